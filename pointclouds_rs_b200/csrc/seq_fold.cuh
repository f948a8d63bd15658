// seq_fold.cuh -- EXACT parallel evaluation of a sequential f32 fold of non-negative values.
//
// The reference sums with `iter().sum::<f32>()`: a strictly left-to-right f32 fold
// (statistical_outlier.rs:54-59, icp.rs:277-280).  Its rounding error at N = 1e5..1e6 is ~1e-5..1e-4
// relative, i.e. larger than the band the SOR mask is allowed to differ in, so the fold ORDER is
// part of the result and has to be reproduced bit for bit -- but a 122 K-long dependent FADD chain
// on one GPU thread costs milliseconds.  This file evaluates the same chain in parallel:
//
//   While the running sum s stays inside one binade [2^E, 2^(E+1)), its ulp u = 2^(E-23) is fixed
//   and s = S*u with a 24-bit integer S.  Adding x >= 0 then is integer arithmetic: with
//   x = (q + r)*u, 0 <= r < 1,   fl(s + x) = (S + q + [r > 1/2] + [r == 1/2 and S+q odd]) * u
//   (round-to-nearest-even).  The only dependence on the running value is the PARITY of S in the
//   tie case, so every element is a map  S -> S + a[S & 1]  with two increments (a0, a1); such maps
//   compose associatively (the output parity is determined by the input parity), hence a block
//   SCAN gives every prefix exactly.  When a prefix reaches 2^24 the sum leaves the binade: that one
//   addition is done with a real FADD and the remainder of the tile is rescanned in the new
//   binade.  Sums double between such crossings, so there are only ~log2(N) of them, and the first
//   ones (short prefixes) are taken by a plain sequential head.
//
// Validated against numpy's sequential float32 accumulation (denormals, ties, overflow) before it
// was ported; on the GPU the SOR statistics are compared bit for bit with the CPU oracle.
#pragma once
#include <cooperative_groups.h>

#include "pcr_internal.cuh"

namespace pcr {

constexpr int kFoldThreads = 1024;
constexpr int kFoldItems = 4;
constexpr int kFoldTile = kFoldThreads * kFoldItems;
constexpr int kFoldHead = 1024;        // elements folded sequentially by thread 0 first (<= kFoldThreads)
constexpr uint32_t kFoldSat = 1u << 25;  // increments saturate here (anything >= 2^24 is a crossing)
constexpr uint32_t kFoldLimit = 1u << 24;

struct FoldMap {
    uint32_t a0, a1;  // increment of S (in ulps) for incoming parity 0 / 1
};

__device__ __forceinline__ uint32_t fold_sat_add(uint32_t a, uint32_t b) {
    uint32_t c = a + b;
    return c > kFoldSat ? kFoldSat : c;
}

// apply f first, then g
__device__ __forceinline__ FoldMap fold_compose(FoldMap f, FoldMap g) {
    FoldMap h;
    h.a0 = fold_sat_add(f.a0, (f.a0 & 1u) ? g.a1 : g.a0);
    h.a1 = fold_sat_add(f.a1, ((f.a1 + 1u) & 1u) ? g.a1 : g.a0);
    return h;
}

// effective exponent (>= -126) and integer mantissa of s >= 0:  s = S * 2^(E - 23)
__device__ __forceinline__ void fold_split(float s, int &E, uint32_t &S) {
    uint32_t b = __float_as_uint(s);
    uint32_t ex = (b >> 23) & 0xffu, m = b & 0x7fffffu;
    if (ex == 0) {
        E = -126;
        S = m;
    } else {
        E = (int)ex - 127;
        S = m | 0x800000u;
    }
}
__device__ __forceinline__ float fold_join(int E, uint32_t S) {
    if (E == -126) return __uint_as_float(S);  // denormal or the first normal binade: bits == S
    return __uint_as_float(((uint32_t)(E + 127) << 23) | (S & 0x7fffffu));
}

// x = (q + r) * 2^(E-23): q (saturated) and kind = 0: r < 1/2 (or exact), 1: r > 1/2, 2: r == 1/2.
// Done in f64, where x / u is exact (a 24-bit mantissa times a power of two): floor and the
// remainder are exact as well.  inv_u = 2^(23-E) is uniform per pass.
__device__ __forceinline__ double fold_inv_ulp(int E) {  // 2^(23-E), E in [-126, 128]
    return __longlong_as_double((long long)(1023 + 23 - E) << 52);
}
__device__ __forceinline__ void fold_decompose(float x, double inv_u, uint32_t &q, int &kind) {
    double y = (double)x * inv_u;  // +inf stays +inf -> saturates
    double fl = floor(y);
    double r = y - fl;
    q = fl >= (double)kFoldSat ? kFoldSat : (uint32_t)fl;
    kind = r > 0.5 ? 1 : (r == 0.5 ? 2 : 0);
}

__device__ __forceinline__ FoldMap fold_element_map(float x, double inv_u) {
    uint32_t q;
    int kind;
    fold_decompose(x, inv_u, q, kind);
    FoldMap m;
    if (kind == 2) {
        m.a0 = fold_sat_add(q, q & 1u);
        m.a1 = fold_sat_add(q, (q + 1u) & 1u);
    } else {
        m.a0 = m.a1 = fold_sat_add(q, kind == 1 ? 1u : 0u);
    }
    return m;
}

// S after adding x in binade E, exact (valid while the result stays below 2^24)
__device__ __forceinline__ uint32_t fold_step(uint32_t S, float x, double inv_u) {
    uint32_t q;
    int kind;
    fold_decompose(x, inv_u, q, kind);
    uint32_t t = fold_sat_add(S, q);
    if (kind == 1 || (kind == 2 && (t & 1u))) t = fold_sat_add(t, 1u);
    return t;
}

struct FoldShared {
    FoldMap warp_tot[32];
    FoldMap cta_total;    // composed map of this CTA's sub-tile (read by the other CTAs of the cluster)
    FoldMap cta_prefix;   // composition of the totals of the lower-ranked CTAs
    uint32_t cross_idx;   // cluster-tile-local index of this CTA's first crossing element
    uint32_t s_before;    // S just before it
    float x_cross;
    uint32_t s_last;      // S after this CTA's last element (valid if nothing crossed before it)
    float s;              // running sum (every CTA of the cluster holds the same value)
    uint32_t done_upto;   // cluster-tile-local: elements below this index are already folded
    uint32_t cnt[32];
    float head[kFoldHead];
};

// Left-to-right f32 sum of fn(v[i]) over the finite v[i], i in [b, e), evaluated by a whole thread
// block CLUSTER (kFoldThreads threads per CTA; a cluster of one CTA works too).  Each CTA scans its
// own sub-tile; the CTAs exchange their composed maps and their first crossing through distributed
// shared memory, so one pass covers cluster_size * kFoldTile elements.  fn must return a
// non-negative value.  Returns the sum in every thread of every CTA; *n_used (if not null) receives
// the number of finite elements.
template <class Fn>
__device__ float cluster_exact_fold(const float *__restrict__ v, size_t b, size_t e, Fn fn, FoldShared &sh, uint32_t *n_used) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), CS = cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t my_cnt = 0;
    // ---- sequential head (every CTA computes it redundantly: no communication) ---------------------
    size_t head_end = b + kFoldHead < e ? b + kFoldHead : e;
    if (tid < kFoldHead) {  // stage through shared memory so that thread 0 never waits on DRAM
        float t = b + tid < head_end ? v[b + tid] : INFINITY;
        bool ok = isfinite(t);
        sh.head[tid] = ok ? fn(t) : -1.0f;  // fn >= 0: a negative entry marks "skip"
        my_cnt += (ok && rank == 0) ? 1u : 0u;
    }
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        const int nh = (int)(head_end - b);
        for (int i = 0; i < nh; i++) {
            float t = sh.head[i];
            if (t >= 0.0f) s = __fadd_rn(s, t);
        }
        sh.s = s;
    }
    __syncthreads();
    // ---- tiles ------------------------------------------------------------------------------------
    const size_t ctile = (size_t)CS * kFoldTile;
    for (size_t tile = head_end; tile < e; tile += ctile) {
        float x[kFoldItems];
        bool use[kFoldItems];
        const uint32_t li0 = rank * kFoldTile + (uint32_t)tid * kFoldItems;  // cluster-tile-local index of item 0
        const size_t base = tile + li0;                                       // thread-contiguous items
#pragma unroll
        for (int j = 0; j < kFoldItems; j++) {
            size_t i = base + j;
            float t = i < e ? v[i] : INFINITY;
            use[j] = isfinite(t);
            x[j] = use[j] ? fn(t) : 0.0f;
            my_cnt += use[j] ? 1u : 0u;
        }
        if (tid == 0) sh.done_upto = 0;
        __syncthreads();
        for (;;) {
            const float s = sh.s;
            if (!isfinite(s)) break;  // overflowed to +inf: stays there (all terms are >= 0); uniform across the cluster
            const uint32_t done = sh.done_upto;
            int E;
            uint32_t S_in;
            fold_split(s, E, S_in);
            const double inv_u = fold_inv_ulp(E);
            // 1. per-thread aggregate map over the pending items
            FoldMap agg = {0u, 0u};
#pragma unroll
            for (int j = 0; j < kFoldItems; j++)
                if (use[j] && li0 + j >= done) agg = fold_compose(agg, fold_element_map(x[j], inv_u));
            // 2. block scan of the aggregates
            FoldMap inc = agg;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                FoldMap u;
                u.a0 = __shfl_up_sync(PCR_FULL, inc.a0, o);
                u.a1 = __shfl_up_sync(PCR_FULL, inc.a1, o);
                if (lane >= o) inc = fold_compose(u, inc);
            }
            if (lane == 31) sh.warp_tot[w] = inc;
            if (tid == 0) sh.cross_idx = 0xffffffffu;
            __syncthreads();
            if (w == 0) {
                FoldMap t = sh.warp_tot[lane], ti = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    FoldMap u;
                    u.a0 = __shfl_up_sync(PCR_FULL, ti.a0, o);
                    u.a1 = __shfl_up_sync(PCR_FULL, ti.a1, o);
                    if (lane >= o) ti = fold_compose(u, ti);
                }
                if (lane == 31) sh.cta_total = ti;
                FoldMap ex;
                ex.a0 = __shfl_up_sync(PCR_FULL, ti.a0, 1);
                ex.a1 = __shfl_up_sync(PCR_FULL, ti.a1, 1);
                if (lane == 0) ex = FoldMap{0u, 0u};
                sh.warp_tot[lane] = ex;
            }
            cluster.sync();  // every CTA's total is visible
            // 3. prefix over the lower-ranked CTAs (warp 0: one lane per CTA, then a shuffle scan)
            if (w == 0) {
                FoldMap t = {0u, 0u};
                if ((unsigned)lane < CS) t = cluster.map_shared_rank(&sh, lane)->cta_total;
                FoldMap ti = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    FoldMap u;
                    u.a0 = __shfl_up_sync(PCR_FULL, ti.a0, o);
                    u.a1 = __shfl_up_sync(PCR_FULL, ti.a1, o);
                    if (lane >= o) ti = fold_compose(u, ti);
                }
                FoldMap ex;
                ex.a0 = __shfl_up_sync(PCR_FULL, ti.a0, 1);
                ex.a1 = __shfl_up_sync(PCR_FULL, ti.a1, 1);
                if (lane == 0) ex = FoldMap{0u, 0u};
                if ((unsigned)lane == rank) sh.cta_prefix = ex;
            }
            __syncthreads();
            FoldMap lane_ex;
            lane_ex.a0 = __shfl_up_sync(PCR_FULL, inc.a0, 1);
            lane_ex.a1 = __shfl_up_sync(PCR_FULL, inc.a1, 1);
            if (lane == 0) lane_ex = FoldMap{0u, 0u};
            const FoldMap pre = fold_compose(sh.cta_prefix, fold_compose(sh.warp_tot[w], lane_ex));
            // 4. walk the own items with the actual running value; find the first crossing
            uint32_t S = fold_sat_add(S_in, (S_in & 1u) ? pre.a1 : pre.a0);
            uint32_t my_cross = 0xffffffffu, S_before = 0;
            float xc = 0.f;
            if (S < kFoldLimit) {
#pragma unroll
                for (int j = 0; j < kFoldItems; j++) {
                    if (use[j] && li0 + j >= done && my_cross == 0xffffffffu) {
                        uint32_t S2 = fold_step(S, x[j], inv_u);
                        if (S2 >= kFoldLimit) {
                            my_cross = li0 + j;
                            S_before = S;
                            xc = x[j];
                        } else {
                            S = S2;
                        }
                    }
                }
            }
            if (my_cross != 0xffffffffu) atomicMin(&sh.cross_idx, my_cross);
            if (tid == kFoldThreads - 1) sh.s_last = S;
            __syncthreads();
            if (my_cross != 0xffffffffu && my_cross == sh.cross_idx) {  // this CTA's first crossing
                sh.s_before = S_before;
                sh.x_cross = xc;
            }
            cluster.sync();  // every CTA's crossing record is visible
            // 5. the cluster-wide first crossing decides the new state (every CTA computes the same)
            if (w == 0) {
                uint32_t c = 0xffffffffu;
                if ((unsigned)lane < CS) c = cluster.map_shared_rank(&sh, lane)->cross_idx;
                const uint32_t cmin = __reduce_min_sync(PCR_FULL, c);
                float ns;
                uint32_t nd = 0;
                if (cmin == 0xffffffffu) {
                    ns = fold_join(E, cluster.map_shared_rank(&sh, CS - 1)->s_last);
                } else {
                    const unsigned owner = __ffs(__ballot_sync(PCR_FULL, c == cmin)) - 1;
                    const FoldShared *o = cluster.map_shared_rank(&sh, owner);
                    ns = __fadd_rn(fold_join(E, o->s_before), o->x_cross);
                    nd = cmin + 1u;
                }
                if (lane == 0) {
                    sh.warp_tot[0].a0 = cmin;  // scratch: lets the other warps see whether to loop
                    sh.s = ns;
                    sh.done_upto = nd;
                }
            }
            cluster.sync();  // all remote reads of this pass are done before anything is overwritten
            if (sh.warp_tot[0].a0 == 0xffffffffu) break;
        }
        __syncthreads();
    }
    // ---- count of finite elements ------------------------------------------------------------------
    if (n_used) {
        uint32_t c = __reduce_add_sync(PCR_FULL, my_cnt);
        if (lane == 0) sh.cnt[w] = c;
        __syncthreads();
        if (w == 0) {
            uint32_t t = __reduce_add_sync(PCR_FULL, sh.cnt[lane]);
            if (lane == 0) sh.cnt[0] = t;
        }
        cluster.sync();
        uint32_t tot = 0;
        for (unsigned r = 0; r < CS; r++) tot += cluster.map_shared_rank(&sh, r)->cnt[0];
        cluster.sync();
        *n_used = tot;
    }
    const float r = sh.s;
    __syncthreads();
    return r;
}


// ------------------------------------------------------------------------------------------------
// The same fold without a cluster-wide pass per binade crossing (round 2).
//
// cluster_exact_fold rescans the rest of its tile after every crossing: ~9 passes of three cluster barriers each per
// fold, 69 us for the two folds of a 119 K-point SOR.  Where the crossings are is known almost exactly beforehand: an
// f64 prefix sum P_i of the terms differs from the f32 running sum by its accumulated rounding (~1e-5 relative), so
//   * an element whose P lies within 2^-10 of a power of two (or straddles one) belongs to a ZONE: the few hundred
//     elements around each crossing, added one by one with real FADDs by one warp;
//   * every other element is added while the sum is inside a binade that is known from P: its map is built for that
//     binade's ulp, and the maps of a SEGMENT (the elements between two zones) compose into one map per CTA;
//   * one warp then walks  segment 0, zone 0, segment 1, zone 1, ...  from the running sum of the head: a segment is
//     one map application per CTA, a zone a short sequential loop.
// Nothing is taken on trust: a segment's maps are applied only if the running sum IS in the predicted binade when the
// segment starts and still is when it ends (S < 2^24: x >= 0, so it then was all the way), every element is in
// exactly one segment or zone (the zones are index ranges that both the map builder and the walker read), and the
// index order of segments and zones is checked.  Any violated check -- or a sum that spans more than kSpecZones
// binades, or an empty head -- returns false and the caller runs cluster_exact_fold.  The result is the same bits
// either way: this only removes passes.
// ------------------------------------------------------------------------------------------------
constexpr int kSpecItems = 8;            // elements per thread and super-tile
constexpr int kSpecTile = kFoldThreads * kSpecItems;
constexpr int kSpecHead = 128;           // elements folded sequentially first (gives the running sum a binade to start from)
constexpr int kSpecZones = 31;           // binades one super-tile may cross: zone k = lane k of a warp, segment k = 0 .. 31
constexpr int kSpecZoneCap = 1024;       // zone elements per super-tile (<= kFoldThreads: one thread fetches one)
constexpr unsigned long long kSpecNearLo = 1ull << 42;                  // mantissa (52 bits) <= 2^-10: just above a power of two
constexpr unsigned long long kSpecNearHi = (1ull << 52) - (1ull << 42);  // >= 1 - 2^-10: just below the next one

struct SpecShared {
    double warp_sum[32];
    double cta_sum, cta_off;
    uint32_t zlo[32], zhi[32];      // this CTA's flagged elements per zone (super-tile-local indices; 0xffffffff / 0 = none)
    uint32_t gzlo[32], gzhi[32];    // cluster-wide
    uint32_t prev_end[32];          // segment k: one past the last index of the zones below it (0 = none)
    uint32_t next_lo[33];           // segment k: first index of the zones from k on (0xffffffff = none)
    uint32_t zoff[33];              // first slot of zone k in zval (prefix of the zone lengths)
    FoldMap warp_seg[32][2];        // per warp: composed maps of its first and last segment (a warp spans more only at a start)
    uint32_t warp_k[32][2];         // ... and their ids (0xffffffff = none)
    FoldMap seg_map[32];            // this CTA's composed map per segment
    uint32_t seg_present;           // bit k: this CTA has elements in segment k
    FoldMap all_map[16][32];        // CTA 0: every CTA's segment maps, gathered for the walk
    uint32_t pmask[32];             // CTA 0: bit r = CTA r has elements in segment k
    float zval[kSpecZoneCap];       // CTA 0: the zone elements' terms (-1: not a term), in index order
    uint32_t fail;
    float s;                        // running sum (walker's result, then copied to every CTA)
    uint32_t ok;
    float head[kSpecHead];
    uint32_t cnt[32];
};

__device__ __forceinline__ int spec_binade(double p) {  // binade of a running sum p >= 0 as an f32 (>= -126)
    const int e = (int)((__double_as_longlong(p) >> 52) & 0x7ff) - 1023;
    return e < -126 ? -126 : e;
}
__device__ __forceinline__ bool spec_near(double p) {
    const unsigned long long m = (unsigned long long)__double_as_longlong(p) & ((1ull << 52) - 1ull);
    return m <= kSpecNearLo || m >= kSpecNearHi;
}
// index of the power of two nearest to p (in log scale)
__device__ __forceinline__ int spec_boundary(double p) {
    const unsigned long long m = (unsigned long long)__double_as_longlong(p) & ((1ull << 52) - 1ull);
    return spec_binade(p) + (m >= (1ull << 51) ? 1 : 0);
}
// ordered composition of the 32 lanes' maps (composition is not commutative): lane l ends with the map of lanes l .. 31
__device__ __forceinline__ FoldMap spec_warp_compose(FoldMap m, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        FoldMap u;
        u.a0 = __shfl_down_sync(PCR_FULL, m.a0, o);
        u.a1 = __shfl_down_sync(PCR_FULL, m.a1, o);
        if (lane + o < 32) m = fold_compose(m, u);
    }
    return m;
}

// The term of an element as a value, not a template parameter: ONE copy of the fold's code serves both SOR folds (the
// kernel runs every instruction once or twice per call on 16 SMs; two inlined copies of this function plus two of the
// fallback were 182 KB of SASS).
struct FoldTerm {
    int squared_deviation;  // 0: v itself, 1: (v - mean)^2 (statistical_outlier.rs:57-59, powi(2))
    float mean;
    __device__ __forceinline__ float operator()(float v) const {
        if (!squared_deviation) return v;
        const float d = __fsub_rn(v, mean);
        return __fmul_rn(d, d);
    }
};

__device__ __noinline__ bool cluster_exact_fold_fast(const float *__restrict__ v, size_t b, size_t e, FoldTerm fn, SpecShared &sh,
                                                     float *result, uint32_t *n_used, bool debug) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), CS = cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t my_cnt = 0;
    long long tk[10];
    int ntk = 0;
#define PCR_FOLD_TICK() \
    if (debug && ntk < 10) tk[ntk++] = clock64();
    PCR_FOLD_TICK()
    // ---- sequential head (every CTA, redundantly) -------------------------------------------------------------
    const size_t head_end = b + kSpecHead < e ? b + kSpecHead : e;
    if (tid < kSpecHead) {
        const float t = b + tid < head_end ? v[b + tid] : INFINITY;
        const bool okv = isfinite(t);
        sh.head[tid] = okv ? fn(t) : -1.0f;
        my_cnt += (okv && rank == 0) ? 1u : 0u;
    }
    if (tid == 0) sh.fail = 0u;
    __syncthreads();
    if (tid == 0) {
        float s = 0.0f;
        const int nh = (int)(head_end - b);
        for (int i = 0; i < nh; i++) {
            const float t = sh.head[i];
            if (t >= 0.0f) s = __fadd_rn(s, t);
        }
        sh.s = s;
    }
    __syncthreads();
    bool ok = true;
    PCR_FOLD_TICK()  // 1: head
    const size_t ctile = (size_t)CS * kSpecTile;
    for (size_t tile = head_end; tile < e && ok; tile += ctile) {
        const float s_start = sh.s;
        if (!(s_start > 0.0f) || !isfinite(s_start)) {  // no binade to start from (uniform across the cluster)
            ok = false;
            break;
        }
        const int E0 = spec_binade((double)s_start);
        float x[kSpecItems];
        bool use[kSpecItems];
        const uint32_t li0 = rank * kSpecTile + (uint32_t)tid * kSpecItems;
        const size_t base = tile + li0;
        double tsum = 0.0;
#pragma unroll
        for (int j = 0; j < kSpecItems; j++) {
            const size_t i = base + j;
            const float t = i < e ? v[i] : INFINITY;
            use[j] = isfinite(t);
            x[j] = use[j] ? fn(t) : 0.0f;
            my_cnt += use[j] ? 1u : 0u;
            tsum += (double)x[j];
        }
        if (tid < 32) {
            sh.zlo[tid] = 0xffffffffu;
            sh.zhi[tid] = 0u;
            sh.pmask[tid] = 0u;
        }
        if (tid == 0) sh.seg_present = 0u;
        // ---- A: f64 prefix of the terms over the cluster ---------------------------------------------------------
        double inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u = __shfl_up_sync(PCR_FULL, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) sh.warp_sum[w] = inc;
        __syncthreads();
        if (w == 0) {
            double t = sh.warp_sum[lane], ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(PCR_FULL, ti, o);
                if (lane >= o) ti += u;
            }
            if (lane == 31) sh.cta_sum = ti;
            sh.warp_sum[lane] = ti - t;  // exclusive
        }
        PCR_FOLD_TICK()  // 2: loads + block prefix
        cluster.sync();  // (1) every CTA's sum is visible
        if (w == 0) {
            double t = 0.0;
            if ((unsigned)lane < CS) t = cluster.map_shared_rank(&sh, lane)->cta_sum;
            double ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double u = __shfl_up_sync(PCR_FULL, ti, o);
                if (lane >= o) ti += u;
            }
            if ((unsigned)lane == rank) sh.cta_off = ti - t;
        }
        __syncthreads();
        double P = (double)s_start + sh.cta_off + sh.warp_sum[w] + (inc - tsum);  // sum before this thread's first item
        int sid[kSpecItems];
        bool flagged[kSpecItems];
        uint32_t code = 0;  // first failed check (debug: PCR_FOLD_DEBUG prints it)
        int zcur = -1;      // zone of this thread's flagged items so far, and their index range
        uint32_t zl = 0xffffffffu, zh = 0u;
#pragma unroll
        for (int j = 0; j < kSpecItems; j++) {
            const double Pprev = P;
            P += (double)x[j];
            sid[j] = spec_binade(Pprev) - E0;
            flagged[j] = false;
            if (use[j]) {
                const bool np = spec_near(Pprev), nn = spec_near(P);
                flagged[j] = np || nn || spec_binade(Pprev) != spec_binade(P);
                if (!(P < 1e300)) code = max(code, 1u);  // +inf term
                if (flagged[j]) {
                    const int zid = (np ? spec_boundary(Pprev) : spec_boundary(P)) - E0 - 1;
                    if (zid < 0 || zid >= kSpecZones) {
                        code = max(code, 2u);
                    } else {
                        // (a zone's few hundred elements sit in a few dozen neighbouring threads: the thread keeps a running
                        // record, the warp merges the records, one pair of atomics per warp and zone)
                        if (zcur >= 0 && zcur != zid) {  // (rare: two zones inside one thread's items)
                            atomicMin(&sh.zlo[zcur], zl);
                            atomicMax(&sh.zhi[zcur], zh);
                            zcur = -1;
                        }
                        if (zcur < 0) {
                            zcur = zid;
                            zl = li0 + j;
                        }
                        zh = li0 + j;
                    }
                } else if (sid[j] < 0 || sid[j] > kSpecZones) {
                    code = max(code, 3u);
                }
            }
        }
        {
            const unsigned grp = __match_any_sync(PCR_FULL, zcur);
            const uint32_t gl = __reduce_min_sync(grp, zl), gh = __reduce_max_sync(grp, zh);
            if (zcur >= 0 && lane == __ffs(grp) - 1) {
                atomicMin(&sh.zlo[zcur], gl);
                atomicMax(&sh.zhi[zcur], gh);
            }
        }
        if (code) atomicMax(&sh.fail, code);
        PCR_FOLD_TICK()  // 3: flags
        cluster.sync();  // (2) every CTA's zone ranges are visible
        // cluster-wide zone ranges: thread (r, k) fetches CTA r's range of zone k -- ONE distributed-shared-memory round trip
        // (a lane looping over the 16 CTAs paid 32 of them back to back: 15 K cycles, measured)
        if (tid < 32) {
            sh.gzlo[tid] = 0xffffffffu;
            sh.gzhi[tid] = 0u;
        }
        __syncthreads();
        if ((unsigned)tid < CS * 32u && lane < kSpecZones) {
            const SpecShared *o = cluster.map_shared_rank(&sh, tid >> 5);
            const uint32_t rl = o->zlo[lane], rh = o->zhi[lane];
            if (rl != 0xffffffffu) {
                atomicMin(&sh.gzlo[lane], rl);
                atomicMax(&sh.gzhi[lane], rh);
            }
        }
        __syncthreads();
        if (w == 0) {  // lane k = zone k: order check, and what every segment needs to know about its neighbours
            const uint32_t lo = sh.gzlo[lane], hi = sh.gzhi[lane];
            const bool nonempty = lo != 0xffffffffu;
            const uint32_t endv = nonempty ? hi + 1u : 0u, len = nonempty ? hi - lo + 1u : 0u;
            uint32_t pe = endv, po = len, nl = lo;  // inclusive scans: max of the ends, sum of the lengths, (suffix) min of the starts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t a = __shfl_up_sync(PCR_FULL, pe, o), c = __shfl_up_sync(PCR_FULL, po, o), d = __shfl_down_sync(PCR_FULL, nl, o);
                if (lane >= o) {
                    pe = max(pe, a);
                    po += c;
                }
                if (lane + o < 32) nl = min(nl, d);
            }
            uint32_t pe_ex = __shfl_up_sync(PCR_FULL, pe, 1);  // exclusive: the zones below zone / segment `lane`
            if (lane == 0) pe_ex = 0u;
            if (nonempty && lo < pe_ex) atomicMax(&sh.fail, 4u);  // zones out of order
            sh.prev_end[lane] = pe_ex;
            sh.next_lo[lane] = nl;
            sh.zoff[lane] = po - len;
            if (lane == 31) {
                sh.next_lo[32] = 0xffffffffu;
                sh.zoff[32] = po;
            }
        }
        __syncthreads();
        // ---- B: segment of every element that is in no zone; per-CTA composed map of every segment --------------------
        code = 0;
        uint32_t kfirst = 0xffffffffu, klast = 0u;  // this thread's segments (items are in index order: ids ascend)
#pragma unroll
        for (int j = 0; j < kSpecItems; j++) {
            if (!use[j]) continue;
            const uint32_t li = li0 + j;
            // the only zones an element can lie in are the ones next to its own segment
            bool inz = false;
#pragma unroll
            for (int d = -1; d <= 0; d++) {
                const int z = sid[j] + d;
                if (z >= 0 && z < kSpecZones && sh.gzlo[z] != 0xffffffffu && li >= sh.gzlo[z] && li <= sh.gzhi[z]) inz = true;
            }
            if (flagged[j] && !inz) code = max(code, 5u);  // (cannot happen: a flagged element made its zone's range)
            if (inz) {
                use[j] = false;  // the walker adds it
            } else if (sid[j] < 0 || sid[j] > kSpecZones) {
                code = max(code, 6u);
                use[j] = false;
            } else {
                if (li < sh.prev_end[sid[j]] || li >= sh.next_lo[sid[j]]) code = max(code, 7u);  // out of order
                kfirst = min(kfirst, (uint32_t)sid[j]);
                klast = max(klast, (uint32_t)sid[j]);
            }
        }
        if (code) atomicMax(&sh.fail, code);
        {
            // per warp: the composed map of its first and of its last segment (ordered reductions, no block barrier); a
            // warp whose 256 elements touch a third segment (only where the sum still doubles within a few elements) fails
            const uint32_t wfirst = __reduce_min_sync(PCR_FULL, kfirst), wlast = __reduce_max_sync(PCR_FULL, klast);
            bool other = false;
            FoldMap mf = {0u, 0u}, ml = {0u, 0u};
            if (wfirst != 0xffffffffu) {
                const double iu_f = fold_inv_ulp(E0 + (int)wfirst), iu_l = fold_inv_ulp(E0 + (int)wlast);
#pragma unroll
                for (int j = 0; j < kSpecItems; j++)
                    if (use[j]) {
                        if ((uint32_t)sid[j] == wfirst) mf = fold_compose(mf, fold_element_map(x[j], iu_f));
                        else if ((uint32_t)sid[j] == wlast) ml = fold_compose(ml, fold_element_map(x[j], iu_l));
                        else other = true;
                    }
                mf = spec_warp_compose(mf, lane);
                if (wlast != wfirst) ml = spec_warp_compose(ml, lane);
            }
            if (__any_sync(PCR_FULL, other)) atomicMax(&sh.fail, 9u);
            if (lane == 0) {
                sh.warp_k[w][0] = wfirst;
                sh.warp_k[w][1] = (wfirst != 0xffffffffu && wlast != wfirst) ? wlast : 0xffffffffu;
                sh.warp_seg[w][0] = mf;
                sh.warp_seg[w][1] = ml;
            }
        }
        __syncthreads();
        if (w == 0) {  // lane = warp: compose the warps' maps per segment, in warp order
            const uint32_t k0 = sh.warp_k[lane][0], k1 = sh.warp_k[lane][1];
            const FoldMap m0 = sh.warp_seg[lane][0], m1 = sh.warp_seg[lane][1];
            uint32_t present = 0u;
            if (k0 != 0xffffffffu) present |= 1u << k0;
            if (k1 != 0xffffffffu) present |= 1u << k1;
            present = __reduce_or_sync(PCR_FULL, present);
            if (lane == 0) sh.seg_present = present;
            uint32_t rest = present;
            while (rest) {
                const uint32_t k = (uint32_t)__ffs(rest) - 1u;
                rest &= rest - 1u;
                FoldMap m = {0u, 0u};
                if (k0 == k) m = m0;
                else if (k1 == k) m = m1;
                m = spec_warp_compose(m, lane);
                if (lane == 0) sh.seg_map[k] = m;
            }
        }
        PCR_FOLD_TICK()  // 4: zone merge + segment maps
        cluster.sync();  // (3) every CTA's segment maps are visible
        // ---- C: the walk (CTA 0).  Everything it needs is first gathered into CTA 0's own shared memory by all of its
        // threads -- the other CTAs' maps (distributed shared memory) and the zone elements (global memory): read one by
        // one from inside the sequential walk, those latencies were the whole kernel ------------------------------------
        if (rank == 0) {
            if ((unsigned)tid < CS * 32u) {
                const unsigned r = tid >> 5, k = tid & 31;
                const SpecShared *o = cluster.map_shared_rank(&sh, r);
                const bool present = (o->seg_present >> k) & 1u;
                FoldMap m = {0u, 0u};
                if (present) m = o->seg_map[k];
                sh.all_map[r][k] = m;
                if (present) atomicOr(&sh.pmask[k], 1u << r);
                if (k == 0) atomicMax(&sh.fail, o->fail);
            }
            const uint32_t Z = sh.zoff[32];
            if ((uint32_t)tid < Z && Z <= (uint32_t)kSpecZoneCap) {
                int k = 0;
                while (sh.zoff[k + 1] <= (uint32_t)tid) k++;
                const size_t gi = tile + sh.gzlo[k] + ((uint32_t)tid - sh.zoff[k]);
                float t = -1.0f;
                if (gi < e) {
                    const float raw = v[gi];
                    if (isfinite(raw)) t = fn(raw);
                }
                sh.zval[tid] = t;
            }
            __syncthreads();
            if (w == 0) {
                float s = s_start;
                uint32_t why = sh.fail;
                if (Z > (uint32_t)kSpecZoneCap) why = max(why, 8u);
                bool fail = why != 0u;
                // (lane k holds what step k needs: no shared-memory latency inside the sequential part except the maps)
                const uint32_t my_pmask = sh.pmask[lane], my_zb = sh.zoff[lane], my_ze = sh.zoff[lane + 1];
                unsigned steps = __ballot_sync(PCR_FULL, my_pmask != 0u || my_ze > my_zb);
                while (steps && !fail) {
                    const int k = __ffs(steps) - 1;
                    steps &= steps - 1u;
                    unsigned rest = __shfl_sync(PCR_FULL, my_pmask, k);
                    if (rest) {  // segment k: one map per CTA, in rank order
                        int E;
                        uint32_t S;
                        fold_split(s, E, S);
                        if (E != E0 + k) {
                            fail = true;
                            why = 100u + (uint32_t)k;
                        } else {
                            while (rest) {
                                const int r = __ffs(rest) - 1;
                                rest &= rest - 1u;
                                const FoldMap m = sh.all_map[r][k];
                                S = fold_sat_add(S, (S & 1u) ? m.a1 : m.a0);
                            }
                            if (S >= kFoldLimit) {
                                fail = true;
                                why = 200u + (uint32_t)k;
                            } else {
                                s = fold_join(E, S);
                            }
                        }
                    }
                    // zone k: real additions, one by one (32 terms per load: lane l holds the l-th)
                    const uint32_t zb = __shfl_sync(PCR_FULL, my_zb, k), ze = __shfl_sync(PCR_FULL, my_ze, k);
                    if (!fail) {
                        for (uint32_t t0 = zb; t0 < ze; t0 += 32) {
                            const float mine = t0 + (uint32_t)lane < ze ? sh.zval[t0 + lane] : -1.0f;
                            // (a slot past the end or of a skipped element holds -1: adding is a select, not a branch)
#pragma unroll
                            for (int l = 0; l < 32; l++) {
                                const float tl = __shfl_sync(PCR_FULL, mine, l);
                                const float s2 = __fadd_rn(s, tl);
                                s = tl >= 0.0f ? s2 : s;
                            }
                        }
                    }
                }
                if (lane == 0) {
                    if (fail && debug)
                        printf("[pcr] fast fold gave up: check %u (E0 %d, s_start %g, s %g, tile %llu)\n", why, E0, s_start, s, (unsigned long long)tile);
                    sh.s = s;
                    sh.ok = fail ? 0u : 1u;
                }
            }
        }
        PCR_FOLD_TICK()  // 5: gather + walk
        cluster.sync();  // (4) the result is in CTA 0
        const SpecShared *c0 = cluster.map_shared_rank(&sh, 0);
        const float ns = c0->s;
        const bool good = c0->ok != 0u;
        cluster.sync();  // every CTA has read it before CTA 0 moves on
        if (tid == 0) sh.s = ns;
        __syncthreads();
        ok = good;
        PCR_FOLD_TICK()  // 6
    }
    if (n_used) {
        uint32_t c = __reduce_add_sync(PCR_FULL, my_cnt);
        if (lane == 0) sh.cnt[w] = c;
        __syncthreads();
        if (w == 0) {
            uint32_t t = __reduce_add_sync(PCR_FULL, sh.cnt[lane]);
            if (lane == 0) sh.cnt[0] = t;
        }
        cluster.sync();
        uint32_t tot = 0;
        for (unsigned r = 0; r < CS; r++) tot += cluster.map_shared_rank(&sh, r)->cnt[0];
        cluster.sync();
        *n_used = tot;
    }
    if (debug && rank == 0 && tid == 0) {
        printf("[pcr] fast fold ticks:");
        for (int i = 1; i < ntk; i++) printf(" %lld", tk[i] - tk[i - 1]);
        printf("\n");
    }
#undef PCR_FOLD_TICK
    *result = sh.s;
    __syncthreads();
    return ok;
}

}  // namespace pcr
