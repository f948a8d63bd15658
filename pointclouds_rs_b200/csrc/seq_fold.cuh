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

}  // namespace pcr
