// knn_search.cuh -- the grid ring search shared by the KNN / SOR / normals kernels (warp per query)
// and by the ICP 1-NN kernel (thread per query).  Replaces kiddo's nearest_n behind
// KdTree::knn / knn_indices (crates/spatial/src/kdtree.rs:64-96).
//
// Search rule (mirrored on the CPU by oracle/orc_grid_knn_model, which is how it was validated):
//   1. scan the 3x3x3 cells around the query's cell; every row of 3 cells along the fast axis is
//      one contiguous run of the cell-sorted array;
//   2. after shell R, everything not scanned lies beyond a face of the scanned cube; with f = the
//      query's fractional position in its cell, the nearest such face is (R + min f) * h away.
//      Stop iff the k-th best d^2 is STRICTLY below (that distance)^2 * (1 - 1e-6): the slack
//      covers the rounding of f32 d^2 (<= 3 ulp) and of the f64 cell assignment, and strictness
//      keeps an equal-distance, lower-index point outside the cube from being missed;
//   3. otherwise scan shell R+1 (Chebyshev distance exactly R+1).  After `max_rings` shells the
//      query is DEFERRED (returns false): the caller re-runs it on a coarser level of the grid
//      (knn.cu: level l+1 has 8x the cell size), so isolated outliers -- the very points SOR exists
//      to find -- never walk hundreds of empty shells.  Only on the coarsest level does a query
//      fall back to scanning its whole frame, pruned by the k-th best found so far.
// Candidates are ranked by the packed (d^2 bits, original index) key, so the result does not
// depend on the scan order.
#pragma once
#include "pcr_internal.cuh"
#include "sortnet32.cuh"

#include <type_traits>

namespace pcr {

constexpr int kMaxRings = 12;         // thread-per-query search (ICP) and the coarsest level
constexpr int kLevelRings = 4;        // shells scanned per grid level before a query is deferred
constexpr int kMaxLevels = 4;
constexpr uint32_t kBruteFrame = 96;  // frames this small are scanned directly
constexpr uint32_t kSelHistMaxN = 384;  // more candidates than this in the 27 cells: a volume (dense object), not a surface

// Squared lower bound (metres^2) on the distance from the query to anything outside the scanned cube
// of radius R cells around cell (c0,c1,c2); f = fractional position of the query in that cell
// (< 0 or >= 1 when the query lies outside the grid and the cell was clamped).  Beyond an open face
// of axis a a point is at least (R + f)h away along a AND at least gap_b away along every other axis
// b on which the query is outside the grid -- without the gap term a query 20 m outside the cloud
// would need 20 m worth of rings to prove its neighbour optimal.  Returns false if no face is open
// (the whole grid has been scanned).
__device__ __forceinline__ bool ring_bound2(const GridDesc &g, int c0, int c1, int c2, double f0, double f1, double f2, int R,
                                            double &bound2) {
    const double gp0 = f0 < 0.0 ? -f0 : (f0 > 1.0 ? f0 - 1.0 : 0.0);
    const double gp1 = f1 < 0.0 ? -f1 : (f1 > 1.0 ? f1 - 1.0 : 0.0);
    const double gp2 = f2 < 0.0 ? -f2 : (f2 > 1.0 ? f2 - 1.0 : 0.0);
    const double o0 = gp1 * gp1 + gp2 * gp2, o1 = gp0 * gp0 + gp2 * gp2, o2 = gp0 * gp0 + gp1 * gp1;
    const double r = (double)R;
    double b = 1e300;
    bool open = false;
    if (c0 - R > 0) { open = true; double d = r + f0; b = fmin(b, d * d + o0); }
    if (c0 + R < g.dims[0] - 1) { open = true; double d = r + 1.0 - f0; b = fmin(b, d * d + o0); }
    if (c1 - R > 0) { open = true; double d = r + f1; b = fmin(b, d * d + o1); }
    if (c1 + R < g.dims[1] - 1) { open = true; double d = r + 1.0 - f1; b = fmin(b, d * d + o1); }
    if (c2 - R > 0) { open = true; double d = r + f2; b = fmin(b, d * d + o2); }
    if (c2 + R < g.dims[2] - 1) { open = true; double d = r + 1.0 - f2; b = fmin(b, d * d + o2); }
    bound2 = b * g.h * g.h;
    return open;
}

// distance, in cells, from a query at fraction f of its cell to the cells `e` columns away along one axis
__device__ __forceinline__ float axis_gap(int e, float f) {
    const float d = e > 0 ? (float)e - f : (e < 0 ? f - (float)e - 1.0f : 0.0f);
    return fmaxf(d, 0.0f);
}

// ------------------------------------------------------------------------------------------------
// Top-k containers.  Both keep the k best keys in ascending order.
// ------------------------------------------------------------------------------------------------

// k <= 32: lane j holds the j-th best key in a register; insertion = ballot + shuffle.
struct RegTopK {
    unsigned long long K;    // this lane's entry
    unsigned long long thr;  // acceptance threshold (uniform): candidates with key < thr enter
    unsigned long long cap;  // upper limit of thr (used by the whole-frame fallback)
    int kk, lane;

    __device__ __forceinline__ void reset(unsigned long long cap_) {
        K = PCR_EMPTY_KEY;
        cap = cap_;
        thr = cap_;
    }
    __device__ __forceinline__ bool full() const { return __shfl_sync(PCR_FULL, K, kk - 1) != PCR_EMPTY_KEY; }
    __device__ __forceinline__ unsigned long long kth() const { return __shfl_sync(PCR_FULL, K, kk - 1); }
    __device__ __forceinline__ void insert(unsigned long long x) {
        int pos = __popc(__ballot_sync(PCR_FULL, K < x));
        unsigned long long up = __shfl_up_sync(PCR_FULL, K, 1);
        if (lane > pos) K = up;
        else if (lane == pos) K = x;
        unsigned long long kth_ = __shfl_sync(PCR_FULL, K, kk - 1);
        thr = kth_ < cap ? kth_ : cap;
    }
    __device__ __forceinline__ int count() const {
        int c = __popc(__ballot_sync(PCR_FULL, K != PCR_EMPTY_KEY));
        return c < kk ? c : kk;
    }
};

// k <= PCR_MAX_K: sorted list in shared memory (one per warp), cooperative insertion.
struct SmemTopK {
    unsigned long long *s;  // kk entries
    unsigned long long thr, cap;
    int kk, lane, n;

    __device__ __forceinline__ void reset(unsigned long long cap_) {
        n = 0;
        cap = cap_;
        thr = cap_;
    }
    __device__ __forceinline__ bool full() const { return n == kk; }
    __device__ __forceinline__ unsigned long long kth() const { return n == kk ? s[kk - 1] : PCR_EMPTY_KEY; }
    __device__ __forceinline__ void insert(unsigned long long x) {
        int c = 0;
        for (int i = lane; i < n; i += 32) c += s[i] < x ? 1 : 0;
        int pos = __reduce_add_sync(PCR_FULL, c);
        int last = n < kk ? n : kk - 1;  // entries [pos, last) move up by one
        for (int top = last; top > pos; top -= 32) {
            int i = top - 1 - lane;
            unsigned long long v = 0;
            if (i >= pos) v = s[i];
            __syncwarp();
            if (i >= pos) s[i + 1] = v;
            __syncwarp();
        }
        if (lane == 0) s[pos] = x;
        __syncwarp();
        if (n < kk) n++;
        unsigned long long kth_ = n == kk ? s[kk - 1] : PCR_EMPTY_KEY;
        thr = kth_ < cap ? kth_ : cap;
    }
    __device__ __forceinline__ int count() const { return n; }
};

// offer one batch (one candidate key per lane, EMPTY for idle lanes) to the top-k
template <class TopK>
__device__ __forceinline__ void offer_batch(TopK &tk, unsigned long long key) {
    unsigned m = __ballot_sync(PCR_FULL, key < tk.thr);
    while (m) {
        int s = __ffs(m) - 1;
        tk.insert(__shfl_sync(PCR_FULL, key, s));
        // re-test the remaining lanes against the tightened threshold
        m = __ballot_sync(PCR_FULL, key < tk.thr) & ~((2u << s) - 1u);
    }
}

// One contiguous run [b, e) of the cell-sorted array against the query, 32 candidates per batch.
// Long runs are taken four batches at a time with all four loads issued up front (the insertion
// loop is data dependent, so without this a warp has a single load in flight).
template <class TopK>
__device__ __forceinline__ void scan_run(TopK &tk, const float4 *__restrict__ pts, uint32_t b, uint32_t e, float qx,
                                         float qy, float qz) {
    uint32_t base = b;
    for (; base + 128 <= e; base += 128) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; u++) p[u] = __ldg(&pts[base + u * 32 + tk.lane]);
#pragma unroll
        for (int u = 0; u < 4; u++)
            offer_batch(tk, make_key(dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z), __float_as_uint(p[u].w)));
    }
    for (; base < e; base += 32) {
        uint32_t i = base + tk.lane;
        unsigned long long key = PCR_EMPTY_KEY;
        if (i < e) {
            float4 p = __ldg(&pts[i]);
            key = make_key(dist2_exact(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w));
        }
        offer_batch(tk, key);
    }
}

// lane-parallel lookup of up to 32 runs, then the warp streams the non-empty ones
// k-th best so far in cell units^2 with the pruning slack of scan_row_pruned (+inf while the list is short)
template <class TopK>
__device__ __forceinline__ float warp_tau_cells(const TopK &tk, float inv_h2) {
    const unsigned long long kth = tk.kth();
    return kth == PCR_EMPTY_KEY ? INFINITY : key_d2(kth) * inv_h2 * (1.0f + 1e-5f) + 1e-8f;
}

// gp = squared gap (cell units) between the query and this lane's run: a run further away than the current
// k-th best is skipped when its turn comes (the list has usually tightened since the run was looked up)
template <class TopK>
__device__ __forceinline__ void scan_runs(TopK &tk, const float4 *__restrict__ pts, uint32_t b, uint32_t e,
                                          unsigned first_mask, float qx, float qy, float qz, float gp, float inv_h2) {
    unsigned ne = __ballot_sync(PCR_FULL, e > b);
    unsigned pri = ne & first_mask;  // runs to take first (the query's own row: tightens thr early)
    ne &= ~first_mask;
    while (pri) {
        int s = __ffs(pri) - 1;
        pri &= pri - 1;
        scan_run(tk, pts, __shfl_sync(PCR_FULL, b, s), __shfl_sync(PCR_FULL, e, s), qx, qy, qz);
    }
    // Nearest run first (one redux per run on gap bits | lane): the list tightens as early as it can, and the runs beyond
    // the current k-th best are skipped when their turn comes.  A point metres above a
    // dense surface used to stream the whole footprint of the shell in which the surface first appears (the rows were
    // taken in lane order, the far corner first); now it reads the row below it and little else.
    uint32_t order = ((ne >> tk.lane) & 1u) ? ((__float_as_uint(gp) & ~31u) | (uint32_t)tk.lane) : 0xffffffffu;  // (gp >= 0: its bits order like its value)
    for (;;) {
        const uint32_t m = __reduce_min_sync(PCR_FULL, order);
        if (m == 0xffffffffu) break;
        const int s = (int)(m & 31u);
        if (tk.lane == s) order = 0xffffffffu;
        if (__shfl_sync(PCR_FULL, gp, s) > warp_tau_cells(tk, inv_h2)) continue;  // (not "break": the order key drops 5 bits of the gap)
        scan_run(tk, pts, __shfl_sync(PCR_FULL, b, s), __shfl_sync(PCR_FULL, e, s), qx, qy, qz);
    }
}

// Trims a row's fast-axis cell range [z0, z1] to the cells that can hold a candidate within the current k-th
// best (same bound and slack as scan_row_pruned); returns false if nothing is left.
__device__ __forceinline__ bool trim_row(float tau_c, float gp, int c2, float f2, int &z0, int &z1) {
    if (gp > tau_c) return false;
    if (tau_c < 1e12f && fabsf(f2) < 1e4f) {
        const float w = sqrtf(tau_c - gp) * (1.0f + 1e-5f) + 1e-5f;
        z0 = max(z0, c2 + (int)floorf(f2 - w));
        z1 = min(z1, c2 + (int)floorf(f2 + w));
    }
    return z0 <= z1;
}

// ------------------------------------------------------------------------------------------------
// Selection front-end of the warp-per-query search (round 2): the 3 x 3 x 3 cube of a query without a single
// insertion.  The 9 rows of the cube are ONE flat candidate list (prefix sums of the run lengths in shared memory, a
// 4-step search maps a flat index to its run), the warp walks it 32 candidates at a time -- every lane busy, whatever
// the run lengths -- twice:
//   pass 1  d^2 -> one of 64 linear bins, counted with shared-memory atomics; a warp scan of the counters finds the
//           first bin edge with at least kk candidates below it;
//   pass 2  the candidates up to that edge (kk plus the rest of one bin, <= 32) are compacted into shared memory with
//           a ballot,
// then one bitonic sort across the lanes on the packed (d^2, index) keys leaves lane j with the j-th best: exactly
// the state RegTopK keeps, so the ring rule and the outer shells continue as before.  More than 32 candidates up to the
// edge (ties, a kk-th beyond the last bin) or kk > 24: the function declines and the caller scans the rows the old
// way.  The bins span four times the kk-th d^2 of a uniform sheet (three times that of a uniform volume for more than
// kSelHistMaxN candidates) with as many points in the cube: a wrong guess only costs the fallback.
// Why a warp per query at all: one thread per query leaves a B200 with 6 warps per scheduler on a 120 K-point frame
// (and the 80 registers + 32 KB of its selection buffers cap it there even on an 8 M-point batch); a warp that issues
// one instruction every ~17 cycles cannot hide its own latencies.  This form needs ~40 registers and 0.7 KB per warp.
// ------------------------------------------------------------------------------------------------
constexpr int kWarpSelMaxK = 24;
struct WarpSelScratch {
    uint32_t hist[64];
    unsigned long long keys[32];
    uint32_t run_end[16];   // inclusive prefix of the run lengths (entries 9..15: 0xffffffff)
    uint32_t run_base[16];  // pts index of flat candidate 0 of the run, minus the run's first flat index
};

__device__ __forceinline__ uint32_t warp_sel_locate(const WarpSelScratch &S, uint32_t i) {  // pts index of flat candidate i
    uint32_t lo = 0;
    if (S.run_end[lo + 7] <= i) lo += 8;
    if (S.run_end[lo + 3] <= i) lo += 4;
    if (S.run_end[lo + 1] <= i) lo += 2;
    if (S.run_end[lo] <= i) lo += 1;
    return S.run_base[lo] + i;
}

// b, e: this lane's run (lanes 0..8, empty elsewhere).  Returns true if tk now holds the exact kk best of the 9 runs.
__device__ __forceinline__ bool warp_select_cube(RegTopK &tk, WarpSelScratch &S, const float4 *__restrict__ pts, uint32_t b, uint32_t e,
                                                 float qx, float qy, float qz, float h2) {
    const int lane = tk.lane, kk = tk.kk;
    const uint32_t n = e - b;
    uint32_t inc = n;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const uint32_t t = __shfl_up_sync(PCR_FULL, inc, d);
        if (lane >= d) inc += t;
    }
    const uint32_t N = __shfl_sync(PCR_FULL, inc, 15);  // (runs only in lanes 0..8)
    if (N == 0) return true;  // nothing in the cube: tk stays empty
    __syncwarp();
    if (lane < 16) {
        S.run_end[lane] = lane < 9 ? inc : 0xffffffffu;
        S.run_base[lane] = b - (inc - n);
    }
    S.hist[lane] = 0;
    S.hist[lane + 32] = 0;
    __syncwarp();
    unsigned long long key = PCR_EMPTY_KEY;
    if (N <= 32) {  // one batch: sort it
        if ((uint32_t)lane < N) {
            const float4 p = __ldg(&pts[warp_sel_locate(S, lane)]);
            const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
            if (d2 == d2) key = make_key(d2, __float_as_uint(p.w));  // (NaN: a tombstoned point)
        }
    } else {
        const float q = (float)kk / (float)N;
        const float r2 = N > kSelHistMaxN ? h2 * cbrtf(6.4456f * q) * cbrtf(6.4456f * q) : 2.8648f * h2 * q;
        const float scale = 64.0f / ((N > kSelHistMaxN ? 3.0f : 4.0f) * r2);
        // four loads in flight per lane (a dense query reads thousands of candidates from L2: one dependent load per step was
        // the whole pass); the open last bin is not counted at all -- most candidates of a volume land there, and 32 lanes
        // incrementing one shared word serialise.  (Measured and dropped: widening / refining the range in further passes as
        // the tile kernel does -- 64 -> 69 us for the dense class; what the first range misses goes to the insertion scan.)
        for (uint32_t i0 = lane; i0 < N; i0 += 128) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + 32u * u;
                p[u] = __ldg(&pts[warp_sel_locate(S, i < N ? i : N - 1u)]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float d2 = dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z);
                const int bin = __float2int_rz(fminf(__fmul_rn(d2, scale), 63.0f));  // (NaN lands in bin 63)
                if (i0 + 32u * u < N && bin < 63) atomicAdd(&S.hist[bin], 1u);
            }
        }
        __syncwarp();
        // first bin edge with at least kk candidates below it: lane l owns bins 2l, 2l+1
        const uint32_t c0 = S.hist[2 * lane], c1 = lane < 31 ? S.hist[2 * lane + 1] : 0u;  // (bin 63 is the open one)
        uint32_t cum = c0 + c1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(PCR_FULL, cum, d);
            if (lane >= d) cum += t;
        }
        const unsigned reached = __ballot_sync(PCR_FULL, cum >= (uint32_t)kk);
        if (!reached) return false;  // the kk-th best lies beyond the bins
        const int L = __ffs(reached) - 1;
        const uint32_t cumL = __shfl_sync(PCR_FULL, cum, L), c1L = __shfl_sync(PCR_FULL, c1, L);
        const bool first_half = cumL - c1L >= (uint32_t)kk;
        const int bstar = 2 * L + (first_half ? 0 : 1);
        const uint32_t A = first_half ? cumL - c1L : cumL;
        if (A > 32u) return false;  // too many up to that edge (ties / a dense bin): the caller's insertion scan handles it
        uint32_t base = 0;
        for (uint32_t i0 = 0; i0 < N; i0 += 128) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t i = i0 + 32u * u + lane;
                p[u] = __ldg(&pts[warp_sel_locate(S, i < N ? i : N - 1u)]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float d2 = dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z);
                const bool take = i0 + 32u * u + lane < N && d2 == d2 && __float2int_rz(fminf(__fmul_rn(d2, scale), 63.0f)) <= bstar;
                const unsigned m = __ballot_sync(PCR_FULL, take);
                if (take) S.keys[base + __popc(m & ((1u << lane) - 1u))] = make_key(d2, __float_as_uint(p[u].w));
                base += __popc(m);
            }
        }
        __syncwarp();
        if ((uint32_t)lane < base) key = S.keys[lane];
    }
    // bitonic sort across the lanes, ascending
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long other = __shfl_xor_sync(PCR_FULL, key, j);
            const bool keep_min = ((lane & k) == 0) == ((lane & j) == 0);
            key = (other < key) == keep_min ? other : key;
        }
    tk.K = lane < kk ? key : PCR_EMPTY_KEY;
    const unsigned long long kth_ = __shfl_sync(PCR_FULL, tk.K, kk - 1);
    tk.thr = kth_ < tk.cap ? kth_ : tk.cap;
    return true;
}

// Warp-cooperative exact k-NN of (qx,qy,qz) in frame g.  All arguments are warp-uniform.
// Returns false if the query has to be re-run on a coarser level (only when !last_level).
template <class TopK>
__device__ __forceinline__ bool warp_knn_search(TopK &tk, const GridDesc &g, const uint32_t *__restrict__ cell_start,
                                                const float4 *__restrict__ pts, float qx, float qy, float qz,
                                                int max_rings, bool last_level, unsigned long long seed = PCR_EMPTY_KEY,
                                                WarpSelScratch *sel = nullptr) {
    tk.reset(PCR_EMPTY_KEY);
    if (seed != PCR_EMPTY_KEY) tk.insert(seed);  // a known candidate (ICP: last iteration's neighbour) bounds the search from the start
    const uint32_t m = g.pt_end - g.pt_begin;
    if (m == 0) return true;
    if (m <= kBruteFrame || m <= (uint32_t)tk.kk) {
        scan_run(tk, pts, g.pt_begin, g.pt_end, qx, qy, qz);
        return true;
    }
    const int lane = tk.lane;
    double f0, f1, f2;
    const int c0 = cell_coord(g, 0, pick_axis(g.ax[0], qx, qy, qz), &f0);
    const int c1 = cell_coord(g, 1, pick_axis(g.ax[1], qx, qy, qz), &f1);
    const int c2 = cell_coord(g, 2, pick_axis(g.ax[2], qx, qy, qz), &f2);
    const int d0n = g.dims[0], d1n = g.dims[1], d2n = g.dims[2];
    // rows and cells beyond the current k-th best are skipped (exactness: see scan_row_pruned); an isolated
    // outlier a few metres above a dense surface otherwise streams the whole footprint of every shell
    const float ff0 = (float)f0, ff1 = (float)f1, ff2 = (float)f2;
    const float inv_h2 = (float)(g.inv_h * g.inv_h);

    {  // the 3x3x3 cube: 9 rows, one lane each
        uint32_t b = 0, e = 0;
        float gp = 0.f;
        const float tau_c = warp_tau_cells(tk, inv_h2);  // finite only for a seeded search
        if (lane < 9) {
            const int e0 = lane / 3 - 1, e1 = lane % 3 - 1;
            int a0 = c0 + e0, a1 = c1 + e1;
            if (a0 >= 0 && a0 < d0n && a1 >= 0 && a1 < d1n) {
                int z0 = max(c2 - 1, 0), z1 = min(c2 + 1, d2n - 1);
                const float g0 = axis_gap(e0, ff0), g1 = axis_gap(e1, ff1);
                gp = g0 * g0 + g1 * g1;
                if (trim_row(tau_c, gp, c2, ff2, z0, z1)) {
                    uint32_t lin = cell_linear(g, a0, a1, z0);
                    b = __ldg(&cell_start[lin]);
                    e = __ldg(&cell_start[lin + (uint32_t)(z1 - z0) + 1u]);
                }
            }
        }
        bool selected = false;
        if constexpr (std::is_same<TopK, RegTopK>::value) {
            if (sel && seed == PCR_EMPTY_KEY && tk.kk <= kWarpSelMaxK) selected = warp_select_cube(tk, *sel, pts, b, e, qx, qy, qz, (float)(g.h * g.h));
        }
        if (!selected) scan_runs(tk, pts, b, e, 1u << 4, qx, qy, qz, gp, inv_h2);
    }
    for (int R = 1;; R++) {
        // nearest face of the scanned cube that still has cells behind it
        double bound2;
        if (!ring_bound2(g, c0, c1, c2, f0, f1, f2, R, bound2)) return true;  // the whole grid has been scanned
        const unsigned long long kth = tk.kth();
        if (kth != PCR_EMPTY_KEY && (double)key_d2(kth) < bound2 * (1.0 - 1e-6)) return true;
        if (R >= max_rings) {
            if (!last_level) return false;  // deferred to the next (coarser) level
            // coarsest level: rescan the frame, pruned by the k-th best so far
            tk.reset(kth == PCR_EMPTY_KEY ? PCR_EMPTY_KEY : kth + 1ull);
            scan_run(tk, pts, g.pt_begin, g.pt_end, qx, qy, qz);
            return true;
        }
        // shell R+1: rows (i0,i1) of the (2S+1)^2 square; border rows are full runs, interior rows
        // contribute their two end cells (slot 0 = low end, slot 1 = high end)
        const int S = R + 1, side = 2 * S + 1, T = 2 * side * side;
        for (int t0 = 0; t0 < T; t0 += 32) {
            int t = t0 + lane;
            uint32_t b = 0, e = 0;
            float gp = 0.f;
            const float tau_c = warp_tau_cells(tk, inv_h2);
            if (t < T) {
                int slot = t & 1, r = t >> 1;
                int i0 = r / side, i1 = r - i0 * side;
                int e0 = i0 - S, e1 = i1 - S;
                int a0 = c0 + e0, a1 = c1 + e1;
                if (a0 >= 0 && a0 < d0n && a1 >= 0 && a1 < d1n) {
                    bool border = (e0 == S) || (e0 == -S) || (e1 == S) || (e1 == -S);
                    int z0, z1;
                    bool ok;
                    if (border) {
                        ok = slot == 0;
                        z0 = max(c2 - S, 0);
                        z1 = min(c2 + S, d2n - 1);
                    } else {
                        int z = slot == 0 ? c2 - S : c2 + S;
                        ok = z >= 0 && z < d2n;
                        z0 = z1 = z;
                    }
                    if (ok) {
                        const float g0 = axis_gap(e0, ff0), g1 = axis_gap(e1, ff1);
                        gp = g0 * g0 + g1 * g1;
                        if (trim_row(tau_c, gp, c2, ff2, z0, z1)) {
                            uint32_t lin = cell_linear(g, a0, a1, z0);
                            b = __ldg(&cell_start[lin]);
                            e = __ldg(&cell_start[lin + (uint32_t)(z1 - z0) + 1u]);
                        }
                    }
                }
            }
            scan_runs(tk, pts, b, e, 0u, qx, qy, qz, gp, inv_h2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Thread-per-query search (level 0 of the KNN / SOR / normals kernels with k <= 32, and the ICP
// 1-NN): same rule, but every THREAD owns one query and keeps its k best keys in registers.
// Neighbouring threads hold neighbouring queries of the cell-sorted order, so a warp walks (almost)
// the same runs in lock step: loads are warp-broadcast L1 hits, no lane idles on a short run, and
// an insertion costs ~6 instructions per slot instead of a ~30-instruction shuffle round.
// ------------------------------------------------------------------------------------------------

// k best keys, ascending, capacity KC (compile time).  With kk < KC the list simply keeps KC
// entries (the threshold is a little looser than needed); only the first kk are reported.
template <int KC>
struct ThreadTopK {
    unsigned long long K[KC];
    int kk;
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int j = 0; j < KC; j++) K[j] = PCR_EMPTY_KEY;
    }
    __device__ __forceinline__ void offer(unsigned long long x) {
        if (x < K[KC - 1]) {
#pragma unroll
            for (int j = KC - 1; j > 0; --j) {
                bool up = x < K[j - 1];
                K[j] = up ? K[j - 1] : (x < K[j] ? x : K[j]);
            }
            K[0] = x < K[0] ? x : K[0];
        }
    }
    __device__ __forceinline__ unsigned long long kth() const {  // kk-th best, EMPTY if not there yet
        unsigned long long r = K[KC - 1];
#pragma unroll
        for (int j = 0; j < KC - 1; j++) r = (j == kk - 1) ? K[j] : r;
        return r;
    }
    __device__ __forceinline__ int count() const {
        int c = 0;
#pragma unroll
        for (int j = 0; j < KC; j++) c += (K[j] != PCR_EMPTY_KEY && j < kk) ? 1 : 0;
        return c;
    }
};

struct ThreadBest1 {  // k = 1
    unsigned long long best;
    __device__ __forceinline__ void reset() { best = PCR_EMPTY_KEY; }
    __device__ __forceinline__ void offer(unsigned long long x) { best = x < best ? x : best; }
    __device__ __forceinline__ unsigned long long kth() const { return best; }
    __device__ __forceinline__ int count() const { return best != PCR_EMPTY_KEY ? 1 : 0; }
};

// kWide: four loads in flight instead of two.  Measured: the unpruned walk of the k >= 16 kernels gains 5-7 %
// (its runs are long and the kernel waits on these loads), the pruned walk of the short lists loses 20-30 %
// (its trimmed runs are a few points long).
template <bool kWide = false, class Acc>
__device__ __forceinline__ void thread_scan_run(Acc &acc, const float4 *__restrict__ pts, uint32_t b, uint32_t e, float qx,
                                                float qy, float qz) {
    uint32_t i = b;
    if (kWide)
    for (; i + 4 <= e; i += 4) {  // four loads in flight
        float4 p0 = __ldg(&pts[i]), p1 = __ldg(&pts[i + 1]), p2 = __ldg(&pts[i + 2]), p3 = __ldg(&pts[i + 3]);
        acc.offer(make_key(dist2_exact(qx, qy, qz, p0.x, p0.y, p0.z), __float_as_uint(p0.w)));
        acc.offer(make_key(dist2_exact(qx, qy, qz, p1.x, p1.y, p1.z), __float_as_uint(p1.w)));
        acc.offer(make_key(dist2_exact(qx, qy, qz, p2.x, p2.y, p2.z), __float_as_uint(p2.w)));
        acc.offer(make_key(dist2_exact(qx, qy, qz, p3.x, p3.y, p3.z), __float_as_uint(p3.w)));
    }
    for (; i + 2 <= e; i += 2) {  // two loads in flight
        float4 p0 = __ldg(&pts[i]), p1 = __ldg(&pts[i + 1]);
        acc.offer(make_key(dist2_exact(qx, qy, qz, p0.x, p0.y, p0.z), __float_as_uint(p0.w)));
        acc.offer(make_key(dist2_exact(qx, qy, qz, p1.x, p1.y, p1.z), __float_as_uint(p1.w)));
    }
    if (i < e) {
        float4 p = __ldg(&pts[i]);
        acc.offer(make_key(dist2_exact(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w)));
    }
}

// Returns false if the query is deferred to the next (coarser) grid level.  `kk` = neighbours wanted
// (for the early-deferral rule): a query that has found less than kk/4 candidates in its 27 cells,
// or is still short of kk after two shells, sits in (near-)empty space -- walking more shells of
// this fine grid is the expensive way to find its neighbours.
// The run scan has ONE call site (the top-k insertion is unrolled and large): shells and the
// whole-frame pass are both expressed as "a list of runs".
template <class Acc>
__device__ __forceinline__ bool thread_grid_search(Acc &acc, const GridDesc &g, const uint32_t *__restrict__ cell_start,
                                                   const float4 *__restrict__ pts, float qx, float qy, float qz, int kk,
                                                   int max_rings, bool last_level) {
    acc.reset();
    const uint32_t m = g.pt_end - g.pt_begin;
    if (m == 0) return true;
    bool whole = m <= kBruteFrame || m <= (uint32_t)kk;  // tiny frame: one run = everything
    double f0 = 0, f1 = 0, f2 = 0;
    int c0 = 0, c1 = 0, c2 = 0;
    if (!whole) {
        c0 = cell_coord(g, 0, pick_axis(g.ax[0], qx, qy, qz), &f0);
        c1 = cell_coord(g, 1, pick_axis(g.ax[1], qx, qy, qz), &f1);
        c2 = cell_coord(g, 2, pick_axis(g.ax[2], qx, qy, qz), &f2);
    }
    const int d0n = g.dims[0], d1n = g.dims[1], d2n = g.dims[2];
    for (int S = 1;; S++) {
        // shell S (S == 1 also takes the centre): rows (e0, e1) of the (2S+1)^2 square; border rows
        // are full runs (slot 0), interior rows contribute their two end cells (slot 0 / 1)
        int e0 = -S, e1 = -S, slot = 0;
        const int T = whole ? 1 : 2 * (2 * S + 1) * (2 * S + 1);
        for (int t = 0; t < T; t++) {
            uint32_t b = 0, e = 0;
            if (whole) {
                b = g.pt_begin;
                e = g.pt_end;
            } else {
                // S == 1: take the query's own row first (the threshold tightens early)
                const int m0 = S == 1 ? (e0 == -1 ? 0 : (e0 == 0 ? -1 : 1)) : e0;
                const int m1 = S == 1 ? (e1 == -1 ? 0 : (e1 == 0 ? -1 : 1)) : e1;
                const int a0 = c0 + m0, a1 = c1 + m1;
                if (a0 >= 0 && a0 < d0n && a1 >= 0 && a1 < d1n) {
                    const bool border = S == 1 || e0 == S || e0 == -S || e1 == S || e1 == -S;
                    int z0, z1;
                    bool ok;
                    if (border) {
                        ok = slot == 0;
                        z0 = max(c2 - S, 0);
                        z1 = min(c2 + S, d2n - 1);
                    } else {
                        const int z = slot == 0 ? c2 - S : c2 + S;
                        ok = z >= 0 && z < d2n;
                        z0 = z1 = z;
                    }
                    if (ok) {
                        const uint32_t lin = cell_linear(g, a0, a1, z0);
                        b = __ldg(&cell_start[lin]);
                        e = __ldg(&cell_start[lin + (uint32_t)(z1 - z0) + 1u]);
                    }
                }
                slot ^= 1;
                if (!slot) {
                    e1++;
                    if (e1 > S) {
                        e1 = -S;
                        e0++;
                    }
                }
            }
            thread_scan_run<true>(acc, pts, b, e, qx, qy, qz);
        }
        if (whole) return true;
        const int R = S;
        double bound2;
        if (!ring_bound2(g, c0, c1, c2, f0, f1, f2, R, bound2)) return true;
        const unsigned long long kth = acc.kth();
        if (kth != PCR_EMPTY_KEY && (double)key_d2(kth) < bound2 * (1.0 - 1e-6)) return true;
        if (!last_level) {
            if (R >= max_rings) return false;
            const int cnt = acc.count();
            if ((R == 1 && cnt * 4 < kk) || (R >= 2 && cnt < kk)) return false;
        } else if (R >= max_rings) {
            acc.reset();  // restart with a whole-frame pass (re-offering a listed key would duplicate it)
            whole = true;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The same search with PRUNED rows.  ncu on the ICP 1-NN kernel (1 M queries, 2 points per cell)
// showed half of its 185 warp instructions per query in the enumeration of cell rows and most of the
// rest in candidates that could not win: once the query's own row has been scanned the best d^2 is
// usually far below the cell size.  Here every row of a shell is first tested against the current
// k-th best tau: with gp = squared in-plane gap (in cells) between the query and the row, the row is
// skipped if gp > tau, and otherwise only the cells within sqrt(tau - gp) of the query along the fast
// axis are read.
// Exactness: the trim keeps every point whose TRUE squared distance is <= tau * (1 + 1e-5) (+ an
// absolute 1e-8 cell^2 for the f32 copies of the fractions), which covers the <= 3 ulp rounding of the
// f32 d^2 the lists are built from; a point outside it can neither enter the list nor tie with its
// last entry.  While the list is not full tau is +inf and nothing is trimmed.  The ring rule and the
// deferral rule are those of thread_grid_search, so both functions return the same lists.
// ------------------------------------------------------------------------------------------------
template <class Acc>
__device__ __forceinline__ float acc_tau(const Acc &acc) {  // k-th best d^2 so far, +inf while the list is short
    const unsigned long long k = acc.kth();
    return k == PCR_EMPTY_KEY ? INFINITY : key_d2(k);
}

template <class Acc>
__device__ __forceinline__ void scan_row_pruned(Acc &acc, const GridDesc &g, const uint32_t *__restrict__ cell_start,
                                                const float4 *__restrict__ pts, int a0, int a1, int zl, int zh, float gp, int c2,
                                                float f2, float inv_h2, float qx, float qy, float qz) {
    if (zl > zh) return;
    const float tau_c = acc_tau(acc) * inv_h2 * (1.0f + 1e-5f) + 1e-8f;  // cell units^2 (+inf: no pruning)
    if (gp > tau_c) return;
    if (tau_c < 1e12f && fabsf(f2) < 1e4f) {
        const float w = sqrtf(tau_c - gp) * (1.0f + 1e-5f) + 1e-5f;
        zl = max(zl, c2 + (int)floorf(f2 - w));
        zh = min(zh, c2 + (int)floorf(f2 + w));
        if (zl > zh) return;
    }
    const uint32_t lin = cell_linear(g, a0, a1, zl);
    thread_scan_run(acc, pts, __ldg(&cell_start[lin]), __ldg(&cell_start[lin + (uint32_t)(zh - zl) + 1u]), qx, qy, qz);
}

template <class Acc>
__device__ __forceinline__ bool thread_grid_search_pruned(Acc &acc, const GridDesc &g, const uint32_t *__restrict__ cell_start,
                                                          const float4 *__restrict__ pts, float qx, float qy, float qz, int kk,
                                                          int max_rings, bool last_level, int seed_mode = 0) {
    // seed_mode != 0: `acc` already holds a real candidate of this index (k = 1 only: ICP's neighbour of
    // the previous iteration) -- it bounds the search from the first row on and cannot change its result.
    //   1: walk a further shell only if that shell is guaranteed to settle the query, else defer
    //   2: walk up to max_rings shells (the deferred pass on the fine grid: the seed keeps the shells cheap)
    const bool seeded = seed_mode != 0;
    if (!seeded) acc.reset();
    const uint32_t m = g.pt_end - g.pt_begin;
    if (m == 0) return true;
    if (m <= kBruteFrame || m <= (uint32_t)kk) {  // tiny frame: one run = everything
        thread_scan_run(acc, pts, g.pt_begin, g.pt_end, qx, qy, qz);
        return true;
    }
    double f0, f1, f2;
    const int c0 = cell_coord(g, 0, pick_axis(g.ax[0], qx, qy, qz), &f0);
    const int c1 = cell_coord(g, 1, pick_axis(g.ax[1], qx, qy, qz), &f1);
    const int c2 = cell_coord(g, 2, pick_axis(g.ax[2], qx, qy, qz), &f2);
    const float ff0 = (float)f0, ff1 = (float)f1, ff2 = (float)f2;
    const int d0n = g.dims[0], d1n = g.dims[1], d2n = g.dims[2];
    const float inv_h2 = (float)(g.inv_h * g.inv_h);
    for (int S = 1;; S++) {
        // shell S (S == 1 also takes the centre, the query's own row first): border rows are full runs,
        // interior rows contribute their two end cells
        const int side = 2 * S + 1, T = side * side;
        const int zlo = max(c2 - S, 0), zhi = min(c2 + S, d2n - 1);
#pragma unroll 1
        for (int r = 0; r < T; r++) {
            int e0 = r / side, e1 = r - e0 * side;
            if (S == 1) {
                e0 = e0 == 0 ? 0 : (e0 == 1 ? -1 : 1);
                e1 = e1 == 0 ? 0 : (e1 == 1 ? -1 : 1);
            } else {
                e0 -= S;
                e1 -= S;
            }
            const int a0 = c0 + e0, a1 = c1 + e1;
            if (a0 < 0 || a0 >= d0n || a1 < 0 || a1 >= d1n) continue;
            const float g0 = axis_gap(e0, ff0), g1 = axis_gap(e1, ff1);
            const float gp = g0 * g0 + g1 * g1;
            if (S == 1 || e0 == S || e0 == -S || e1 == S || e1 == -S) {
                scan_row_pruned(acc, g, cell_start, pts, a0, a1, zlo, zhi, gp, c2, ff2, inv_h2, qx, qy, qz);
            } else {
                if (c2 - S >= 0) scan_row_pruned(acc, g, cell_start, pts, a0, a1, c2 - S, c2 - S, gp, c2, ff2, inv_h2, qx, qy, qz);
                if (c2 + S < d2n) scan_row_pruned(acc, g, cell_start, pts, a0, a1, c2 + S, c2 + S, gp, c2, ff2, inv_h2, qx, qy, qz);
            }
        }
        double bound2;
        if (!ring_bound2(g, c0, c1, c2, f0, f1, f2, S, bound2)) return true;
        const unsigned long long kth = acc.kth();
        if (kth != PCR_EMPTY_KEY && (double)key_d2(kth) < bound2 * (1.0 - 1e-6)) return true;
        if (!last_level) {
            if (S >= max_rings) return false;
            // a seeded query knows how far it has to look: walk one more shell only if that shell settles
            // it ((S + 1) cells >= its current best distance); otherwise the next level is the cheaper
            // place to look (the seed goes along)
            if (seed_mode == 1 && kth != PCR_EMPTY_KEY) {
                const double reach = (double)(S + 1) * g.h;
                if ((double)key_d2(kth) > reach * reach) return false;
            }
            // (the count below can be smaller than in thread_grid_search once rows are trimmed, but
            // rows are only trimmed when the list is already full, i.e. cnt == kk)
            const int cnt = acc.count();
            if ((S == 1 && cnt * 4 < kk) || (S >= 2 && cnt < kk)) return false;
        } else if (S >= max_rings) {
            if (!seeded) acc.reset();  // coarsest level: scan the whole frame instead (min over everything: a kept seed is harmless for k = 1)
            thread_scan_run(acc, pts, g.pt_begin, g.pt_end, qx, qy, qz);
            return true;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Ball scan for a SEEDED 1-NN query (ICP: the neighbour of the previous iteration, r0 away).  The answer
// lies inside the ball of radius r0 around the query, so instead of walking shells and proving a ring bound
// the thread visits the rows (e0, e1) of the square |e| <= ceil(r0 / h) centre-out, skips every row whose
// in-plane gap exceeds the current best and reads only the cells within sqrt(tau - gap^2) along the fast
// axis: O(R^2) row tests instead of the O(R^3) cell visits of a shell walk, and tau shrinks as it goes.
// Returns false (nothing scanned) if the ball spans more than max_cells cells: the caller defers the query.
// Exactness: scan_row_pruned keeps every point with d^2 <= tau (1 + 1e-5); rows and cells outside that bound
// hold nothing that could beat or tie the current best, and every row inside it is visited.
// ------------------------------------------------------------------------------------------------
template <class Acc>
__device__ __forceinline__ bool thread_ball_search(Acc &acc, const GridDesc &g, const uint32_t *__restrict__ cell_start,
                                                   const float4 *__restrict__ pts, float qx, float qy, float qz, int max_cells) {
    const float tau0 = acc_tau(acc);
    const float r_cells = sqrtf(tau0) * (float)g.inv_h * (1.0f + 1e-5f) + 1e-3f;
    if (!(r_cells < (float)max_cells)) return false;  // also catches an empty list (tau = +inf)
    const int R0 = (int)ceilf(r_cells);
    double f0, f1, f2;
    const int c0 = cell_coord(g, 0, pick_axis(g.ax[0], qx, qy, qz), &f0);
    const int c1 = cell_coord(g, 1, pick_axis(g.ax[1], qx, qy, qz), &f1);
    const int c2 = cell_coord(g, 2, pick_axis(g.ax[2], qx, qy, qz), &f2);
    const float ff0 = (float)f0, ff1 = (float)f1, ff2 = (float)f2;
    if (fabsf(ff0) > 1e4f || fabsf(ff1) > 1e4f || fabsf(ff2) > 1e4f) return false;  // far outside the grid: not a ball-scan case
    const int d0n = g.dims[0], d1n = g.dims[1], d2n = g.dims[2];
    const float inv_h2 = (float)(g.inv_h * g.inv_h);
    for (int r = 0; r <= R0; r++) {
        // every row of square ring r is at least (r - 1) cells away in the plane: stop once that exceeds the best
        if (r >= 2) {
            const float lo = (float)(r - 1);
            if (lo * lo > acc_tau(acc) * inv_h2 * (1.0f + 1e-5f) + 1e-8f) break;
        }
        const int side = 2 * r + 1;
        const int T = r == 0 ? 1 : 8 * r;  // cells on the perimeter of the (2r+1)^2 square
#pragma unroll 1
        for (int t = 0; t < T; t++) {
            int e0, e1;
            if (r == 0) {
                e0 = 0; e1 = 0;
            } else if (t < side) {  // top edge
                e0 = -r; e1 = t - r;
            } else if (t < 2 * side) {  // bottom edge
                e0 = r; e1 = t - side - r;
            } else {  // left / right edges without the corners
                const int u = t - 2 * side;
                e0 = (u >> 1) - r + 1;
                e1 = (u & 1) ? r : -r;
            }
            const int a0 = c0 + e0, a1 = c1 + e1;
            if (a0 < 0 || a0 >= d0n || a1 < 0 || a1 >= d1n) continue;
            const float g0 = axis_gap(e0, ff0), g1 = axis_gap(e1, ff1);
            scan_row_pruned(acc, g, cell_start, pts, a0, a1, 0, d2n - 1, g0 * g0 + g1 * g1, c2, ff2, inv_h2, qx, qy, qz);
        }
    }
    return true;
}


// ------------------------------------------------------------------------------------------------
// Selection instead of insertion (round 2).  ncu on the insertion kernels: half of all warp
// instructions were ThreadTopK::offer's unrolled shift network, executed at 14 of 32 lanes -- whenever ONE
// lane accepts a candidate the whole warp pays ~150 instructions, and with ~60 acceptances per query
// spread over ~150 candidates that is almost every step.  Here a candidate that passes the threshold is
// only WRITTEN into a free slot of the thread's column of a shared-memory buffer (32 slots, one STS.64), and
// the warp tightens the threshold together when any of its lanes runs out of slots:
//   sort      the 32 slots are ranked by a register sorting network (Batcher, 191 exchanges) over ONE u32 per
//             slot: the upper 27 bits of the d^2 bit pattern with the slot number in the low 5 bits, so an
//             exchange is a umin/umax pair and no key ever moves in shared memory;
//   compact   the kk-th ranked slot gives the new threshold tau (rounded up to the end of its 27-bit bucket);
//             slots ranked behind that bucket are marked free again;
//   finalize  (end of a shell) the ranking is EXACT unless two neighbours among the first kk+1 ranks share a
//             27-bit bucket (d^2 equal to 4e-6 relative: ties, lattices, duplicates); then the warp orders its
//             buffers by the full (d^2, index) keys with a plain insertion sort in shared memory (rare, small code).
// All of it runs with the 32 lanes in lock step (the walk is warp-synchronous: row loops run to the warp's
// longest run, lanes without work are predicated off), so no lane idles while another sorts.
// Acceptance is `d2 <= tau` (one FSETP; ties are kept) and, once an exact list is known, `key < tau_key`.  A
// candidate that fails either can neither enter the final list nor tie with its last entry, so the result is
// the list ThreadTopK builds.
// ------------------------------------------------------------------------------------------------
constexpr int kSelCap = 32;      // slots per thread (the sorting network's width)
constexpr int kSelStride = 128;  // threads per block
constexpr int kSelMaxK = 21;     // kk <= kSelCap - 11: at least two groups of candidates fit between compactions
constexpr uint32_t kSelPad = 0xffffffe0u;  // rank key of an empty slot (| slot): behind every real d^2 (bits <= 0x7f800000)
constexpr int kSelBins = 2 * kSelCap;      // the threshold histogram lives in the (still empty) slots: 64 u32 counters

// 32 u64 slots per thread, laid out [warp][slot][lane] (file scope: every access below is a plain LDS / STS): a warp's
// slots are one contiguous 8 KB block that no other warp touches (the warps of a block are not in step with each
// other), and slot j of a warp's 32 lanes is one conflict-free 256 B row.
__shared__ unsigned long long g_sel_buf[kSelCap * kSelStride];

__device__ __forceinline__ unsigned long long &sel_slot(int j) {
    return g_sel_buf[(threadIdx.x & ~31u) * kSelCap + j * 32 + (threadIdx.x & 31u)];
}
__device__ __forceinline__ uint32_t sel_slot_hi(int j) {
    return reinterpret_cast<const uint32_t *>(g_sel_buf)[((threadIdx.x & ~31u) * kSelCap + j * 32 + (threadIdx.x & 31u)) * 2 + 1];
}
// The same block as 64 u32 counters per lane (the threshold histogram), [counter][lane]: conflict-free, one IMAD per address.
__device__ __forceinline__ uint32_t &sel_word(int w) {
    return reinterpret_cast<uint32_t *>(g_sel_buf)[(threadIdx.x & ~31u) * kSelCap * 2 + w * 32 + (threadIdx.x & 31u)];
}

// Exact order for the rare warp whose ranking has a bucket collision: occupied slots are moved to the front and
// insertion-sorted by the full key; the kk smallest stay.  Returns the new free mask (slots >= min(cnt, kk)).
static __device__ __noinline__ uint32_t sel_exact_slow(uint32_t free_mask, int kk) {
    int n = 0;
    for (int j = 0; j < kSelCap; j++)
        if (!((free_mask >> j) & 1u)) {
            const unsigned long long v = sel_slot(j);
            int i = n++;
            for (; i > 0 && sel_slot(i - 1) > v; i--) sel_slot(i) = sel_slot(i - 1);  // (i - 1 < j: already moved)
            sel_slot(i) = v;
        }
    const int m = n < kk ? n : kk;
    return m >= 32 ? 0u : ~((1u << m) - 1u);
}

struct ThreadSel {
    unsigned long long tau_key;  // exact kk-th best key once a finalize has seen kk entries, else EMPTY
    float tau;                   // accept d2 <= tau
    uint32_t free_mask;          // bit j set: slot j is empty
    int kk;
    int m;                       // entries of the final list (valid after finalize)
    uint32_t n_eval;             // distance evaluations of this query (both passes), for the bench's FP32 roofline
    uint32_t ord[kSelCap];       // after finalize(): ord[j] & 31 = slot of the j-th best (static indexing only: registers)

    __device__ __forceinline__ void reset() {
        free_mask = 0xffffffffu;
        tau = INFINITY;
        tau_key = PCR_EMPTY_KEY;
        m = 0;
    }
    __device__ __forceinline__ void offer(float d2, uint32_t idx) {
        if (d2 <= tau) {  // false for NaN (tombstoned points)
            const unsigned long long key = make_key(d2, idx);
            if (key < tau_key) {
                const int slot = __ffs(free_mask) - 1;
                free_mask &= free_mask - 1u;
                sel_slot(slot) = key;
            }
        }
    }
    // one rank key per slot, sorted
    __device__ __forceinline__ void rank_slots(uint32_t (&k)[kSelCap]) const {
#pragma unroll
        for (int j = 0; j < kSelCap; j++) k[j] = ((free_mask >> j) & 1u) ? (kSelPad | j) : ((sel_slot_hi(j) & ~31u) | j);
#define PCR_CE32(i, j)                     \
    {                                      \
        const uint32_t x = k[i], y = k[j]; \
        k[i] = min(x, y);                  \
        k[j] = max(x, y);                  \
    }
        PCR_SORTNET32(PCR_CE32)
#undef PCR_CE32
    }
    // warp-synchronous: call with all 32 lanes, after at most 4 offers
    __device__ __forceinline__ void make_room() {
        if (!__any_sync(PCR_FULL, __popc(free_mask) < 4)) return;
        uint32_t k[kSelCap];
        rank_slots(k);
        const int cnt = kSelCap - __popc(free_mask);
        if (cnt >= kk) {  // fewer than kk: nothing can be dropped yet
            uint32_t kb = k[0];
#pragma unroll
            for (int j = 1; j < kSelCap; j++) kb = (j == kk - 1) ? k[j] : kb;
            const uint32_t tb = kb | 31u;  // end of the kk-th best's bucket: every slot ranked behind it is farther
            tau = fminf(tau, __uint_as_float(min(tb, 0x7f800000u)));  // (the bucket of +inf ends in NaN patterns)
#pragma unroll
            for (int j = 1; j < kSelCap; j++)
                if (j >= kk && k[j] > tb && k[j] < kSelPad) free_mask |= 1u << (k[j] & 31u);
        }
        // a bucket that holds more than 11 candidates (ties, duplicates) keeps a buffer full: the exact list has kk entries
        if (__any_sync(PCR_FULL, __popc(free_mask) < 4)) {
            free_mask = sel_exact_slow(free_mask, kk);
            if (kSelCap - __popc(free_mask) == kk) {
                tau_key = sel_slot(kk - 1);
                tau = key_d2(tau_key);
            }
        }
    }
    // warp-synchronous: the exact list (end of a shell)
    __device__ __forceinline__ void finalize() {
        rank_slots(ord);
        int cnt = kSelCap - __popc(free_mask);
        m = cnt < kk ? cnt : kk;
        bool bad = false;
#pragma unroll
        for (int j = 0; j + 1 < kSelCap; j++)
            if (j < m && j + 1 < cnt && ((ord[j] ^ ord[j + 1]) < 32u)) bad = true;  // two neighbours of the ranking in one bucket
        if (__any_sync(PCR_FULL, bad)) {
            free_mask = sel_exact_slow(free_mask, kk);  // slots 0 .. m-1 now hold the list in order
#pragma unroll
            for (int j = 0; j < kSelCap; j++) ord[j] = j < m ? ((sel_slot_hi(j) & ~31u) | j) : (kSelPad | j);
        } else {
#pragma unroll
            for (int j = 1; j < kSelCap; j++)
                if (j >= m && ord[j] < kSelPad) free_mask |= 1u << (ord[j] & 31u);
        }
        if (m == kk) {
            tau_key = kth();
            tau = key_d2(tau_key);
        }
    }
    // valid after finalize()
    __device__ __forceinline__ int count() const { return m; }
    __device__ __forceinline__ unsigned long long key_at(uint32_t o) const { return sel_slot((int)(o & 31u)); }
    __device__ __forceinline__ unsigned long long kth() const {
        if (m != kk) return PCR_EMPTY_KEY;
        uint32_t o = ord[0];
#pragma unroll
        for (int j = 1; j < kSelCap; j++) o = (j == kk - 1) ? ord[j] : o;
        return key_at(o);
    }

    // ---- threshold from a histogram (before anything is buffered: the slots serve as 64 counters) ----
    __device__ __forceinline__ void hist_clear() {
#pragma unroll
        for (int w = 0; w < kSelBins; w++) sel_word(w) = 0u;
    }
    __device__ __forceinline__ void hist_add(float d2, float scale) {
        // fminf returns its non-NaN operand: a tombstoned point (d2 = NaN) lands in the last bin, which never counts
        const int bin = __float2int_rz(fminf(__fmul_rn(d2, scale), (float)(kSelBins - 1)));
        atomicAdd(&sel_word(bin), 1u);  // this thread's own counter: the atomic only spares the read-modify-write chain
    }
    // smallest bin edge with at least kk candidates below it -> tau (+inf if the kk-th falls into the last, open bin)
    __device__ __forceinline__ void hist_threshold(float bin_width) {
        uint32_t cum = 0;
        int bstar = kSelBins - 1;
#pragma unroll
        for (int w = 0; w < kSelBins - 1; w++) {
            cum += sel_word(w);
            bstar = min(bstar, cum >= (uint32_t)kk ? w : kSelBins - 1);
        }
        // bin(d2) <= bstar  =>  d2 * scale < bstar + 1 (rounded product)  =>  d2 < (bstar + 1) * width * (1 + 1e-5)
        tau = bstar >= kSelBins - 1 ? INFINITY : (float)(bstar + 1) * bin_width * (1.0f + 1e-5f);
    }
};

// One run per lane against that lane's query, all 32 lanes in lock step: the loop runs to the warp's longest
// run, four candidates per step.  kHist: the candidates only feed the threshold histogram.
template <bool kHist>
__device__ __forceinline__ void ws_scan_run(ThreadSel &acc, const float4 *__restrict__ pts, uint32_t b, uint32_t e, float qx, float qy,
                                            float qz, float scale) {
    const uint32_t n = e - b;
    acc.n_eval += n;
    const uint32_t nmax = __reduce_max_sync(PCR_FULL, n);
    for (uint32_t i = 0; i < nmax; i += 4) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            p[u].x = __int_as_float(0x7fc00000);  // a lane past the end of its run sees a NaN point: it fails every test below
            if (i + u < n) p[u] = __ldg(&pts[b + i + u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float d2 = dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z);
            if (kHist) acc.hist_add(d2, scale);
            else acc.offer(d2, __float_as_uint(p[u].w));
        }
        if (!kHist) acc.make_room();
    }
}

// points in the 3 x 3 x 3 cells around a query (what the first shell of every search reads): 9 row lookups
__device__ __forceinline__ uint32_t count_27_cells(const GridDesc *__restrict__ gp, const uint32_t *__restrict__ cell_start, float qx, float qy,
                                                   float qz, uint32_t *own_row = nullptr /* the three cells of the query's own row */) {
    const GridDesc g = *gp;
    const int c0 = cell_coord(g, 0, pick_axis(g.ax[0], qx, qy, qz), nullptr);
    const int c1 = cell_coord(g, 1, pick_axis(g.ax[1], qx, qy, qz), nullptr);
    const int c2 = cell_coord(g, 2, pick_axis(g.ax[2], qx, qy, qz), nullptr);
    const int z0 = max(c2 - 1, 0), z1 = min(c2 + 1, g.dims[2] - 1);
    uint32_t n = 0;
#pragma unroll
    for (int e0 = -1; e0 <= 1; e0++)
#pragma unroll
        for (int e1 = -1; e1 <= 1; e1++) {
            const int a0 = c0 + e0, a1 = c1 + e1;
            if (a0 < 0 || a0 >= g.dims[0] || a1 < 0 || a1 >= g.dims[1]) continue;
            const uint32_t lin = cell_linear(g, a0, a1, z0);
            const uint32_t nr = __ldg(&cell_start[lin + (uint32_t)(z1 - z0) + 1u]) - __ldg(&cell_start[lin]);
            n += nr;
            if (own_row && e0 == 0 && e1 == 0) *own_row = nr;
        }
    return n;
}

enum SelOutcome { kSelDone = 0, kSelDefer = 1, kSelContinue = 2 };

// thread_grid_search / thread_grid_search_pruned as a warp-synchronous program (same rows, same ring rule, same
// deferral rule, same pruning bound -- see the comments there).  Call with all 32 lanes of a warp; `live` = this
// lane has a query.  first_shells > 0: stop after that many shells; a lane whose list the ring rule cannot confirm
// by then returns kSelContinue and the caller queues it for the follow-up pass (few lanes need more, and a warp
// that walks another shell for one of them idles the other 31).
__device__ __forceinline__ int ws_grid_search(ThreadSel &acc, bool live, const GridDesc *__restrict__ gp, const uint32_t *__restrict__ cell_start,
                                              const float4 *__restrict__ pts, float qx, float qy, float qz, int max_rings, bool last_level,
                                              int first_shells) {
    // (the descriptor stays in memory: the walk keeps only the cell coordinates, the f32 fractions and the table
    // shape in registers, and the shell-end test reloads what it needs -- the sorting network wants the registers)
    acc.reset();
    acc.n_eval = 0;
    const int kk = acc.kk;
    const uint32_t pt_begin = gp->pt_begin, pt_end = gp->pt_end;
    const uint32_t m = live ? pt_end - pt_begin : 0u;
    bool done = m == 0;
    int outcome = kSelDone;
    bool whole = !done && (m <= kBruteFrame || m <= (uint32_t)kk);  // tiny frame: one run = everything
    int c0 = 0, c1 = 0, c2 = 0;
    float ff0 = 0.f, ff1 = 0.f, ff2 = 0.f;
    if (!done && !whole) {
        const GridDesc g = *gp;
        double f0, f1, f2;
        c0 = cell_coord(g, 0, pick_axis(g.ax[0], qx, qy, qz), &f0);
        c1 = cell_coord(g, 1, pick_axis(g.ax[1], qx, qy, qz), &f1);
        c2 = cell_coord(g, 2, pick_axis(g.ax[2], qx, qy, qz), &f2);
        ff0 = (float)f0;
        ff1 = (float)f1;
        ff2 = (float)f2;
    }
    const int d0n = gp->dims[0], d1n = gp->dims[1], d2n = gp->dims[2];
    const uint32_t cell_base = gp->cell_base;
    const float inv_h2 = (float)(gp->inv_h * gp->inv_h);

    // this lane's run of row (e0, e1) of shell S (slot: the two end cells of an interior row), trimmed to the cells
    // within the current threshold (bound and slack of scan_row_pruned; tau = +inf trims nothing)
    auto row_run = [&](int S, int e0, int e1, bool border, int slot, uint32_t &b, uint32_t &e) {
        b = 0;
        e = 0;
        const int a0 = c0 + e0, a1 = c1 + e1;
        if (a0 < 0 || a0 >= d0n || a1 < 0 || a1 >= d1n) return;
        int z0, z1;
        if (border) {
            z0 = max(c2 - S, 0);
            z1 = min(c2 + S, d2n - 1);
        } else {
            z0 = z1 = slot == 0 ? c2 - S : (slot == 1 ? c2 + S : c2);  // (slot 2: the row's middle cell)
            if (z0 < 0 || z0 >= d2n) return;
        }
        const float g0 = axis_gap(e0, ff0), g1 = axis_gap(e1, ff1);
        const float tau_c = acc.tau * inv_h2 * (1.0f + 1e-5f) + 1e-8f;
        if (!trim_row(tau_c, g0 * g0 + g1 * g1, c2, ff2, z0, z1)) return;
        const uint32_t lin = cell_base + ((uint32_t)a0 * (uint32_t)d1n + (uint32_t)a1) * (uint32_t)d2n + (uint32_t)z0;
        b = __ldg(&cell_start[lin]);
        e = __ldg(&cell_start[lin + (uint32_t)(z1 - z0) + 1u]);
    };
    // row order of the 3x3x3 cube: the query's own row first, then its four face neighbours
    auto cube_row = [](int r, int &e0, int &e1) {
        e0 = r / 3;
        e1 = r - e0 * 3;
        e0 = e0 == 0 ? 0 : (e0 == 1 ? -1 : 1);
        e1 = e1 == 0 ? 0 : (e1 == 1 ? -1 : 1);
    };

    // ---- threshold for the 27 cells from a histogram of the d^2 ------------------------------------------------
    // Without it the first 32 candidates all enter the buffer and the threshold only tightens compaction by
    // compaction (ncu: 4-7 per warp, half of the kernel's instructions; 15 and more for a query inside a dense
    // object, whose warp then runs five times longer than the others and IS the kernel's duration).  One extra pass
    // that only counts candidates into 64 linear d^2 bins gives a threshold with kk (+ the rest of one bin)
    // candidates below it, so the collecting pass buffers ~kk keys and rarely compacts.
    //   * up to kSelHistMaxN candidates in the 27 cells (surfaces): all 9 rows are counted; the bins span four times
    //     the kk-th d^2 of a uniform sheet with that many points on the 3 x 3 cells;
    //   * more (a volume: the 27 cells hold 20 x the candidates the list needs): only the query's cell and its six
    //     face neighbours are counted -- the kk-th best of those 7 cells bounds the kk-th best of all, and with it the
    //     collecting pass trims the edge and corner cells it would otherwise read; the bins span three times the
    //     kk-th d^2 of a uniform volume with that many points in 7 cells.
    // A kk-th beyond the last bin leaves tau = +inf, more than 11 candidates in its bin just mean compactions as
    // before: the histogram only speeds things up, the result does not depend on it.
    {
        uint32_t n27 = 0, n7 = 0;
        const bool walk = !done && !whole;
        for (int r = 0; r < 9; r++) {
            int e0, e1;
            cube_row(r, e0, e1);
            uint32_t b = 0, e = 0;
            if (walk) row_run(1, e0, e1, true, 0, b, e);
            n27 += e - b;
            if (e0 == 0 || e1 == 0) {  // own row: all three cells; a face row: its middle cell
                if (e0 != 0 || e1 != 0) {
                    b = e = 0;
                    if (walk) row_run(1, e0, e1, false, 2, b, e);
                }
                n7 += e - b;
            }
        }
        const bool volume = n27 > kSelHistMaxN;
        const uint32_t nh = volume ? n7 : n27;
        const bool hist = walk && nh > (uint32_t)kSelCap - 4u;
        if (__any_sync(PCR_FULL, hist)) {
            const float h2 = 1.0f / inv_h2;
            // sheet: n points on 9 h^2 -> r^2 = 9 h^2 kk / (pi n);  volume: n points in 7 h^3 -> r^2 = h^2 (21 kk / (4 pi n))^(2/3)
            const float q = (float)kk / (float)(nh ? nh : 1u);
            const float r2 = volume ? h2 * cbrtf(1.6711f * q) * cbrtf(1.6711f * q) : 2.8648f * h2 * q;
            const float width = hist ? (volume ? 3.0f : 4.0f) * r2 / (float)kSelBins : 1.0f;
            const float scale = 1.0f / width;
            acc.hist_clear();
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                int e0, e1;
                cube_row(r, e0, e1);
                uint32_t b = 0, e = 0;
                if (hist && !(volume && e0 != 0 && e1 != 0)) row_run(1, e0, e1, !(volume && (e0 != 0 || e1 != 0)), 2, b, e);
                ws_scan_run<true>(acc, pts, b, e, qx, qy, qz, scale);
            }
            if (hist) acc.hist_threshold(width);
        }
    }

    for (int S = 1;; S++) {
        const int side = 2 * S + 1, T = side * side;
#pragma unroll 1
        for (int r = 0; r < T; r++) {
            int e0, e1;
            if (S == 1) {
                cube_row(r, e0, e1);
            } else {
                e0 = r / side;
                e1 = r - e0 * side - S;
                e0 -= S;
            }
            const bool border = S == 1 || e0 == S || e0 == -S || e1 == S || e1 == -S;  // warp-uniform
            // border rows are full runs; interior rows contribute their two end cells (slot 0 / 1)
#pragma unroll 1
            for (int slot = 0; slot < (border ? 1 : 2); slot++) {
                uint32_t b = 0, e = 0;
                if (!done) {
                    if (whole) {
                        if (r == 0 && slot == 0) {
                            b = pt_begin;
                            e = pt_end;
                        }
                    } else {
                        row_run(S, e0, e1, border, slot, b, e);
                    }
                }
                ws_scan_run<false>(acc, pts, b, e, qx, qy, qz, 0.f);
            }
        }
        acc.finalize();
        if (!done) {
            if (whole) {
                done = true;
            } else {
                const GridDesc g = *gp;
                double f0, f1, f2, bound2;
                cell_coord(g, 0, pick_axis(g.ax[0], qx, qy, qz), &f0);
                cell_coord(g, 1, pick_axis(g.ax[1], qx, qy, qz), &f1);
                cell_coord(g, 2, pick_axis(g.ax[2], qx, qy, qz), &f2);
                const unsigned long long kth = acc.kth();
                if (!ring_bound2(g, c0, c1, c2, f0, f1, f2, S, bound2)) {
                    done = true;  // the whole grid has been scanned
                } else if (kth != PCR_EMPTY_KEY && (double)key_d2(kth) < bound2 * (1.0 - 1e-6)) {
                    done = true;
                } else if (!last_level) {
                    const int cnt = acc.count();
                    if (S >= max_rings || (S == 1 && cnt * 4 < kk) || (S >= 2 && cnt < kk)) {
                        done = true;
                        outcome = kSelDefer;
                    } else if (first_shells && S >= first_shells) {
                        done = true;
                        outcome = kSelContinue;
                    }
                } else if (S >= max_rings) {
                    acc.reset();  // coarsest level: restart with a whole-frame pass
                    whole = true;
                }
            }
        }
        if (__all_sync(PCR_FULL, done)) break;
    }
    return outcome;
}

}  // namespace pcr
