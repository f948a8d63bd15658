// knn_tile.cuh -- level 0 of the grid KNN as a cell-tile program (round 2).
//
// Replaces spatial::KdTree::knn (/root/reference/crates/spatial/src/kdtree.rs:64-96) for the self-queries of SOR and
// normal estimation: same result as every other search in this library (the k smallest (d^2, index) keys), different
// schedule.
//
// A warp owns 32 queries that are neighbours in the cell-sorted order.  Queries of one grid row (same frame, same two
// slow cell coordinates) form a SUB-TILE: the cells their 27-cell cubes touch are 9 contiguous runs of the cell-sorted
// point array -- rows (a0 + e0, a1 + e1), cells [zlo - 1, zhi + 1] -- and those 9 runs are staged in shared memory by
// nine `cp.async.bulk` copies (one per lane 0..8, completion counted by one mbarrier per warp; SASS: UBLKCP).  Every
// query then finds its own 3-cell window inside each staged row (two reads of the cell table per row) and works on
// shared memory only:
//   pass 1   d^2 of every candidate of its 27 cells -> one of 32 linear bins (per-thread counters in shared memory); the
//            first bin edge with at least k candidates below it is the acceptance threshold;
//   pass 2   the candidates up to that edge are recorded as ONE u32 each: the upper 23 bits of the d^2 pattern and the
//            candidate's 9-bit position in the staged tile -- key and payload of the ranking in one register;
//   rank     Batcher's 32-input network on those u32 (191 min / max pairs, all in registers); two neighbours of the
//            ranking that share a 23-bit bucket (d^2 equal to 6e-5 relative) are put in (d^2, index) order by
//            recomputing their exact keys from the staged points (a short bubble pass in shared memory, rare);
//   shell 2  a query whose k-th neighbour is not yet inside the scanned cube's guaranteed radius (or that found fewer
//            than k) reads the second shell from global memory, row by row, trimmed to the k-th distance; what it accepts
//            is appended to the staged tile and ranked again.
// What does not fit -- more than 512 candidates around one cell (dense objects), more than 32 candidates below the
// threshold (exact ties), a third shell -- is handed to the older kernels through lists: `ovf_list` (thread per query,
// knn_sel_kernel) and `cont_list` / `defer_list` (warp per query, same level / coarser level).  Every path returns the
// same exact list; the tile only decides where a query runs.
#pragma once

#include "knn_search.cuh"

namespace pcr {

constexpr int kTileWarps = 1;                   // one warp per block: a warp is as slow as its heaviest query, a block as its slowest warp
constexpr uint32_t kTileDenseN = 128;           // more candidates than this in a query's 27 cells: not a tile's job (classify_kernel)
constexpr int kTileThreads = 32 * kTileWarps;
constexpr int kTileCap = 512;                   // staged candidates per warp
constexpr uint32_t kTilePos = 511;              // low 9 bits of a rank key: position in the staged tile
static_assert(kTileCap <= 512, "a staged position has 9 bits");
constexpr int kTileSlots = 32;                  // accepted candidates per query (the sorting network's width)
constexpr int kTileBins = 32;                   // threshold histogram (lives in the slot words before anything is accepted)
constexpr int kTileMaxK = 24;                   // k <= this: at least 8 slots of margin behind the k-th
constexpr int kTileRowSpan = 8;                 // rows (of one slow-axis layer) a sub-tile may span: 3 * (8 + 2) staged runs, one per lane
constexpr int kTileZReach = 6;                  // ... among the queries within this many cells of the first one along the fast axis
constexpr uint32_t kTilePad = 0xfffffe00u;      // rank key of an empty slot: behind every real d^2 (bits <= 0x7f800000)

struct __align__(16) TileWarp {
    float4 stage[kTileCap];          // the sub-tile's candidate points, staged row after row
    uint32_t slot[kTileSlots][32];   // [slot][lane]: histogram counters, then rank keys of the accepted candidates
    uint32_t seg[9][32];             // [row][lane]: this lane's window in staged row r, begin | end << 16
    int4 row[3 * (kTileRowSpan + 2)];  // {cell-table offset of the row's cell 0, first point of the staged run, its position in stage, length}
    unsigned long long bar;          // mbarrier the bulk copies complete on
    uint32_t pad_[2];
};
static_assert(sizeof(TileWarp) % 16 == 0, "TileWarp is an array element");

// ---- PTX: mbarrier + bulk copy global -> shared (SASS: SYNCS.*, UBLKCP) ------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// ---- passes 1 + 2 over the lane's 27 cells (9 windows of the staged rows) -----------------------------------------
// Returns the number of accepted candidates; more than kTileSlots = overflow (only the first kTileSlots were recorded),
// -1 = the threshold cannot be trusted (rounding at a refined bin edge): the caller hands the query back.
__device__ __forceinline__ int tile_collect(TileWarp &S, int lane, int kk, uint32_t N, float qx, float qy, float qz, float h2, uint32_t &n_eval) {
    float tau = INFINITY;
    if (N > (uint32_t)kTileSlots) {
        // The histogram covers [lo, lo + 32 width).  First guess: four times the k-th d^2 of a uniform sheet (three times
        // that of a uniform volume) with N points in the cube -- the model of ws_grid_search.  A k-th beyond the last bin
        // widens the range eightfold; a k-th whose bin holds too many candidates for the slots makes that bin the next
        // range (nb = candidates known to lie below it).  Three passes at most; whatever is left is handed back.
        const float q = (float)kk / (float)N;
        const bool volume = N > kSelHistMaxN;
        const float c = cbrtf(6.4456f * q);
        const float r2 = volume ? h2 * c * c : 2.8648f * h2 * q;
        float width = (volume ? 3.0f : 4.0f) * r2 * (1.0f / (float)kTileBins);
        float lo = 0.0f;
        uint32_t nb = 0;
#pragma unroll 1
        for (int pass = 0; pass < 3; pass++) {
            const float scale = 1.0f / width;
#pragma unroll
            for (int b = 0; b < kTileBins; b++) S.slot[b][lane] = 0u;
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                const uint32_t sg = S.seg[r][lane];
                const uint32_t e = sg >> 16;
                for (uint32_t i = sg & 0xffffu; i < e; i += 4) {  // four candidates in flight (the loads are the latency)
                    float4 p[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) p[u] = S.stage[min(i + u, e - 1u)];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const float d2 = dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z);
                        // fminf returns its non-NaN operand: a tombstoned point (d2 = NaN) lands in the last bin, which never counts
                        const int bin = __float2int_rz(fminf(__fmul_rn(__fsub_rn(d2, lo), scale), (float)(kTileBins - 1)));
                        // (this thread's own counter: the atomic only spares the read-modify-write chain)
                        if (i + u < e && !(d2 < lo)) atomicAdd(&S.slot[bin][lane], 1u);
                    }
                }
            }
            n_eval += N;
            uint32_t cum = nb, below = nb, upto = nb;
            int bstar = kTileBins - 1;
#pragma unroll
            for (int b = 0; b < kTileBins - 1; b++) {
                const uint32_t cb = S.slot[b][lane];
                const bool first = bstar == kTileBins - 1 && cum + cb >= (uint32_t)kk;
                if (first) {
                    bstar = b;
                    below = cum;
                    upto = cum + cb;
                }
                cum += cb;
            }
            if (bstar == kTileBins - 1) {  // the k-th lies beyond the bins
                width *= 8.0f;
                continue;
            }
            // bin(d2) <= bstar  =>  (d2 - lo) * scale < bstar + 1 (rounded)  =>  d2 < (lo + (bstar + 1) * width) * (1 + 1e-5)
            tau = (lo + (float)(bstar + 1) * width) * (1.0f + 1e-5f);
            if (upto <= (uint32_t)kTileSlots) break;
            if (pass < 2) {  // too many up to that edge: look inside the k-th's bin
                lo = lo + (float)bstar * width;
                nb = below;
                width *= 1.0f / (float)kTileBins;
            }
        }
    }
    int cnt = 0;
#pragma unroll 1
    for (int r = 0; r < 9; r++) {
        const uint32_t sg = S.seg[r][lane];
        const uint32_t e = sg >> 16;
        for (uint32_t i = sg & 0xffffu; i < e; i += 4) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; u++) p[u] = S.stage[min(i + u, e - 1u)];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float d2 = dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z);
                const bool acc = i + u < e && d2 <= tau;  // false for NaN (tombstoned points)
                // (one predicated store, no branch: a 33rd candidate lands on the 32nd's slot, and the count says overflow)
                uint32_t *dst = &S.slot[min(cnt, kTileSlots - 1)][lane];
                if (acc) *dst = (__float_as_uint(d2) & ~kTilePos) | (i + u);
                cnt += acc ? 1 : 0;
            }
        }
    }
    n_eval += N;
    // a finite threshold promised at least kk candidates below it; fewer means a candidate fell between two passes' bin edges
    if (cnt < kk && tau != INFINITY) return -1;
    return cnt;
}

// exact (d^2, index) key of the staged candidate behind a rank key
__device__ __forceinline__ unsigned long long tile_full_key(const TileWarp &S, uint32_t rk, float qx, float qy, float qz) {
    const float4 p = S.stage[rk & kTilePos];
    return make_key(dist2_exact(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w));
}

// Ranks slots [0, cnt) of this lane: k[] ascending by (23-bit d^2, position); pads behind.
__device__ __forceinline__ void tile_rank(const TileWarp &S, int lane, int cnt, uint32_t (&k)[kTileSlots]) {
#pragma unroll
    for (int j = 0; j < kTileSlots; j++) k[j] = j < cnt ? S.slot[j][lane] : (kTilePad | (uint32_t)j);
#define PCR_CE32(i, j)                     \
    {                                      \
        const uint32_t x = k[i], y = k[j]; \
        k[i] = min(x, y);                  \
        k[j] = max(x, y);                  \
    }
    PCR_SORTNET32(PCR_CE32)
#undef PCR_CE32
}

// The ranking is exact unless two neighbours of it share a 23-bit bucket within the first m + 1 ranks: then the slots are
// rewritten in rank order and bubble passes over the equal-bucket pairs (exact keys recomputed from the staged points)
// finish the order.  Per lane, divergent, rare.
__device__ __forceinline__ void tile_fix(TileWarp &S, int lane, int cnt, int m, uint32_t (&k)[kTileSlots], float qx, float qy, float qz) {
    bool bad = false;
#pragma unroll
    for (int j = 0; j < kTileMaxK; j++)
        if (j < m && j + 1 < cnt && ((k[j] ^ k[j + 1]) <= kTilePos)) bad = true;
    if (!bad) return;
#pragma unroll
    for (int j = 0; j < kTileSlots; j++)
        if (j < cnt) S.slot[j][lane] = k[j];
    bool swapped = true;
    while (swapped) {
        swapped = false;
        for (int j = 0; j + 1 < cnt; j++) {
            const uint32_t x = S.slot[j][lane], y = S.slot[j + 1][lane];
            if ((x ^ y) <= kTilePos && tile_full_key(S, y, qx, qy, qz) < tile_full_key(S, x, qx, qy, qz)) {
                S.slot[j][lane] = y;
                S.slot[j + 1][lane] = x;
                swapped = true;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kTileSlots; j++)
        if (j < cnt) k[j] = S.slot[j][lane];
}

}  // namespace pcr
