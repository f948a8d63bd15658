// grid_build.cu -- uniform-grid spatial index build (replaces KdTree::build,
// crates/spatial/src/kdtree.rs:25-44).
//
//   K0  bbox_kernel      finite-point bounding box + count per frame          (reads 12 B/pt)
//   K1  count_kernel     cell id per point, per-cell counts (L2 atomics), AoS copy
//                        (reads 12 B/pt, writes 16 B orig4 + 8 B cell id / rank)
//   K1p stats_kernel     occupancy statistics used to choose the cell size     (probe only)
//   K2  scan (2 kernels) exclusive prefix of the dense cell table              (8 B/cell)
//   K3  scatter_kernel   single-pass counting (radix) sort by dense cell id -> cell-sorted float4
//                        (reads 24 B/pt, writes 16 B/pt)
//
// The cell id is a dense linear key with the longest axis fastest, so a row of neighbouring cells
// is ONE contiguous run of points; the order of points inside a cell is irrelevant because every
// consumer ranks candidates by (d^2, original index).
#include "pcr_internal.cuh"

#include <algorithm>
#include <cstring>
#include <chrono>
#include <cmath>

namespace pcr {

namespace {

constexpr int kBuildThreads = 256;
constexpr int kItems = 4;  // points per thread in the streaming kernels

using FrameStats = CloudStats;  // mn / mx: order-preserving uint encoding of f32

struct ProbeStats {
    double sum_log_own, sum_log_super;
};

static inline float ordered_f32(unsigned u) {
    unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &b, 4);
    return f;
}

// ---- K0: bounding box of the finite (and, with a mask, kept) points of each frame -------------
__global__ void __launch_bounds__(kBuildThreads) bbox_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                             const float *__restrict__ z,
                                                             const uint32_t *__restrict__ frame_off, size_t n,
                                                             const uint8_t *__restrict__ mask,
                                                             FrameStats *__restrict__ stats) {
    const int f = blockIdx.y;
    const uint32_t b = frame_off ? frame_off[f] : 0u;
    const uint32_t e = frame_off ? frame_off[f + 1] : (uint32_t)n;
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u}, cnt = 0;
    for (uint32_t i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x) {
        float px = x[i], py = y[i], pz = z[i];
        if (finite3(px, py, pz) && (!mask || mask[i])) {
            unsigned ux = f32_ordered(px), uy = f32_ordered(py), uz = f32_ordered(pz);
            mn[0] = min(mn[0], ux); mx[0] = max(mx[0], ux);
            mn[1] = min(mn[1], uy); mx[1] = max(mx[1], uy);
            mn[2] = min(mn[2], uz); mx[2] = max(mx[2], uz);
            cnt++;
        }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
        mn[a] = __reduce_min_sync(PCR_FULL, mn[a]);
        mx[a] = __reduce_max_sync(PCR_FULL, mx[a]);
    }
    cnt = __reduce_add_sync(PCR_FULL, cnt);
    __shared__ unsigned s_mn[3], s_mx[3], s_cnt;
    if (threadIdx.x == 0) {
        s_mn[0] = s_mn[1] = s_mn[2] = 0xffffffffu;
        s_mx[0] = s_mx[1] = s_mx[2] = 0u;
        s_cnt = 0;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && cnt) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            atomicMin(&s_mn[a], mn[a]);
            atomicMax(&s_mx[a], mx[a]);
        }
        atomicAdd(&s_cnt, cnt);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
            atomicMin(&stats[f].mn[a], s_mn[a]);
            atomicMax(&stats[f].mx[a], s_mx[a]);
        }
        atomicAdd(&stats[f].count, s_cnt);
    }
}

__global__ void init_stats_kernel(FrameStats *stats, int n_frames) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n_frames) {
        stats[f].mn[0] = stats[f].mn[1] = stats[f].mn[2] = 0xffffffffu;
        stats[f].mx[0] = stats[f].mx[1] = stats[f].mx[2] = 0u;
        stats[f].count = 0;
        stats[f].valid = 0;
    }
}

// ---- K1: cell id + per-cell count + rank inside the cell + AoS copy -----------------------------
// rank = old value of the L2 atomic: the later scatter is then a plain store (no second atomic).
__global__ void __launch_bounds__(kBuildThreads) count_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                              const float *__restrict__ z,
                                                              const uint32_t *__restrict__ frame_off, size_t n,
                                                              const uint8_t *__restrict__ mask,
                                                              const GridDesc *__restrict__ grids,
                                                              uint32_t *__restrict__ cell_count,
                                                              uint32_t *__restrict__ cell_id, uint32_t *__restrict__ rank,
                                                              float4 *__restrict__ orig4) {
    PCR_GRID_DEP_SYNC();
    const int f = blockIdx.y;
    const uint32_t b = frame_off ? frame_off[f] : 0u;
    const uint32_t e = frame_off ? frame_off[f + 1] : (uint32_t)n;
    const GridDesc g = grids[f];
    const uint32_t base = b + blockIdx.x * (kBuildThreads * kItems) + threadIdx.x;
#pragma unroll
    for (int it = 0; it < kItems; it++) {
        uint32_t i = base + it * kBuildThreads;
        if (i >= e) break;
        float px = x[i], py = y[i], pz = z[i];
        if (orig4) orig4[i] = make_float4(px, py, pz, 0.f);
        uint32_t cid = 0xffffffffu, r = 0;
        if (finite3(px, py, pz) && (!mask || mask[i])) {
            int c0 = cell_coord(g, 0, pick_axis(g.ax[0], px, py, pz), nullptr);
            int c1 = cell_coord(g, 1, pick_axis(g.ax[1], px, py, pz), nullptr);
            int c2 = cell_coord(g, 2, pick_axis(g.ax[2], px, py, pz), nullptr);
            cid = cell_linear(g, c0, c1, c2);
            r = atomicAdd(&cell_count[cid], 1u);
        }
        cell_id[i] = cid;
        rank[i] = r;
    }
}

// same, for a coarser level of an existing index: reads the finer level's CELL-SORTED points, so
// neighbouring threads fall into the same coarse cell and one warp-aggregated atomic per distinct
// cell replaces up to 32 same-address atomics.  Tombstoned (NaN) entries are skipped.
__global__ void __launch_bounds__(kBuildThreads) count_sorted_kernel(const float4 *__restrict__ fine_sorted, uint32_t n_sorted,
                                                                     const GridDesc *__restrict__ grids, int n_frames,
                                                                     uint32_t *__restrict__ cell_count,
                                                                     uint32_t *__restrict__ cell_id, uint32_t *__restrict__ rank) {
    PCR_GRID_DEP_SYNC();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t cid = 0xffffffffu;
    if (i < n_sorted) {
        const float4 p = fine_sorted[i];
        if (p.x == p.x) {
            int f = 0;
            if (n_frames > 1) {  // frames own contiguous slices of the sorted array
                int lo = 0, hi = n_frames - 1;
                while (lo < hi) {
                    int mid = (lo + hi + 1) >> 1;
                    if (grids[mid].pt_begin <= i) lo = mid;
                    else hi = mid - 1;
                }
                while (lo + 1 < n_frames && grids[lo].pt_end <= i) lo++;
                f = lo;
            }
            const GridDesc g = grids[f];
            int c0 = cell_coord(g, 0, pick_axis(g.ax[0], p.x, p.y, p.z), nullptr);
            int c1 = cell_coord(g, 1, pick_axis(g.ax[1], p.x, p.y, p.z), nullptr);
            int c2 = cell_coord(g, 2, pick_axis(g.ax[2], p.x, p.y, p.z), nullptr);
            cid = cell_linear(g, c0, c1, c2);
        }
    }
    // warp-aggregated atomic: one leader per distinct cell id
    const unsigned peers = __match_any_sync(PCR_FULL, cid);
    uint32_t r = 0;
    if (cid != 0xffffffffu) {
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&cell_count[cid], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        r = base + __popc(peers & ((1u << lane) - 1u));
    }
    if (i < n_sorted) {
        cell_id[i] = cid;
        rank[i] = r;
    }
}

__global__ void __launch_bounds__(kBuildThreads) scatter_sorted_kernel(const float4 *__restrict__ fine_sorted, uint32_t n_sorted,
                                                                       const uint32_t *__restrict__ cell_start,
                                                                       const uint32_t *__restrict__ cell_id,
                                                                       const uint32_t *__restrict__ rank,
                                                                       float4 *__restrict__ sorted) {
    PCR_GRID_DEP_SYNC();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sorted) return;
    uint32_t cid = cell_id[i];
    if (cid == 0xffffffffu) return;
    sorted[cell_start[cid] + rank[i]] = fine_sorted[i];
}

// ---- K1p: occupancy statistics of a probe grid ---------------------------------------------------
// For every indexed point: log2(points in its cell) and log2(points in its 2x2x2 super-cell).  The
// two geometric means give the local occupancy m(h) and its scaling exponent D (m ~ h^D), from
// which the host solves m(h*) = target for the final cell size.
__global__ void __launch_bounds__(kBuildThreads) stats_kernel(const uint32_t *__restrict__ frame_off, size_t n,
                                                              const GridDesc *__restrict__ grids,
                                                              const uint32_t *__restrict__ cell_count,
                                                              const uint32_t *__restrict__ cell_id,
                                                              ProbeStats *__restrict__ out) {
    const int f = blockIdx.y;
    const uint32_t b = frame_off ? frame_off[f] : 0u;
    const uint32_t e = frame_off ? frame_off[f + 1] : (uint32_t)n;
    const GridDesc g = grids[f];
    float s_own = 0.f, s_sup = 0.f;
    for (uint32_t i = b + blockIdx.x * blockDim.x + threadIdx.x; i < e; i += gridDim.x * blockDim.x) {
        uint32_t cid = cell_id[i];
        if (cid == 0xffffffffu) continue;
        uint32_t l = cid - g.cell_base;
        int c2 = l % g.dims[2];
        uint32_t t = l / g.dims[2];
        int c1 = t % g.dims[1];
        int c0 = t / g.dims[1];
        uint32_t own = cell_count[cid], sup = 0;
        int b0 = c0 & ~1, b1 = c1 & ~1, b2 = c2 & ~1;
#pragma unroll
        for (int d0 = 0; d0 < 2; d0++)
#pragma unroll
            for (int d1 = 0; d1 < 2; d1++)
#pragma unroll
                for (int d2 = 0; d2 < 2; d2++) {
                    int a0 = b0 + d0, a1 = b1 + d1, a2 = b2 + d2;
                    if (a0 < g.dims[0] && a1 < g.dims[1] && a2 < g.dims[2]) sup += cell_count[cell_linear(g, a0, a1, a2)];
                }
        s_own += log2f((float)own);
        s_sup += log2f((float)sup);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s_own += __shfl_xor_sync(PCR_FULL, s_own, o);
        s_sup += __shfl_xor_sync(PCR_FULL, s_sup, o);
    }
    __shared__ float sh[2][kBuildThreads / 32];
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sh[0][w] = s_own;
        sh[1][w] = s_sup;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, c = 0;
        for (int i = 0; i < kBuildThreads / 32; i++) {
            a += sh[0][i];
            c += sh[1][i];
        }
        atomicAdd(&out[f].sum_log_own, a);
        atomicAdd(&out[f].sum_log_super, c);
    }
}

// ---- K3: scatter into cell order --------------------------------------------------------------
__global__ void __launch_bounds__(kBuildThreads) scatter_kernel(const float4 *__restrict__ orig4, size_t n,
                                                                const uint32_t *__restrict__ cell_start,
                                                                const uint32_t *__restrict__ cell_id,
                                                                const uint32_t *__restrict__ rank,
                                                                float4 *__restrict__ sorted) {
    PCR_GRID_DEP_SYNC();
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t cid = cell_id[i];
    if (cid == 0xffffffffu) return;
    float4 p = orig4[i];
    p.w = __uint_as_float(i);
    sorted[cell_start[cid] + rank[i]] = p;
}

// ---- tombstones: drop masked-out points from a built level without rebuilding it ------------------
struct TombLevels {
    float4 *sorted[4];
    uint32_t n[4];
    int levels;
};
__global__ void __launch_bounds__(256) tombstone_kernel(TombLevels t, const uint8_t *__restrict__ keep) {
    PCR_GRID_DEP_SYNC();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int l = 0; l < t.levels; l++) {
        if (i >= t.n[l]) continue;
        float4 p = t.sorted[l][i];
        if (!keep[__float_as_uint(p.w)]) {
            const float qnan = __int_as_float(0x7fc00000);
            p.x = p.y = p.z = qnan;  // every distance to it is NaN: its key can never enter a top-k
            t.sorted[l][i] = p;
        }
    }
}

// ---- K2: exclusive scan of a u32 table (two kernels, no inter-block waiting) -------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t *sh) {
    v = __reduce_add_sync(PCR_FULL, v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    if (threadIdx.x < kScanThreads / 32) t = sh[threadIdx.x];
    if (threadIdx.x < 32) t = __reduce_add_sync(PCR_FULL, t);
    if (threadIdx.x == 0) sh[0] = t;
    __syncthreads();
    t = sh[0];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t *__restrict__ data, size_t n,
                                                                      uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t sh[kScanThreads / 32];
    size_t base = (size_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int it = 0; it < kScanItems; it++) {
        size_t i = base + it * kScanThreads + threadIdx.x;
        if (i < n) s += data[i];
    }
    s = block_sum_u32(s, sh);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s;
}

// ---- single-pass exclusive scan (decoupled look-back) ---------------------------------------------
// One launch instead of two, the data read once.  Tiles take a ticket (so a tile only ever waits for tiles that
// are already running), publish their aggregate, then warp 0 walks back over the published words of the tiles
// before it, 32 at a time, until it meets one that already holds an inclusive prefix.
// State word: [63:34] epoch | [33:32] status (1 = tile aggregate, 2 = inclusive prefix) | [31:0] value.  The
// epoch changes with every scan of the context, so the words are never cleared between scans.
constexpr unsigned long long kScanAgg = 1ull << 32, kScanPre = 2ull << 32;

__device__ __forceinline__ uint4 scan_load4(const uint32_t *data, size_t i, size_t n, bool vec) {
    if (vec && i + 3 < n) return *reinterpret_cast<const uint4 *>(data + i);
    uint4 v;
    v.x = i < n ? data[i] : 0u;
    v.y = i + 1 < n ? data[i + 1] : 0u;
    v.z = i + 2 < n ? data[i + 2] : 0u;
    v.w = i + 3 < n ? data[i + 3] : 0u;
    return v;
}

__device__ __forceinline__ void scan_store4(uint32_t *data, size_t i, size_t n, bool vec, uint4 v) {
    if (vec && i + 3 < n) {
        *reinterpret_cast<uint4 *>(data + i) = v;
        return;
    }
    if (i < n) data[i] = v.x;
    if (i + 1 < n) data[i + 1] = v.y;
    if (i + 2 < n) data[i + 2] = v.z;
    if (i + 3 < n) data[i + 3] = v.w;
}

// Warp 0 of a tile: publish the tile's aggregate, sum the published words of the tiles before it (32 at a time, back to the
// first one that already holds an inclusive prefix), publish the tile's own inclusive prefix.  Returns the exclusive prefix.
__device__ __forceinline__ uint32_t scan_lookback_warp0(unsigned long long *state, uint32_t tile, uint32_t epoch, uint32_t total, int lane) {
    const unsigned long long tag = (unsigned long long)epoch << 34;
    volatile unsigned long long *vs = state;
    uint32_t pre = 0;
    if (tile > 0) {
        if (lane == 0) vs[tile] = tag | kScanAgg | total;
        long long look = (long long)tile - 1;
        for (;;) {
            const long long t = look - lane;
            const bool need = t >= 0;
            unsigned long long sv = 0;
            uint32_t polls = 0;
            bool ready;
            do {
                if (need) sv = vs[t];
                ready = !need || ((sv >> 34) == epoch && ((sv >> 32) & 3ull) != 0);
                if (++polls > (1u << 26)) __trap();  // a predecessor never published: fail loudly, do not hang
            } while (!__all_sync(PCR_FULL, ready));
            const unsigned pm = __ballot_sync(PCR_FULL, need && ((sv >> 32) & 3ull) == 2ull);
            const int stop = pm ? __ffs(pm) - 1 : 32;
            pre += __reduce_add_sync(PCR_FULL, (need && lane <= stop) ? (uint32_t)sv : 0u);
            if (pm || look < 32) break;
            look -= 32;
        }
    }
    if (lane == 0) vs[tile] = tag | kScanPre | (uint32_t)(pre + total);
    return pre;
}

__global__ void __launch_bounds__(kScanThreads) scan_lookback_kernel(uint32_t *__restrict__ data, size_t n,
                                                                     unsigned long long *state, uint32_t *ticket,
                                                                     uint32_t epoch, uint32_t tiles, int vec) {
    constexpr int kWarps = kScanThreads / 32, kRows = kScanItems / 4;
    __shared__ uint32_t s_tile, s_pre;
    __shared__ uint32_t warp_tot[kWarps];
    PCR_GRID_DEP_SYNC();
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= tiles) return;  // (a ticket left over from an aborted scan: never write out of bounds)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // a warp owns 32 * kScanItems consecutive elements, read as kRows coalesced rows of uint4
    const size_t wbase = (size_t)tile * kScanTile + (size_t)w * (kScanItems * 32);
    uint4 v[kRows];
    uint32_t ex[kRows];
    uint32_t carry = 0;
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        v[r] = scan_load4(data, wbase + (size_t)r * 128 + lane * 4, n, vec != 0);
        const uint32_t s = v[r].x + v[r].y + v[r].z + v[r].w;
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(PCR_FULL, incl, o);
            if (lane >= o) incl += u;
        }
        ex[r] = carry + incl - s;
        carry += __shfl_sync(PCR_FULL, incl, 31);
    }
    if (lane == 0) warp_tot[w] = carry;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int j = 0; j < kWarps; j++) {
        const uint32_t t = warp_tot[j];
        if (j < w) woff += t;
        total += t;
    }
    if (w == 0) {
        const uint32_t pre = scan_lookback_warp0(state, tile, epoch, total, lane);
        if (lane == 0) s_pre = pre;
    }
    __syncthreads();
    const uint32_t base = s_pre + woff;
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        uint4 o;
        o.x = base + ex[r];
        o.y = o.x + v[r].x;
        o.z = o.y + v[r].y;
        o.w = o.z + v[r].z;
        scan_store4(data, wbase + (size_t)r * 128 + lane * 4, n, vec != 0, o);
    }
    if (tile == tiles - 1 && threadIdx.x == 0) *ticket = 0;  // every ticket has been handed out by now
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_u64_kernel(const uint32_t *__restrict__ in,
                                                                      uint64_t *__restrict__ out, size_t n,
                                                                      const uint32_t *__restrict__ tile_sums) {
    // two-pass variant (tile sums, then apply) widening to u64 and writing out[n] = total as well
    __shared__ unsigned long long sh64[kScanThreads / 32];
    __shared__ unsigned long long warp_off[kScanThreads / 32];
    unsigned long long pre = 0;
    for (uint32_t t = threadIdx.x; t < blockIdx.x; t += kScanThreads) pre += tile_sums[t];
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(PCR_FULL, pre, o);
    if ((threadIdx.x & 31) == 0) sh64[threadIdx.x >> 5] = pre;
    __syncthreads();
    pre = 0;
    for (int i = 0; i < kScanThreads / 32; i++) pre += sh64[i];
    __syncthreads();
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    unsigned long long tsum = 0;
#pragma unroll
    for (int it = 0; it < kScanItems; it++) {
        size_t i = base + it;
        v[it] = i < n ? in[i] : 0u;
        tsum += v[it];
    }
    unsigned long long incl = tsum;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long u = __shfl_up_sync(PCR_FULL, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) warp_off[w] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < kScanThreads / 32; i++) {
            unsigned long long t = warp_off[i];
            warp_off[i] = acc;
            acc += t;
        }
    }
    __syncthreads();
    unsigned long long run = pre + warp_off[w] + (incl - tsum);
#pragma unroll
    for (int it = 0; it < kScanItems; it++) {
        size_t i = base + it;
        if (i < n) out[i] = run;
        run += v[it];
        if (i + 1 == n) out[n] = run;
    }
}

}  // namespace

// (scratch: ticket + tile states; scans that may run at the same time on different streams need their own)
static int exclusive_scan_u32_with(Ctx *ctx, uint32_t *d_data, size_t n, DevBuf &scratch, uint32_t &scan_epoch) {
    if (n == 0) return PCR_OK;
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    const size_t cap_before = scratch.cap;
    PCR_TRY(ensure(ctx, scratch, 16 + tiles * sizeof(unsigned long long)));
    if (scratch.cap != cap_before || scan_epoch >= (1u << 30) - 2) {  // fresh memory (or epoch wrap): no tag may match
        PCR_CUDA(ctx, cudaMemsetAsync(scratch.p, 0, scratch.cap, ctx->stream));
        scan_epoch = 0;
    }
    const uint32_t epoch = ++scan_epoch;
    uint32_t *ticket = (uint32_t *)scratch.p;
    unsigned long long *state = (unsigned long long *)((char *)scratch.p + 16);
    const int vec = ((uintptr_t)d_data & 15) == 0;
    PCR_CUDA(ctx, launch_chained(scan_lookback_kernel, dim3((unsigned)tiles), dim3(kScanThreads), 0, ctx->stream, d_data, n, state, ticket, epoch,
                                 (uint32_t)tiles, vec));
    ctx->launches++;
    return PCR_OK;
}

int exclusive_scan_u32_dev(Ctx *ctx, uint32_t *d_data, size_t n) { return exclusive_scan_u32_with(ctx, d_data, n, ctx->b_scan, ctx->scan_epoch); }

int exclusive_scan_u64_from_u32_dev(Ctx *ctx, const uint32_t *d_in, uint64_t *d_out, size_t n) {
    if (n == 0) {
        PCR_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(uint64_t), ctx->stream));
        return PCR_OK;
    }
    size_t tiles = (n + kScanTile - 1) / kScanTile;
    PCR_TRY(ensure(ctx, ctx->b_misc2, tiles * sizeof(uint32_t)));
    uint32_t *tile_sums = (uint32_t *)ctx->b_misc2.p;
    scan_tile_sums_kernel<<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(d_in, n, tile_sums);
    PCR_LAUNCH_CHECK(ctx);
    scan_apply_u64_kernel<<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(d_in, d_out, n, tile_sums);
    PCR_LAUNCH_CHECK(ctx);
    return PCR_OK;
}

// ------------------------------------------------------------------------------------------------
// host side: cell-size selection and build orchestration
// ------------------------------------------------------------------------------------------------
namespace {

constexpr uint64_t kMaxCellsPerFrame = 1ull << 24;
constexpr uint64_t kMaxCellsTotal = 1ull << 27;

struct FrameBox {
    double mn[3], ext[3];
    uint32_t count;
};

// occupancy target: points sharing a cell with a typical point.  Tuned with the scalar model
// (oracle/orc_grid_knn_model) on the KITTI / aerial / uniform-cube shapes.
// Self-queries with k <= 24 run as cell tiles (knn_tile.cuh): one shell only, so the cells are sized for the k-th
// neighbour to lie inside the 27-cell cube's guaranteed radius for all but ~1 % of the queries (swept on the 122 K frame:
// 0.4 -> 8 758 continuations, 0.5 -> 1 456, 0.6 -> 3 360 with more of the frame in the dense class).
double occupancy_target(size_t k_hint, bool self_knn) {
    double k = k_hint ? (double)k_hint : 10.0;
    static const bool tile_off = [] {
        const char *e = getenv("PCR_KNN_TILE");
        return e && !strcmp(e, "0");
    }();
    double scale = self_knn && !tile_off && k_hint <= 24 ? 0.5 : 0.25;
    if (const char *e = getenv("PCR_OCC_SCALE")) scale = atof(e);  // tuning hook
    return std::min(64.0, std::max(2.0, scale * k));
}

double initial_cell(const FrameBox &b, double target) {
    if (b.count == 0) return 1.0;
    double e[3] = {b.ext[0], b.ext[1], b.ext[2]};
    std::sort(e, e + 3);
    double tiny = std::max(e[2] * 1e-6, 1e-30);
    double vol = std::max(e[0], tiny) * std::max(e[1], tiny) * std::max(e[2], tiny);
    double area = std::max(e[1], tiny) * std::max(e[2], tiny);
    double h3 = std::cbrt(vol * target / b.count);
    double h2 = std::sqrt(area * target / b.count);
    double h = std::sqrt(h2 * h3);
    if (!(h > 0) || !std::isfinite(h)) h = 1.0;
    return h;
}

// Fill dims / permutation for cell size h, growing h until the dense table fits the cap.
void shape_grid(GridDesc &g, const FrameBox &b, double h, uint64_t cap) {
    if (b.count == 0) h = 1.0;
    int dims[3];
    for (int iter = 0; iter < 64; iter++) {
        double inv = 1.0 / h;
        double cells = 1.0;
        for (int a = 0; a < 3; a++) {
            double d = std::floor(b.ext[a] * inv) + 1.0;
            if (!(d >= 1.0)) d = 1.0;
            cells *= d;
            dims[a] = d > 2147483000.0 ? 2147483000 : (int)d;
        }
        if (cells <= (double)cap) break;
        h *= std::cbrt(cells / (double)cap) * 1.02;
    }
    int order[3] = {0, 1, 2};
    std::stable_sort(order, order + 3, [&](int p, int q) { return dims[p] < dims[q]; });
    g.h = h;
    g.inv_h = 1.0 / h;
    for (int j = 0; j < 3; j++) {
        g.ax[j] = order[j];
        g.dims[j] = dims[order[j]];
        g.o[j] = b.count ? b.mn[order[j]] : 0.0;
    }
    g.n_cells = (uint32_t)((uint64_t)g.dims[0] * g.dims[1] * g.dims[2]);
}

}  // namespace

void index_free(Index *ix) {
    if (!ix) return;
    if (ix->coarser) index_free(ix->coarser);
    if (ix->finer) index_free(ix->finer);
    if (ix->owns_memory && ix->ctx) {
        cudaStream_t s = ix->ctx->stream;
        if (ix->grids && !ix->level_bufs_borrowed) cudaFreeAsync(ix->grids, s);
        if (ix->sorted && !ix->level_bufs_borrowed) cudaFreeAsync(ix->sorted, s);
        if (ix->cell_start && ix->cell_slot < 0) cudaFreeAsync(ix->cell_start, s);
        if (!ix->shares_orig4) {
            if (ix->frame_in_off) cudaFreeAsync(ix->frame_in_off, s);
            if (ix->orig4) cudaFreeAsync(ix->orig4, s);
        }
    }
    delete ix;
}

int index_apply_mask_dev(Index *ix, const uint8_t *d_keep) {
    Ctx *ctx = ix->ctx;
    TombLevels t = {};
    uint32_t n_max = 0;
    auto flush = [&]() -> int {  // (all levels of an index fit one launch; the loop is for generality)
        if (t.levels && n_max) {
            PCR_CUDA(ctx, launch_chained(tombstone_kernel, dim3((n_max + 255) / 256), dim3(256), 0, ctx->stream, t, d_keep));
            ctx->launches++;
        }
        t = {};
        n_max = 0;
        return PCR_OK;
    };
    auto add = [&](Index *l) -> int {
        l->d_mask = d_keep;  // levels built from now on leave the removed points out
        if (!l->n_indexed) return PCR_OK;
        t.sorted[t.levels] = l->sorted;
        t.n[t.levels] = (uint32_t)l->n_indexed;
        n_max = std::max(n_max, t.n[t.levels]);
        if (++t.levels == 4) PCR_TRY(flush());
        return PCR_OK;
    };
    for (Index *l = ix; l; l = l->coarser) PCR_TRY(add(l));
    if (ix->finer) PCR_TRY(add(ix->finer));
    return flush();
}

// Next-coarser level: same points, same frames, cell size x kLevelFactor.  Reuses the bounding
// boxes and the AoS copy of level 0, so it costs one count + scan + scatter (no host round trip).
// fine = the other direction: cell size / 2 (or as fine as the cell-table cap allows), attached as ix->finer; it is built
// on its own side stream at the same time as the coarser level, hence its own scratch.
static int index_level_build(Index *ix, bool fine, Index **out) {
    Index *&slot_ptr = fine ? ix->finer : ix->coarser;
    if (slot_ptr) {
        *out = slot_ptr;
        return PCR_OK;
    }
    Ctx *ctx = ix->ctx;
    cudaStream_t st = ctx->stream;
    const int F = ix->n_frames;
    const size_t n = ix->n;
    Index *c = new (std::nothrow) Index();
    if (!c) return fail(ctx, PCR_ERR_OOM, "host allocation failed");
    c->ctx = ctx;
    c->n = n;
    c->n_indexed = ix->n_indexed;
    c->n_frames = F;
    c->grids_h.resize(F);
    c->box_min = ix->box_min;
    c->box_ext = ix->box_ext;
    c->box_count = ix->box_count;
    c->d_mask = ix->d_mask;
    c->orig4 = ix->orig4;
    c->frame_in_off = ix->frame_in_off;
    c->shares_orig4 = true;
    struct Guard {
        Index *ix;
        ~Guard() {
            if (ix) index_free(ix);
        }
    } guard{c};
    constexpr int kSlots = (int)(sizeof(ctx->b_cells) / sizeof(ctx->b_cells[0]));
    if (fine ? ix->cell_slot == 0 : (ix->cell_slot >= 0 && ix->cell_slot + 1 < kSlots - 1)) {
        c->cell_slot = fine ? kSlots - 1 : ix->cell_slot + 1;  // (the last slot is the finer level's)
        c->level_bufs_borrowed = true;  // a level of a transient index: every array comes from the context's slot
        PCR_TRY(ensure(ctx, ctx->b_lgrids[c->cell_slot], sizeof(GridDesc) * F));
        c->grids = (GridDesc *)ctx->b_lgrids[c->cell_slot].p;
    } else {
        PCR_CUDA(ctx, cudaMallocAsync((void **)&c->grids, sizeof(GridDesc) * F, st));
    }
    const uint64_t cap = std::max<uint64_t>(1, std::min<uint64_t>(kMaxCellsPerFrame, kMaxCellsTotal / (uint64_t)F));
    uint64_t base = 0;
    uint32_t max_frame = 0;
    for (int f = 0; f < F; f++) {
        FrameBox b;
        for (int a = 0; a < 3; a++) {
            b.mn[a] = ix->box_min[3 * f + a];
            b.ext[a] = ix->box_ext[3 * f + a];
        }
        b.count = ix->box_count[f];
        GridDesc &g = c->grids_h[f];
        double factor = kLevelFactor;
        if (const char *ev = getenv("PCR_LEVEL_FACTOR")) factor = atof(ev);  // tuning hook
        if (fine) {
            factor = 0.6;  // (swept on the 122 K frame: 0.35 -> 0.592 ms per step, 0.42 -> 0.530, 0.5 -> 0.523, 0.6 -> 0.511, 0.7 -> 0.515)
            if (const char *ev = getenv("PCR_FINE_FACTOR")) factor = atof(ev);  // tuning hook
        }
        shape_grid(g, b, ix->grids_h[f].h * factor, cap);
        g.cell_base = (uint32_t)base;
        base += g.n_cells;
        g.pt_begin = ix->grids_h[f].pt_begin;
        g.pt_end = ix->grids_h[f].pt_end;
        g.in_begin = ix->grids_h[f].in_begin;
        g.in_end = ix->grids_h[f].in_end;
        max_frame = std::max(max_frame, g.in_end - g.in_begin);
    }
    c->total_cells = (uint32_t)base;
    PCR_CUDA(ctx, cudaMemcpyAsync(c->grids, c->grids_h.data(), sizeof(GridDesc) * F, cudaMemcpyHostToDevice, st));
    if (c->level_bufs_borrowed) {
        PCR_TRY(ensure(ctx, ctx->b_cells[c->cell_slot], sizeof(uint32_t) * ((size_t)base + 1)));
        c->cell_start = (uint32_t *)ctx->b_cells[c->cell_slot].p;
        PCR_TRY(ensure(ctx, ctx->b_lsorted[c->cell_slot], sizeof(float4) * std::max<size_t>(n, 1)));
        c->sorted = (float4 *)ctx->b_lsorted[c->cell_slot].p;
    } else {
        PCR_CUDA(ctx, cudaMallocAsync((void **)&c->cell_start, sizeof(uint32_t) * ((size_t)base + 1), st));
        PCR_CUDA(ctx, cudaMallocAsync((void **)&c->sorted, sizeof(float4) * std::max<size_t>(n, 1), st));
    }
    PCR_CUDA(ctx, cudaMemsetAsync(c->cell_start, 0, sizeof(uint32_t) * ((size_t)base + 1), st));
    DevBuf &scratch = fine ? ctx->b_fine_misc : ctx->b_misc;
    PCR_TRY(ensure(ctx, scratch, sizeof(uint32_t) * 2 * std::max<size_t>(n, 1)));
    uint32_t *d_cell_id = (uint32_t *)scratch.p;
    uint32_t *d_rank = d_cell_id + std::max<size_t>(n, 1);
    const uint32_t ns = (uint32_t)ix->n_indexed;
    if (ns > 0) {
        count_sorted_kernel<<<(ns + kBuildThreads - 1) / kBuildThreads, kBuildThreads, 0, st>>>(ix->sorted, ns, c->grids, F, c->cell_start,
                                                                                               d_cell_id, d_rank);
        PCR_LAUNCH_CHECK(ctx);
    }
    if (fine) PCR_TRY(exclusive_scan_u32_with(ctx, c->cell_start, (size_t)base + 1, ctx->b_fine_scan, ctx->fine_scan_epoch));
    else PCR_TRY(exclusive_scan_u32_dev(ctx, c->cell_start, (size_t)base + 1));
    if (ns > 0) {
        PCR_CUDA(ctx, launch_chained(scatter_sorted_kernel, dim3((ns + kBuildThreads - 1) / kBuildThreads), dim3(kBuildThreads), 0, st, ix->sorted, ns,
                                     c->cell_start, d_cell_id, d_rank, c->sorted));
        ctx->launches++;
    }
    guard.ix = nullptr;
    slot_ptr = c;
    *out = c;
    return PCR_OK;
}

int index_coarser_level(Index *ix, Index **out) { return index_level_build(ix, false, out); }
int index_finer_level(Index *ix, Index **out) { return index_level_build(ix, true, out); }

int index_build_dev(Ctx *ctx, const float *dx, const float *dy, const float *dz, size_t n, const BuildOpts &opts,
                    Index **out) {
    *out = nullptr;
    if (n > 0x7fffffffull) return fail(ctx, PCR_ERR_UNSUPPORTED, "clouds above 2^31 points are not supported");
    TimeScope ts(ctx, kTagBuild);
    const int F = opts.n_frames > 0 ? opts.n_frames : 1;
    if (F > 1 && !opts.frame_offsets) return fail(ctx, PCR_ERR_INVALID_ARG, "frame_offsets missing");
    cudaStream_t st = ctx->stream;

    Index *ix = new (std::nothrow) Index();
    if (!ix) return fail(ctx, PCR_ERR_OOM, "host allocation failed");
    ix->ctx = ctx;
    ix->n = n;
    ix->n_frames = F;
    ix->grids_h.resize(F);
    struct Guard {  // frees the half-built index on any early return
        Index *ix;
        ~Guard() {
            if (ix) index_free(ix);
        }
    } guard{ix};

    PCR_CUDA(ctx, cudaMallocAsync((void **)&ix->grids, sizeof(GridDesc) * F, st));
    std::vector<uint32_t> off_h(F + 1);
    if (F > 1) {
        for (int f = 0; f <= F; f++) {
            if (opts.frame_offsets[f] > n || (f && opts.frame_offsets[f] < opts.frame_offsets[f - 1]))
                return fail(ctx, PCR_ERR_INVALID_ARG, "frame_offsets must be non-decreasing and end at n");
            off_h[f] = (uint32_t)opts.frame_offsets[f];
        }
        if (off_h[0] != 0 || off_h[F] != n) return fail(ctx, PCR_ERR_INVALID_ARG, "frame_offsets must span [0, n]");
        PCR_CUDA(ctx, cudaMallocAsync((void **)&ix->frame_in_off, sizeof(uint32_t) * (F + 1), st));
        PCR_CUDA(ctx, cudaMemcpyAsync(ix->frame_in_off, off_h.data(), sizeof(uint32_t) * (F + 1), cudaMemcpyHostToDevice, st));
    } else {
        off_h[0] = 0;
        off_h[1] = (uint32_t)n;
    }
    uint32_t max_frame = 0;
    for (int f = 0; f < F; f++) max_frame = std::max(max_frame, off_h[f + 1] - off_h[f]);

    if (n > 0) {
        PCR_CUDA(ctx, cudaMallocAsync((void **)&ix->orig4, sizeof(float4) * n, st));
        PCR_CUDA(ctx, cudaMallocAsync((void **)&ix->sorted, sizeof(float4) * n, st));
    }

    const bool dbg_t = getenv("PCR_DEBUG_BUILD") != nullptr;
    auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now_us();
    // ---- K0: bounding boxes -------------------------------------------------------------------
    PCR_TRY(ensure(ctx, ctx->b_small, sizeof(FrameStats) * F + sizeof(ProbeStats) * F + 256));
    PCR_TRY(ensure_pinned(ctx, sizeof(FrameStats) * F + sizeof(ProbeStats) * F + 256));
    FrameStats *d_stats = (FrameStats *)ctx->b_small.p;
    ProbeStats *d_probe = (ProbeStats *)((char *)ctx->b_small.p + sizeof(FrameStats) * F);
    FrameStats *h_stats = (FrameStats *)ctx->pinned;
    ProbeStats *h_probe = (ProbeStats *)((char *)ctx->pinned + sizeof(FrameStats) * F);
    if (opts.known_stats && opts.known_stats->valid && F == 1 && !opts.d_mask) {
        h_stats[0] = *opts.known_stats;  // the step that wrote the cloud measured it: no kernel, no round trip
    } else {
    init_stats_kernel<<<(F + 255) / 256, 256, 0, st>>>(d_stats, F);
    PCR_LAUNCH_CHECK(ctx);
    if (n > 0) {
        unsigned bx = (unsigned)std::min<size_t>((max_frame + kBuildThreads * 4 - 1) / (kBuildThreads * 4),
                                                 F > 1 ? 64 : (size_t)ctx->sm_count * 4);
        bx = std::max(bx, 1u);
        bbox_kernel<<<dim3(bx, F), kBuildThreads, 0, st>>>(dx, dy, dz, ix->frame_in_off, n, opts.d_mask, d_stats);
        PCR_LAUNCH_CHECK(ctx);
    }
    PCR_CUDA(ctx, cudaMemcpyAsync(h_stats, d_stats, sizeof(FrameStats) * F, cudaMemcpyDeviceToHost, st));
    PCR_CUDA(ctx, cudaStreamSynchronize(st));
    }

    std::vector<FrameBox> box(F);
    size_t n_indexed = 0;
    for (int f = 0; f < F; f++) {
        box[f].count = h_stats[f].count;
        for (int a = 0; a < 3; a++) {
            if (box[f].count) {
                double lo = ordered_f32(h_stats[f].mn[a]), hi = ordered_f32(h_stats[f].mx[a]);
                box[f].mn[a] = lo;
                box[f].ext[a] = hi - lo;
            } else {
                box[f].mn[a] = 0;
                box[f].ext[a] = 0;
            }
        }
        n_indexed += box[f].count;
    }
    ix->n_indexed = n_indexed;
    ix->d_mask = opts.d_mask;
    ix->box_min.resize(3 * F);
    ix->box_ext.resize(3 * F);
    ix->box_count.resize(F);
    for (int f = 0; f < F; f++) {
        ix->box_count[f] = box[f].count;
        for (int a = 0; a < 3; a++) {
            ix->box_min[3 * f + a] = box[f].mn[a];
            ix->box_ext[3 * f + a] = box[f].ext[a];
        }
    }

    const double target = occupancy_target(opts.k_hint, opts.self_knn);
    const size_t hint_key = opts.k_hint + (opts.self_knn ? (size_t)1 << 20 : 0);  // (what the cached cell size was chosen for)
    const uint64_t cap = std::max<uint64_t>(1, std::min<uint64_t>(kMaxCellsPerFrame, kMaxCellsTotal / (uint64_t)F));
    std::vector<double> hsel(F);
    const bool forced = ctx->forced_cell > 0.f;
    for (int f = 0; f < F; f++) hsel[f] = forced ? (double)ctx->forced_cell : initial_cell(box[f], target);

    PCR_TRY(ensure(ctx, ctx->b_misc, sizeof(uint32_t) * 2 * std::max<size_t>(n, 1)));
    uint32_t *d_cell_id = (uint32_t *)ctx->b_misc.p;
    uint32_t *d_rank = d_cell_id + std::max<size_t>(n, 1);

    // lays the frames' tables back to back and uploads the descriptors
    auto layout = [&](uint32_t *total_cells) -> int {
        uint64_t base = 0;
        uint32_t pt = 0;
        for (int f = 0; f < F; f++) {
            GridDesc &g = ix->grids_h[f];
            shape_grid(g, box[f], hsel[f], cap);
            g.cell_base = (uint32_t)base;
            base += g.n_cells;
            g.pt_begin = pt;
            pt += box[f].count;
            g.pt_end = pt;
            g.in_begin = off_h[f];
            g.in_end = off_h[f + 1];
        }
        if (base >= 0xffffffffull) return fail(ctx, PCR_ERR_UNSUPPORTED, "cell table too large");
        *total_cells = (uint32_t)base;
        PCR_CUDA(ctx, cudaMemcpyAsync(ix->grids, ix->grids_h.data(), sizeof(GridDesc) * F, cudaMemcpyHostToDevice, st));
        return PCR_OK;
    };
    const unsigned count_bx = std::max(1u, (max_frame + kBuildThreads * kItems - 1) / (kBuildThreads * kItems));

    if (dbg_t) { cudaStreamSynchronize(st); fprintf(stderr, "[build] bbox done +%.0f us\n", now_us() - t_begin); }
    // ---- probe rounds: measure occupancy, solve for the cell size ------------------------------
    bool cached = false;
    if (!forced && F == 1 && n_indexed > 0 && ctx->frame_stream && ctx->cell_cache.valid && ctx->cell_cache.k_hint == hint_key &&
        !getenv("PCR_NO_CELL_CACHE")) {
        const auto &cc = ctx->cell_cache;
        auto close = [](double a, double b) { return a <= b * 1.125 + 1e-9 && b <= a * 1.125 + 1e-9; };
        if (close((double)box[0].count, (double)cc.count) && close(box[0].ext[0], cc.ext[0]) && close(box[0].ext[1], cc.ext[1]) &&
            close(box[0].ext[2], cc.ext[2])) {
            hsel[0] = cc.h;
            cached = true;
        }
        if (cached) ctx->stat_cell_hits++;
        else ctx->stat_cell_misses++;
    }
    if (!forced && !cached && n_indexed > 0) {
        for (int round = 0; round < 2; round++) {
            uint32_t total = 0;
            PCR_TRY(layout(&total));
            PCR_TRY(ensure(ctx, ctx->b_table, sizeof(uint32_t) * ((size_t)total + 1)));
            uint32_t *d_tab = (uint32_t *)ctx->b_table.p;
            PCR_CUDA(ctx, cudaMemsetAsync(d_tab, 0, sizeof(uint32_t) * ((size_t)total + 1), st));
            PCR_CUDA(ctx, cudaMemsetAsync(d_probe, 0, sizeof(ProbeStats) * F, st));
            count_kernel<<<dim3(count_bx, F), kBuildThreads, 0, st>>>(dx, dy, dz, ix->frame_in_off, n, opts.d_mask, ix->grids,
                                                                      d_tab, d_cell_id, d_rank, nullptr);
            PCR_LAUNCH_CHECK(ctx);
            unsigned sbx = (unsigned)std::min<size_t>((max_frame + kBuildThreads * 4 - 1) / (kBuildThreads * 4),
                                                      F > 1 ? 64 : (size_t)ctx->sm_count * 4);
            stats_kernel<<<dim3(std::max(sbx, 1u), F), kBuildThreads, 0, st>>>(ix->frame_in_off, n, ix->grids, d_tab,
                                                                               d_cell_id, d_probe);
            PCR_LAUNCH_CHECK(ctx);
            PCR_CUDA(ctx, cudaMemcpyAsync(h_probe, d_probe, sizeof(ProbeStats) * F, cudaMemcpyDeviceToHost, st));
            PCR_CUDA(ctx, cudaStreamSynchronize(st));
            bool again = false;
            for (int f = 0; f < F; f++) {
                if (!box[f].count) continue;
                double l1 = h_probe[f].sum_log_own / box[f].count;     // log2 m(h)
                double l2 = h_probe[f].sum_log_super / box[f].count;   // log2 m(2h)
                double D = std::min(3.0, std::max(1.0, l2 - l1));
                double ratio = std::exp2((std::log2(target) - l1) / D);
                ratio = std::min(6.0, std::max(1.0 / 6.0, ratio));
                double hprobe = ix->grids_h[f].h;  // (may have been grown by the cell cap)
                hsel[f] = hprobe * ratio;
                if (ratio > 2.5 || ratio < 0.4) again = true;
            }
            if (!again) break;
        }
    }

    if (dbg_t) { cudaStreamSynchronize(st); fprintf(stderr, "[build] probe done +%.0f us\n", now_us() - t_begin); }
    if (!forced && !cached && F == 1 && n_indexed > 0) {
        ctx->cell_cache.valid = true;
        ctx->cell_cache.k_hint = hint_key;
        ctx->cell_cache.count = box[0].count;
        for (int a = 0; a < 3; a++) ctx->cell_cache.ext[a] = box[0].ext[a];
        ctx->cell_cache.h = hsel[0];
    }
    // ---- final grid: count, scan, scatter ------------------------------------------------------
    uint32_t total = 0;
    PCR_TRY(layout(&total));
    ix->total_cells = total;
    if (opts.transient) {
        ix->cell_slot = 0;
        PCR_TRY(ensure(ctx, ctx->b_cells[0], sizeof(uint32_t) * ((size_t)total + 1)));
        ix->cell_start = (uint32_t *)ctx->b_cells[0].p;
    } else {
        PCR_CUDA(ctx, cudaMallocAsync((void **)&ix->cell_start, sizeof(uint32_t) * ((size_t)total + 1), st));
    }
    if (dbg_t) { fprintf(stderr, "[build] malloc issued +%.0f us\n", now_us() - t_begin); }
    PCR_CUDA(ctx, cudaMemsetAsync(ix->cell_start, 0, sizeof(uint32_t) * ((size_t)total + 1), st));
    if (dbg_t) { cudaStreamSynchronize(st); fprintf(stderr, "[build] memset done +%.0f us\n", now_us() - t_begin); }
    if (n > 0) {
        count_kernel<<<dim3(count_bx, F), kBuildThreads, 0, st>>>(dx, dy, dz, ix->frame_in_off, n, opts.d_mask, ix->grids,
                                                                  ix->cell_start, d_cell_id, d_rank, ix->orig4);
        PCR_LAUNCH_CHECK(ctx);
    }
    if (dbg_t) { cudaStreamSynchronize(st); fprintf(stderr, "[build] count done +%.0f us\n", now_us() - t_begin); }
    PCR_TRY(exclusive_scan_u32_dev(ctx, ix->cell_start, (size_t)total + 1));
    if (dbg_t) { cudaStreamSynchronize(st); fprintf(stderr, "[build] scan done +%.0f us\n", now_us() - t_begin); }
    if (n > 0) {
        PCR_CUDA(ctx, launch_chained(scatter_kernel, dim3((unsigned)((n + kBuildThreads - 1) / kBuildThreads)), dim3(kBuildThreads), 0, st,
                                     ix->orig4, n, ix->cell_start, d_cell_id, d_rank, ix->sorted));
        ctx->launches++;
    }
    if (dbg_t) { cudaStreamSynchronize(st); fprintf(stderr, "[build] final done +%.0f us (total cells %u)\n", now_us() - t_begin, total); }
    guard.ix = nullptr;
    *out = ix;
    return PCR_OK;
}

}  // namespace pcr
