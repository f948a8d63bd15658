// cluster.cu -- euclidean_cluster (crates/segmentation/src/euclidean_cluster.rs:96-187) on the GPU.
//
// The reference bins the finite points into cells of size r keyed by floor(x * (1/r)) as i32
// (:55-61), tests every pair of points that share a cell or sit in key-adjacent cells with
// dx*dx + dy*dy + dz*dz <= r*r (f32, :147-151) and takes the connected components of that graph.
// The components do not depend on the order the pairs are visited in, so the GPU version may test
// them in any order -- but it has to test exactly the SAME pair set, which is why the cells here are
// keyed by the reference's own f32 expression and not by the KNN grid's f64 cell coordinates (a pair
// closer than r whose keys differ by 2 through f32 rounding is missed by the reference, and must be
// missed here too).
//
//   insert_kernel    open-addressing hash table of cells (one representative point per slot); the
//                    return value of the per-cell counter atomic is the point's rank     12 B/pt + table
//   (scan)           exclusive scan of the table's counters -> cell start offsets
//   scatter_kernel   cell-ordered float4 copy (x, y, z, original index)                  28 B/pt
//   union_kernel     one thread per point: own cell (later positions only) + the 13 forward
//                    neighbour cells (:65-82); hits are queued per warp and united together by a
//                    lock-free union-find over cell-sorted positions with randomised linking
//   roots / labels   roots compressed, per root the smallest original index -> label[i]
// The host side of the entry point turns the labels into the reference's output order (size
// descending, then first index ascending, indices ascending): one counting pass over n labels.
#include "pcr_internal.cuh"

#include <algorithm>

namespace pcr {

namespace {

constexpr uint32_t kEmptySlot = 0xffffffffu;

// euclidean_cluster.rs:55-61: (v * inv_r).floor() as i32 -- cvt.rmi.s32.f32 rounds down, saturates and
// maps NaN to 0, exactly like Rust's `as i32` after floor()
__device__ __forceinline__ int ref_cell(float v, float inv_r) { return __float2int_rd(__fmul_rn(v, inv_r)); }

__device__ __forceinline__ uint32_t hash_cell(int kx, int ky, int kz, uint32_t mask) {
    unsigned long long k = ((unsigned long long)(uint32_t)kx << 32) | (uint32_t)ky;
    k ^= (unsigned long long)(uint32_t)kz * 0x9e3779b97f4a7c15ull;
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k & mask;
}

// The hash table stores, per occupied slot, the index of ONE point of that cell (its representative):
// a full i32 x 3 key does not fit a single CAS word, the representative does, and the key is
// recomputed from its coordinates whenever two cells have to be told apart.  No spinning, exact for
// every key the reference can form (including the saturated keys of far-away points).
__global__ void __launch_bounds__(256) cluster_insert_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                             const float *__restrict__ z, size_t n, float inv_r,
                                                             uint32_t *__restrict__ rep, uint32_t mask, uint32_t *__restrict__ count,
                                                             uint32_t *__restrict__ slot_of, uint32_t *__restrict__ rank_of) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float px = x[i], py = y[i], pz = z[i];
    if (!finite3(px, py, pz)) {  // :114-116
        slot_of[i] = kEmptySlot;
        return;
    }
    const int kx = ref_cell(px, inv_r), ky = ref_cell(py, inv_r), kz = ref_cell(pz, inv_r);
    uint32_t h = hash_cell(kx, ky, kz, mask);
    for (;;) {
        const uint32_t old = atomicCAS(&rep[h], kEmptySlot, (uint32_t)i);
        if (old == kEmptySlot) break;  // this point now represents the cell
        if (ref_cell(x[old], inv_r) == kx && ref_cell(y[old], inv_r) == ky && ref_cell(z[old], inv_r) == kz) break;
        h = (h + 1) & mask;
    }
    slot_of[i] = h;
    rank_of[i] = atomicAdd(&count[h], 1u);
}

__global__ void __launch_bounds__(256) cluster_scatter_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                              const float *__restrict__ z, size_t n,
                                                              const uint32_t *__restrict__ start, const uint32_t *__restrict__ slot_of,
                                                              const uint32_t *__restrict__ rank_of, float4 *__restrict__ cpts,
                                                              uint32_t *__restrict__ cslot) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot_of[i];
    if (s == kEmptySlot) return;
    const uint32_t pos = start[s] + rank_of[i];
    cpts[pos] = make_float4(x[i], y[i], z[i], __uint_as_float((uint32_t)i));
    cslot[pos] = s;
}

// Lock-free union-find over the CELL-SORTED positions (neighbours in space are neighbours in the
// parent array, so the walks stay in L1/L2 lines the pair loop has just touched); the root
// with the smaller PRIORITY wins.  Priorities only ever decrease along a walk, so neither the racy path-halving store nor
// a stale cached read can close a cycle or leave the component: a stale parent is still a member of
// the same tree.  Only the CAS decides a link, and it sees the true value.
__device__ __forceinline__ uint32_t uf_find(volatile uint32_t *parent, uint32_t v) {
    for (;;) {
        const uint32_t p = parent[v];
        if (p == v) return v;
        const uint32_t gp = parent[p];
        if (gp != p) parent[v] = gp;
        v = p;
    }
}

// Link order: a pseudo-random priority of the position, not the position itself.  Positions are in
// cell order, so "smaller position wins" builds chains as long as a row of cells (measured: 9.5 ms for
// the union kernel); with random priorities the expected depth is logarithmic, as in randomised linking.
__device__ __forceinline__ uint32_t uf_prio(uint32_t v) {
    v *= 0x9e3779b1u;
    v ^= v >> 15;
    v *= 0x85ebca77u;
    v ^= v >> 13;
    return v;
}
__device__ __forceinline__ bool uf_before(uint32_t a, uint32_t b) {  // a becomes (stays) the root when linked with b
    const uint32_t pa = uf_prio(a), pb = uf_prio(b);
    return pa != pb ? pa < pb : a < b;
}

// one edge: link the trees of positions a and b (exact walks with halving, CAS decides)
__device__ __forceinline__ void uf_unite(uint32_t *parent, uint32_t a, uint32_t b) {
    if (parent[a] == parent[b]) return;  // already hanging under the same node: the common case once components have formed
    uint32_t ra = uf_find(parent, a), rb = uf_find(parent, b);
    while (ra != rb) {
        const bool a_first = uf_before(ra, rb);
        const uint32_t lo = a_first ? ra : rb, hi = a_first ? rb : ra;
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;  // hi was still a root: linked under the one that comes first
        ra = uf_find(parent, ra);
        rb = uf_find(parent, rb);
    }
}

__constant__ int kHalfOff[13][3] = {{1, 0, 0},  {1, 1, 0},   {1, -1, 0}, {1, 0, 1}, {1, 0, -1}, {1, 1, 1}, {1, 1, -1},
                                    {1, -1, 1}, {1, -1, -1}, {0, 1, 0},  {0, 1, 1}, {0, 1, -1}, {0, 0, 1}};  // :65-82

constexpr int kUnionWarps = 4;

// One thread per point walks its own cell (later positions only, :145) and the 13 forward neighbour
// cells.  The pair test is cheap and regular; the union that follows a hit is a serial pointer walk that
// different lanes reach at different times -- executed in place it ran with 3.5 of 32 lanes active
// (ncu, 191 M warp instructions, 2.8 ms on the 122 K frame).  So hits are only QUEUED (warp-aggregated
// push into shared memory) and the warp drains the queue together, one edge per lane.
__global__ void __launch_bounds__(32 * kUnionWarps) cluster_union_kernel(const float4 *__restrict__ cpts, const uint32_t *__restrict__ cslot,
                                                                         uint32_t m_bound, const float *__restrict__ x,
                                                                         const float *__restrict__ y, const float *__restrict__ z,
                                                                         const uint32_t *__restrict__ rep, uint32_t mask,
                                                                         const uint32_t *__restrict__ start, float inv_r, float r2,
                                                                         uint32_t *parent /* by position */) {
    __shared__ uint2 queue[kUnionWarps][64];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = min(m_bound, start[mask + 1]);  // start[T] = number of finite points
    const bool live = t < m;
    if (!__any_sync(PCR_FULL, live)) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t s = 0;
    if (live) {
        a = cpts[t];
        s = cslot[t];
    }
    int qn = 0;  // entries in this warp's queue (warp-uniform)
    auto drain = [&](int keep_below) {
        while (qn > keep_below) {
            const int take = min(qn, 32);
            uint2 e = make_uint2(0, 0);
            if (lane < take) e = queue[w][qn - take + lane];
            __syncwarp();
            if (lane < take) uf_unite(parent, e.x, e.y);
            qn -= take;
            __syncwarp();
        }
    };
    auto scan = [&](uint32_t b, uint32_t e) {  // warp-uniform loop, per-lane ranges; the pair test of :147-151
        for (uint32_t j = b;; j++) {
            const bool act = j < e;
            if (!__any_sync(PCR_FULL, act)) break;
            bool hit = false;
            if (act) {
                const float4 p = __ldg(&cpts[j]);
                hit = dist2_exact(a.x, a.y, a.z, p.x, p.y, p.z) <= r2;
            }
            const unsigned hm = __ballot_sync(PCR_FULL, hit);
            if (hm) {
                if (hit) queue[w][qn + __popc(hm & ((1u << lane) - 1u))] = make_uint2(t, j);
                qn += __popc(hm);
                __syncwarp();
                if (qn >= 32) drain(31);
            }
        }
    };
    scan(live ? t + 1 : 0, live ? start[s + 1] : 0);  // own cell: every unordered pair once (:145)
    const int kx = ref_cell(a.x, inv_r), ky = ref_cell(a.y, inv_r), kz = ref_cell(a.z, inv_r);
#pragma unroll 1
    for (int o = 0; o < 13; o++) {
        uint32_t b = 0, e = 0;
        if (live) {
            // cx + dx on i32 (wrapping, as in a release build of the reference; only saturated keys get there)
            const int nx = (int)((uint32_t)kx + (uint32_t)kHalfOff[o][0]), ny = (int)((uint32_t)ky + (uint32_t)kHalfOff[o][1]);
            const int nz = (int)((uint32_t)kz + (uint32_t)kHalfOff[o][2]);
            uint32_t h = hash_cell(nx, ny, nz, mask);
            for (;;) {
                const uint32_t r = __ldg(&rep[h]);
                if (r == kEmptySlot) break;
                if (ref_cell(__ldg(&x[r]), inv_r) == nx && ref_cell(__ldg(&y[r]), inv_r) == ny && ref_cell(__ldg(&z[r]), inv_r) == nz) {
                    b = start[h];
                    e = start[h + 1];
                    break;
                }
                h = (h + 1) & mask;
            }
        }
        scan(b, e);
    }
    drain(0);
}

__global__ void __launch_bounds__(256) cluster_init_parent_kernel(uint32_t *__restrict__ parent, uint32_t *__restrict__ min_idx, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    parent[i] = (uint32_t)i;
    min_idx[i] = 0xffffffffu;
}

// root position of every finite point (parents fully compressed), and per root the smallest ORIGINAL index
__global__ void __launch_bounds__(256) cluster_roots_kernel(uint32_t *__restrict__ parent, const float4 *__restrict__ cpts, uint32_t m_bound,
                                                            const uint32_t *__restrict__ start, uint32_t mask, uint32_t *__restrict__ min_idx) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < m_bound && t < start[mask + 1];
    const unsigned lm = __ballot_sync(PCR_FULL, live);
    if (!live) return;
    const uint32_t r = uf_find(parent, t);
    parent[t] = r;
    // one atomic per distinct root in the warp (a 108 K-point ground plane is ONE root: 79 us of
    // same-address atomics without this)
    const unsigned peers = __match_any_sync(lm, r);
    const uint32_t mn = __reduce_min_sync(peers, __float_as_uint(cpts[t].w));
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicMin(&min_idx[r], mn);
}

__global__ void __launch_bounds__(256) cluster_labels_kernel(const uint32_t *__restrict__ parent, const uint32_t *__restrict__ min_idx,
                                                             const uint32_t *__restrict__ start, const uint32_t *__restrict__ slot_of,
                                                             const uint32_t *__restrict__ rank_of, size_t n, uint32_t *__restrict__ labels) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot_of[i];
    if (s == kEmptySlot) {  // non-finite point: its own component (:164-168)
        labels[i] = (uint32_t)i;
        return;
    }
    const uint32_t pos = start[s] + rank_of[i];
    labels[i] = min_idx[parent[parent[pos]]];  // parent[pos] is a root after cluster_roots_kernel (or one step from it)
}

}  // namespace

// labels[i] = smallest index of the connected component of point i (non-finite points: i itself).
// Preconditions (checked by the callers): n > 0, threshold > 0 or NaN.  No host round trip.
int cluster_labels_dev(Ctx *ctx, const float *dx, const float *dy, const float *dz, size_t n, float threshold, uint32_t *d_labels) {
    if (n > 0x3fffffffull) return fail(ctx, PCR_ERR_UNSUPPORTED, "euclidean_cluster: clouds above 2^30 points are not supported");
    cudaStream_t st = ctx->stream;
    TimeScope ts(ctx, kTagOther);
    const float inv_r = 1.0f / threshold;    // :107
    const float r2 = threshold * threshold;  // :108
    uint32_t T = 256;
    while (T < 2u * (uint32_t)n) T <<= 1;
    // scratch: cpts[n] | rep[T] | count[T+1] | slot_of[n] | rank_of[n] | cslot[n] | parent[n] | min_idx[n]
    const size_t need = sizeof(float4) * n + sizeof(uint32_t) * (2 * (size_t)T + 1 + 5 * n) + 256;
    PCR_TRY(ensure(ctx, ctx->b_table, need));  // (b_misc2 is the scan's own scratch)
    char *p = (char *)ctx->b_table.p;
    float4 *cpts = (float4 *)p;
    p += sizeof(float4) * n;
    uint32_t *rep = (uint32_t *)p;
    p += sizeof(uint32_t) * T;
    uint32_t *count = (uint32_t *)p;
    p += sizeof(uint32_t) * ((size_t)T + 1);
    uint32_t *slot_of = (uint32_t *)p;
    p += sizeof(uint32_t) * n;
    uint32_t *rank_of = (uint32_t *)p;
    p += sizeof(uint32_t) * n;
    uint32_t *cslot = (uint32_t *)p;
    p += sizeof(uint32_t) * n;
    uint32_t *parent = (uint32_t *)p;
    p += sizeof(uint32_t) * n;
    uint32_t *min_idx = (uint32_t *)p;
    PCR_CUDA(ctx, cudaMemsetAsync(rep, 0xff, sizeof(uint32_t) * T, st));
    PCR_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(uint32_t) * ((size_t)T + 1), st));
    const unsigned nb = (unsigned)((n + 255) / 256);
    cluster_init_parent_kernel<<<nb, 256, 0, st>>>(parent, min_idx, n);
    PCR_LAUNCH_CHECK(ctx);
    cluster_insert_kernel<<<nb, 256, 0, st>>>(dx, dy, dz, n, inv_r, rep, T - 1, count, slot_of, rank_of);
    PCR_LAUNCH_CHECK(ctx);
    PCR_TRY(exclusive_scan_u32_dev(ctx, count, (size_t)T + 1));
    cluster_scatter_kernel<<<nb, 256, 0, st>>>(dx, dy, dz, n, count, slot_of, rank_of, cpts, cslot);
    PCR_LAUNCH_CHECK(ctx);
    cluster_union_kernel<<<(unsigned)((n + 32 * kUnionWarps - 1) / (32 * kUnionWarps)), 32 * kUnionWarps, 0, st>>>(cpts, cslot, (uint32_t)n, dx, dy, dz, rep, T - 1, count, inv_r, r2,
                                                                       parent);
    PCR_LAUNCH_CHECK(ctx);
    cluster_roots_kernel<<<nb, 256, 0, st>>>(parent, cpts, (uint32_t)n, count, T - 1, min_idx);
    PCR_LAUNCH_CHECK(ctx);
    cluster_labels_kernel<<<nb, 256, 0, st>>>(parent, min_idx, count, slot_of, rank_of, n, d_labels);
    PCR_LAUNCH_CHECK(ctx);
    return PCR_OK;
}

// Labels (root = smallest index of the component) -> the reference's output (:164-186): clusters
// with min_size <= size <= max_size, size descending then first index ascending, indices ascending.
size_t clusters_from_labels(const uint32_t *labels, size_t n, size_t min_size, size_t max_size, uint32_t *offsets, uint32_t *indices,
                            std::vector<uint32_t> &scratch) {
    scratch.assign(2 * n, 0);
    uint32_t *size = scratch.data(), *slot = scratch.data() + n;
    for (size_t i = 0; i < n; i++) size[labels[i]]++;
    std::vector<std::pair<uint32_t, uint32_t>> comps;  // (size, root); root == first index
    for (size_t r = 0; r < n; r++)
        if (size[r] && size[r] >= min_size && size[r] <= max_size) comps.emplace_back(size[r], (uint32_t)r);
    std::sort(comps.begin(), comps.end(), [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) {
        return a.first != b.first ? a.first > b.first : a.second < b.second;
    });
    for (size_t r = 0; r < n; r++) slot[r] = 0xffffffffu;
    offsets[0] = 0;
    for (size_t c = 0; c < comps.size(); c++) {
        slot[comps[c].second] = offsets[c];
        offsets[c + 1] = offsets[c] + comps[c].first;
    }
    for (size_t i = 0; i < n; i++)
        if (slot[labels[i]] != 0xffffffffu) indices[slot[labels[i]]++] = (uint32_t)i;
    return comps.size();
}

}  // namespace pcr
