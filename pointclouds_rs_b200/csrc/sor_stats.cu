// sor_stats.cu -- global statistics and keep-mask of statistical_outlier_removal
// (crates/filters/src/statistical_outlier.rs:43-66), one segment per frame.
//
//   mean  = (sequential f32 sum of the finite mean distances) / n_finite
//   var   = (sequential f32 sum of (d - mean)^2) / n_finite      (population, two-pass)
//   thr   = mean + std_mul * sqrt(var)
//   keep  = mean_d <= thr
#include "seq_fold.cuh"

#include <algorithm>

namespace pcr {

namespace {

struct SorFrameStats {
    float mean, stddev, thr;
    uint32_t n_finite;
};

static __device__ __noinline__ float slow_fold(const float *__restrict__ v, size_t b, size_t e, FoldTerm term, FoldShared &sh, uint32_t *n_used) {
    return cluster_exact_fold(v, b, e, term, sh, n_used);
}

// one block CLUSTER per frame: two exact left-to-right folds (seq_fold.cuh), then the threshold
__global__ void __launch_bounds__(kFoldThreads) sor_stats_kernel(const float *__restrict__ mean_d,
                                                                  const uint32_t *__restrict__ frame_off, size_t n,
                                                                  float std_mul, SorFrameStats *__restrict__ stats, int use_fast) {
    PCR_GRID_DEP_SYNC();
    __shared__ FoldShared sh;
    __shared__ SpecShared sp;
    const int f = blockIdx.x / cooperative_groups::this_cluster().num_blocks();
    const size_t b = frame_off ? frame_off[f] : 0, e = frame_off ? frame_off[f + 1] : n;
    uint32_t nf = 0;
    // (fast: crossings predicted from an f64 prefix and verified; the pass-per-crossing fold is the fallback -- same bits)
    float sum;
    FoldTerm term = {0, 0.0f};
    if (!use_fast || !cluster_exact_fold_fast(mean_d, b, e, term, sp, &sum, &nf, use_fast == 2)) sum = slow_fold(mean_d, b, e, term, sh, &nf);
    SorFrameStats s;
    s.n_finite = nf;
    if (nf == 0) {  // statistical_outlier.rs:49-51 -> empty result (NaN threshold keeps nothing)
        s.mean = s.stddev = s.thr = __int_as_float(0x7fc00000);
    } else {
        const float nn = (float)nf;
        const float gmean = __fdiv_rn(sum, nn);
        term.squared_deviation = 1;
        term.mean = gmean;
        float var;
        if (!use_fast || !cluster_exact_fold_fast(mean_d, b, e, term, sp, &var, nullptr, use_fast == 2)) var = slow_fold(mean_d, b, e, term, sh, nullptr);
        var = __fdiv_rn(var, nn);
        const float sd = __fsqrt_rn(var);
        s.mean = gmean;
        s.stddev = sd;
        s.thr = __fadd_rn(gmean, __fmul_rn(std_mul, sd));
    }
    if (threadIdx.x == 0 && cooperative_groups::this_cluster().block_rank() == 0) stats[f] = s;
}

__global__ void __launch_bounds__(256) sor_mask_kernel(const float *__restrict__ mean_d, const uint32_t *__restrict__ frame_off,
                                                       int n_frames, size_t n, const SorFrameStats *__restrict__ stats,
                                                       uint8_t *__restrict__ keep, unsigned long long *__restrict__ kept) {
    PCR_GRID_DEP_SYNC();
    const int f = blockIdx.y;
    const size_t b = frame_off ? frame_off[f] : 0, e = frame_off ? frame_off[f + 1] : n;
    const SorFrameStats s = stats[f];
    unsigned cnt = 0;
    for (size_t i = b + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (size_t)gridDim.x * blockDim.x) {
        // n_finite == 0 -> thr is NaN -> nothing is kept
        uint8_t k = mean_d[i] <= s.thr ? 1 : 0;
        keep[i] = k;
        cnt += k;
    }
    cnt = __reduce_add_sync(PCR_FULL, cnt);
    __shared__ unsigned sh;
    if (threadIdx.x == 0) sh = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&sh, cnt);
    __syncthreads();
    if (threadIdx.x == 0 && sh) atomicAdd(&kept[f], (unsigned long long)sh);
}

}  // namespace

int sor_threshold_mask_dev(Ctx *ctx, const float *d_mean_d, const uint32_t *d_frame_off, int n_frames, size_t n, float std_mul,
                           uint8_t *d_keep, float *d_stats, unsigned long long *d_kept, bool kept_zeroed) {
    static_assert(sizeof(SorFrameStats) == 4 * sizeof(float), "stats layout");
    if (n == 0) return PCR_OK;
    TimeScope ts(ctx, kTagSorStats);
    if (!kept_zeroed) PCR_CUDA(ctx, cudaMemsetAsync(d_kept, 0, sizeof(unsigned long long) * n_frames, ctx->stream));  // (before the kernels: they chain)
    {
        // a cluster of 8 CTAs per frame (one CTA for small frames: fewer cluster barriers)
        // 16 CTAs per frame for long frames (non-portable cluster size: fewer cluster-wide passes per fold)
        static const unsigned big = getenv("PCR_FOLD_CLUSTER") ? (unsigned)atoi(getenv("PCR_FOLD_CLUSTER")) : 16u;
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(sor_stats_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaGetLastError();
            attr_set = true;
        }
        const size_t per_frame = n / (size_t)n_frames;
        // (many frames already fill the GPU: 16-CTA clusters then only cost scheduling freedom -- 0.84 -> 1.67 ms on 100 frames)
        const unsigned cs = (per_frame > 12 * (size_t)kFoldTile && n_frames <= 4) ? big : (per_frame > 4 * (size_t)kFoldTile ? 8u : 1u);
        unsigned use = cs;
        for (;;) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)n_frames * use);
            cfg.blockDim = dim3(kFoldThreads);
            cfg.stream = ctx->stream;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = use;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // (see launch_chained)
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            static const bool no_pdl = getenv("PCR_NO_PDL") != nullptr;
            cfg.attrs = attr;
            cfg.numAttrs = no_pdl ? 1 : 2;
            static const int use_fast = getenv("PCR_FOLD_SLOW") ? 0 : (getenv("PCR_FOLD_DEBUG") ? 2 : 1);  // A/B hook: the pass-per-crossing fold only
            // (many frames: the GPU is full of independent clusters and the pass-per-crossing fold's smaller code wins -- 12.55 vs
            // 12.9 ms on the 100-frame batch; the predicted-crossing fold is for the single long frame, 85 -> 60 us)
            const int fast = n_frames <= 4 ? use_fast : 0;
            cudaError_t e = cudaLaunchKernelEx(&cfg, sor_stats_kernel, d_mean_d, d_frame_off, n, std_mul, (SorFrameStats *)d_stats, fast);
            if (e == cudaSuccess) break;
            cudaGetLastError();
            if (use <= 8) return fail(ctx, PCR_ERR_CUDA, "sor_stats_kernel launch failed: %s", cudaGetErrorString(e));
            use = 8;  // the non-portable 16-CTA cluster could not be scheduled: the portable size always can
        }
        ctx->launches++;
    }
    size_t per = n / (size_t)n_frames + 1;
    unsigned bx = (unsigned)std::min<size_t>((per + 255) / 256, (size_t)ctx->sm_count * 8);
    PCR_CUDA(ctx, launch_chained(sor_mask_kernel, dim3(bx ? bx : 1, n_frames), dim3(256), 0, ctx->stream, d_mean_d, d_frame_off, n_frames, n,
                                 (const SorFrameStats *)d_stats, d_keep, d_kept));
    ctx->launches++;
    return PCR_OK;
}

}  // namespace pcr
