// seq_fold.cuh -- sequential-order f32 folds.
//
// The reference sums with `iter().sum::<f32>()`: a strictly left-to-right f32 fold
// (statistical_outlier.rs:54-59, icp.rs:277-280).  Its rounding error at N = 1e5..1e6 is ~1e-5..1e-4
// relative, i.e. larger than the band the SOR mask is allowed to differ in, so the fold ORDER is
// part of the result and has to be reproduced, not just the mathematical sum.
#pragma once
#include "pcr_internal.cuh"

namespace pcr {

// Plain left-to-right fold by one thread over v[b, e): sum of the elements for which
// `pred(v)` holds, mapped through `fn`.  Loads are independent of the add chain, so the loop runs
// at one dependent FADD (4 cycles) per element.
template <class Fn>
__device__ __forceinline__ float seq_fold_thread(const float *__restrict__ v, size_t b, size_t e, Fn fn, uint32_t *n_used) {
    float s = 0.0f;
    uint32_t cnt = 0;
    size_t i = b;
    for (; i + 8 <= e; i += 8) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; j++) t[j] = v[i + j];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (isfinite(t[j])) {
                s = __fadd_rn(s, fn(t[j]));
                cnt++;
            }
        }
    }
    for (; i < e; i++) {
        float t = v[i];
        if (isfinite(t)) {
            s = __fadd_rn(s, fn(t));
            cnt++;
        }
    }
    if (n_used) *n_used = cnt;
    return s;
}

}  // namespace pcr
