// brn.cuh -- Batch ReNorm finalize (network/slim/ops.py:130-171): per-channel sums -> mean/var, r/d clip, affine a,b, optional
// UPDATE_OPS (zero-debiased EMA of the moving statistics, r_max/d_max/curr_t schedule).  Called by NTHREADS threads that
// share named barrier BAR_ID: the stand-alone finalize kernel, the last block of the stats kernel, or the last CTA's
// epilogue warps of the tcgen05 conv kernel (fused statistics).
#pragma once
#include "common.cuh"
#include <math.h>

// state layout per BRN conv: mov_mean[C], mov_var[C], biased_mean[C], biased_var[C], r_max, d_max, curr_t, local_step
template <int BAR_ID, int NTHREADS>
__device__ __forceinline__ void brn_finalize_dev(int tid, int C, double n, const double* __restrict__ sums, const float* __restrict__ bg,
                                 float* __restrict__ state, float* __restrict__ aff, float* __restrict__ bstat, int update_state) {
  const float eps = 0.001f, one_minus_decay = 0.01f;           // um_v1.py:9-10 (decay 0.99, epsilon 1e-3)
  const float r_max = state[4 * C], d_max = state[4 * C + 1], t = state[4 * C + 2], step = state[4 * C + 3];
  asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(NTHREADS) : "memory");
  for (int c = tid; c < C; c += NTHREADS) {
    double mean_d = __ldcg(sums + c) / n;
    double var_d = __ldcg(sums + C + c) / n - mean_d * mean_d;
    if (var_d < 0) var_d = 0;
    float mean = (float)mean_d, var = (float)var_d;
    float mov_mean = state[c], mov_var = state[C + c];
    float stdv = sqrtf(var + eps), mov_std = sqrtf(mov_var + eps);
    float r = fminf(fmaxf(stdv / mov_std, 1.0f / r_max), r_max);               // ops.py:158-159
    float d = fminf(fmaxf((mean - mov_mean) / mov_std, -d_max), d_max);       // ops.py:161-162
    float inv_std = 1.0f / sqrtf(var + eps);
    float beta = bg[c], gamma = bg[C + c];
    // y = ((x-mean)*inv_std*r + d)*gamma + beta = x*a + b
    float a = inv_std * r * gamma;
    float b = (d - mean * inv_std * r) * gamma + beta;
    aff[c] = a; aff[C + c] = b;
    bstat[c] = mean; bstat[C + c] = inv_std; bstat[2 * C + c] = r; bstat[3 * C + c] = d;
    if (update_state) {                                                       // ops.py:134-137, zero-debiased EMA
      float bm = state[2 * C + c], bv = state[3 * C + c];
      bm -= (bm - mean) * one_minus_decay;
      bv -= (bv - var) * one_minus_decay;
      float corr = 1.0f - powf(0.99f, step + 1.0f);
      state[2 * C + c] = bm; state[3 * C + c] = bv;
      state[c] = bm / corr; state[C + c] = bv / corr;
    }
  }
  asm volatile("bar.sync %0, %1;" ::"n"(BAR_ID), "n"(NTHREADS) : "memory");
  if (update_state && tid == 0) {
    state[4 * C] = 3.0f / (1.0f + 2.0f * expf(-t));                          // ops.py:141-144
    state[4 * C + 1] = 5.0f / (5000.0f * expf(-2.0f * t));                   // ops.py:146-149
    state[4 * C + 2] = t + 1e-5f;                                             // ops.py:151-153
    state[4 * C + 3] = step + 1.0f;
  }
}

