// ew.cu -- HBM-bound elementwise / reduction kernels around the convolutions.  All tensors are NHWC
// fp32 "views": a pointer already offset to the first channel plus a channel stride, so producers
// write straight into concat buffers (tf.concat on the path costs no copy).
//
// Reference ops replaced (under /root/reference):
//   data/preprocess.py:176-187 norm_dm                              -> norm_dm_kernel
//   network/um_v1.py:109-121 tiny_dm / uu / vv / uvd                -> make_uvd_kernel
//   network/slim/ops.py:640-669 max_pool (+ TF MaxPoolGrad)         -> maxpool_kernel / maxpool_bwd_kernel
//   network/slim/ops.py:671-677 upsampling_nearest + add um_v1.py:69 -> upadd_kernel / upadd_bwd_lo_kernel
//   network/um_v1.py:146-148 tf.where depth mask                    -> copy_view_kernel(mask)
//   network/slim/ops.py:130-171 Batch ReNorm (train)                -> channel_stats / brn_finalize / brn_apply
//   network/slim/ops.py:173-180 BRN eval                            -> fold_affine
//   model/hourglass_um_crop_tiny.py:195-274,323-371 GT synthesis + l2 losses -> loss_kernel
//   network/slim/losses.py:56-72 l2_regularizer                     -> wd_kernel
//   model/train_single_gpu.py:86-88 + tf.train.AdamOptimizer        -> adam_kernel
#include "common.cuh"
#include "brn.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

constexpr int EW_T = 256;
inline int blocks_for(size_t n, int per_block = EW_T, int cap = 148 * 16) {
  size_t b = (n + per_block - 1) / per_block;
  if (b > (size_t)cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

__global__ void norm_dm_kernel(int B, int npix, const float* __restrict__ dm, const float* __restrict__ coms,
                               float* __restrict__ out) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  size_t n = (size_t)B * npix;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int b = (int)(i / npix);
    float cz = coms[b * 3 + 2];
    float max_depth = cz + 300.0f * 0.5f, min_depth = cz - 300.0f * 0.5f;
    float d = dm[i];
    bool m = (d < max_depth) && (d > (min_depth - 300.0f * 0.5f));
    out[i] = m ? __fdiv_rn(__fsub_rn(d, min_depth), 300.0f) : -1.0f;
  }
}

__global__ void make_uvd_kernel(int B, int in_hw, int out_hw, const float* __restrict__ x0, float* __restrict__ tiny,
                                UvdDst dst) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  size_t n = (size_t)B * out_hw * out_hw;
  int s = in_hw / out_hw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int b = (int)(i / (out_hw * out_hw));
    int r = (int)(i - (size_t)b * out_hw * out_hw);
    int y = r / out_hw, x = r - y * out_hw;
    float d = x0[((size_t)b * in_hw + y * s) * in_hw + x * s];
    float uu = (float)x / (float)(out_hw / 2) - 1.0f;
    float vv = (float)y / (float)(out_hw / 2) - 1.0f;
    if (tiny) tiny[i] = d;
    for (int k = 0; k < dst.n; ++k) {
      float* p = dst.p[k] + i * dst.cs[k];
      p[0] = uu; p[1] = vv; p[2] = d;
    }
  }
}

// SAME max pool stride 2: k=2 (pad 0,0) or k=3 (pad 0 before, 1 after on even inputs); padding never wins
__global__ void maxpool_kernel(int B, int H, int W, int C, int k, const float* __restrict__ x, int x_cs,
                               float* __restrict__ y, int y_cs) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  int Ho = H / 2, Wo = W / 2;
  size_t n = (size_t)B * Ho * Wo * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    int ox = (int)(pix % Wo); int oy = (int)((pix / Wo) % Ho); int b = (int)(pix / ((size_t)Wo * Ho));
    float m = -INFINITY;
    for (int dy = 0; dy < k; ++dy) {
      int iy = oy * 2 + dy; if (iy >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        int ix = ox * 2 + dx; if (ix >= W) continue;
        float v = x[((size_t)(b * H + iy) * W + ix) * x_cs + c];
        m = v > m ? v : m;
      }
    }
    y[pix * y_cs + c] = m;
  }
}

// gather form of MaxPoolGrad: input pixel (iy,ix) receives dy of every window whose FIRST maximum (row-major
// scan, strict >) it is -- the convention of TF's and PyTorch's max-pool backward.
// float4-over-channels variant of maxpool_bwd_kernel (C, strides % 4 == 0, 16 B aligned views): same gather formulation and the same
// "first maximum in (dy, dx) scan order wins" tie rule per channel, one thread per (input pixel, channel quad) -- a quarter of the load
// instructions (the scalar kernel issues up to 36 loads of x per element on the 3x3/s2 pools: 192 us for the 32x32 -> 16x16 pool at B=40).
// Default since round 2 (verified against the fp32 engine and measured: profiles/r2_sweep.md); DENSEREG_POOL_BWD_V4=0 selects the scalar kernel.
__global__ void maxpool_bwd_v4_kernel(int B, int H, int W, int C4, int k, const float4* __restrict__ x, int x_cs4,
                                      const float4* __restrict__ dy, int dy_cs4, float4* __restrict__ dx, int dx_cs4, int accumulate) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  const int Ho = H / 2, Wo = W / 2;
  const size_t n = (size_t)B * H * W * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4); const size_t pix = i / C4;
    const int ix = (int)(pix % W); const int iy = (int)((pix / W) % H); const int b = (int)(pix / ((size_t)W * H));
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    int oy_lo = (iy - (k - 1) + 1) / 2; if (iy - (k - 1) < 0) oy_lo = 0;
    int ox_lo = (ix - (k - 1) + 1) / 2; if (ix - (k - 1) < 0) ox_lo = 0;
    for (int oy = oy_lo; oy <= iy / 2 && oy < Ho; ++oy) {
      for (int ox = ox_lo; ox <= ix / 2 && ox < Wo; ++ox) {
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        bool mine[4] = {false, false, false, false};               // is (iy, ix) the arg-max of this window, per channel
        for (int ddy = 0; ddy < k; ++ddy) {
          const int yy = oy * 2 + ddy; if (yy >= H) continue;
          for (int ddx = 0; ddx < k; ++ddx) {
            const int xx = ox * 2 + ddx; if (xx >= W) continue;
            const float4 v4 = x[((size_t)(b * H + yy) * W + xx) * x_cs4 + c];
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
            const bool here = (yy == iy && xx == ix);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (v[e] > m[e]) { m[e] = v[e]; mine[e] = here; }
          }
        }
        const float4 d4 = dy[((size_t)(b * Ho + oy) * Wo + ox) * dy_cs4 + c];
        if (mine[0]) g[0] += d4.x;
        if (mine[1]) g[1] += d4.y;
        if (mine[2]) g[2] += d4.z;
        if (mine[3]) g[3] += d4.w;
      }
    }
    float4* o = dx + pix * dx_cs4 + c;
    float4 r = make_float4(g[0], g[1], g[2], g[3]);
    if (accumulate) { const float4 old = *o; r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
    *o = r;
  }
}

__global__ void maxpool_bwd_kernel(int B, int H, int W, int C, int k, const float* __restrict__ x, int x_cs,
                                   const float* __restrict__ dy, int dy_cs, float* __restrict__ dx, int dx_cs, int accumulate) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  int Ho = H / 2, Wo = W / 2;
  size_t n = (size_t)B * H * W * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    int ix = (int)(pix % W); int iy = (int)((pix / W) % H); int b = (int)(pix / ((size_t)W * H));
    float g = 0.f;
    // windows (oy,ox) covering (iy,ix): oy*2 <= iy <= oy*2+k-1
    int oy_lo = (iy - (k - 1) + 1) / 2; if (iy - (k - 1) < 0) oy_lo = 0;
    int ox_lo = (ix - (k - 1) + 1) / 2; if (ix - (k - 1) < 0) ox_lo = 0;
    for (int oy = oy_lo; oy <= iy / 2 && oy < Ho; ++oy) {
      for (int ox = ox_lo; ox <= ix / 2 && ox < Wo; ++ox) {
        float m = -INFINITY; int ay = -1, ax = -1;
        for (int ddy = 0; ddy < k; ++ddy) {
          int yy = oy * 2 + ddy; if (yy >= H) continue;
          for (int ddx = 0; ddx < k; ++ddx) {
            int xx = ox * 2 + ddx; if (xx >= W) continue;
            float v = x[((size_t)(b * H + yy) * W + xx) * x_cs + c];
            if (v > m) { m = v; ay = yy; ax = xx; }
          }
        }
        if (ay == iy && ax == ix) g += dy[((size_t)(b * Ho + oy) * Wo + ox) * dy_cs + c];
      }
    }
    float* o = dx + pix * dx_cs + c;
    *o = accumulate ? *o + g : g;
  }
}

__global__ void upadd_kernel(int B, int H, int W, int C, const float* __restrict__ a, int a_cs,
                             const float* __restrict__ lo, int lo_cs, float* __restrict__ y, int y_cs) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  size_t n = (size_t)B * H * W * C;
  int Hl = H / 2, Wl = W / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    int x = (int)(pix % W); int yy = (int)((pix / W) % H); int b = (int)(pix / ((size_t)W * H));
    float v = a[pix * a_cs + c] + lo[((size_t)(b * Hl + (yy >> 1)) * Wl + (x >> 1)) * lo_cs + c];
    y[pix * y_cs + c] = v;
  }
}

__global__ void upadd_bwd_lo_kernel(int B, int H, int W, int C, const float* __restrict__ dy, int dy_cs,
                                    float* __restrict__ dlo, int dlo_cs, int accumulate) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  int Hl = H / 2, Wl = W / 2;
  size_t n = (size_t)B * Hl * Wl * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    int x = (int)(pix % Wl); int y = (int)((pix / Wl) % Hl); int b = (int)(pix / ((size_t)Wl * Hl));
    const float* r0 = dy + ((size_t)(b * H + 2 * y) * W + 2 * x) * dy_cs + c;
    const float* r1 = r0 + (size_t)W * dy_cs;
    float g = (r0[0] + r0[dy_cs]) + (r1[0] + r1[dy_cs]);
    float* o = dlo + pix * dlo_cs + c;
    *o = accumulate ? *o + g : g;
  }
}

__global__ void copy_view_kernel(size_t npix, int C, const float* __restrict__ src, int src_cs, float* __restrict__ dst,
                                 int dst_cs, int accumulate, const float* __restrict__ tiny_mask) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  size_t n = npix * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    float v = src[pix * src_cs + c];
    if (tiny_mask && tiny_mask[pix] < -0.9f) v = 0.f;
    float* o = dst + pix * dst_cs + c;
    *o = accumulate ? *o + v : v;
  }
}

__global__ void fill_view_kernel(size_t npix, int C, float* __restrict__ dst, int dst_cs, float v) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  size_t n = npix * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    dst[pix * dst_cs + c] = v;
  }
}



// per-channel sum / sum of squares in double + BRN finalize by the LAST block to finish (one launch instead of two)
__global__ void channel_stats_finalize_kernel(size_t npix, int C, const float* __restrict__ x, int x_cs, double* __restrict__ sums,
                                              unsigned int* __restrict__ counter, const float* __restrict__ bg, float* __restrict__ state,
                                              float* __restrict__ aff, float* __restrict__ bstat, int update_state) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  __shared__ double s1[8][33], s2[8][33];
  __shared__ int is_last;
  int c = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0, b = 0.0;
  if (c < C) {
    for (size_t p = blockIdx.y * 8 + threadIdx.y; p < npix; p += (size_t)gridDim.y * 8) {
      float v = x[p * x_cs + c];
      a += (double)v; b += (double)v * (double)v;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    atomicAdd(sums + c, a); atomicAdd(sums + C + c, b);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    brn_finalize_dev<0, 256>(threadIdx.y * 32 + threadIdx.x, C, (double)npix, sums, bg, state, aff, bstat, update_state);
  }
}


__global__ void brn_apply_kernel(size_t npix, int C, const float* __restrict__ raw, int raw_cs, const float* __restrict__ aff,
                                 int relu, const float* __restrict__ res, int res_cs, float* __restrict__ y, int y_cs) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  size_t n = npix * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    float v = raw[pix * raw_cs + c] * aff[c] + aff[C + c];
    if (relu) v = fmaxf(v, 0.f);
    if (res) v += res[pix * res_cs + c];
    y[pix * y_cs + c] = v;
  }
}

// float4 variant: C % 4 == 0, all strides % 4 == 0, all bases 16 B aligned; 32-bit index math
__global__ void brn_apply_v4_kernel(unsigned n4, unsigned C4, const float4* __restrict__ raw, unsigned raw_cs4,
                                    const float4* __restrict__ aff, int relu, const float4* __restrict__ res, unsigned res_cs4,
                                    float4* __restrict__ y, unsigned y_cs4, int rev) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += gridDim.x * blockDim.x) {
    const unsigned i = rev ? n4 - 1 - i0 : i0;      // rev: walk the map from its END (ew_reverse() below)
    const unsigned pix = i / C4, c = i - pix * C4;
    const float4 x = raw[(size_t)pix * raw_cs4 + c], a = __ldg(aff + c), b = __ldg(aff + C4 + c);
    float4 v = make_float4(fmaf(x.x, a.x, b.x), fmaf(x.y, a.y, b.y), fmaf(x.z, a.z, b.z), fmaf(x.w, a.w, b.w));
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (res) { const float4 r = res[(size_t)pix * res_cs4 + c]; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
    y[(size_t)pix * y_cs4 + c] = v;
  }
}

__global__ void brn_bwd_reduce_kernel(size_t npix, int C, const float* __restrict__ dy, int dy_cs, const float* __restrict__ raw,
                                      int raw_cs, const float* __restrict__ aff, const float* __restrict__ bstat, int relu,
                                      double* __restrict__ sums) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  __shared__ double s1[8][33], s2[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0, b = 0.0;
  if (c < C) {
    float sa = aff[c], sb = aff[C + c], mean = bstat[c], inv_std = bstat[C + c];
    for (size_t p = blockIdx.y * 8 + threadIdx.y; p < npix; p += (size_t)gridDim.y * 8) {
      float x = raw[p * raw_cs + c];
      float g = dy[p * dy_cs + c];
      if (relu && !(x * sa + sb > 0.f)) g = 0.f;
      float xh = (x - mean) * inv_std;
      a += (double)g; b += (double)g * (double)xh;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { a += s1[i][threadIdx.x]; b += s2[i][threadIdx.x]; }
    atomicAdd(sums + c, a); atomicAdd(sums + C + c, b);
  }
}

__global__ void brn_bwd_apply_kernel(size_t npix, int C, const float* __restrict__ dy, int dy_cs, const float* __restrict__ raw,
                                     int raw_cs, const float* __restrict__ aff, const float* __restrict__ bstat,
                                     const float* __restrict__ bg, int relu, const double* __restrict__ sums,
                                     float* __restrict__ draw, int draw_cs, float* __restrict__ gparam) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  size_t n = npix * C;
  const double inv_n = 1.0 / (double)npix;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    float sa = aff[c], sb = aff[C + c], mean = bstat[c], inv_std = bstat[C + c], r = bstat[2 * C + c];
    float gamma = bg[C + c];
    float x = raw[pix * raw_cs + c];
    float g = dy[pix * dy_cs + c];
    if (relu && !(x * sa + sb > 0.f)) g = 0.f;
    float xh = (x - mean) * inv_std;
    float mg = (float)(sums[c] * inv_n), mgx = (float)(sums[C + c] * inv_n);
    draw[pix * draw_cs + c] = gamma * r * inv_std * (g - mg - xh * mgx);
  }
  // dbeta = sum g ; dgamma = sum g*(xhat*r + d) = r*sum_gx + d*sum_g      (block 0 only)
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float r = bstat[2 * C + c], d = bstat[3 * C + c];
      gparam[c] += (float)sums[c];
      gparam[C + c] += (float)((double)r * sums[C + c] + (double)d * sums[c]);
    }
  }
}

// ---- BRN backward, channel-stationary float4 kernels (C % 4 == 0, 16 B aligned views) -----------------------------------------------
// Thread mapping for both: blockDim.x = C4 * lanes (C4 = C/4 channel quads, lanes = floor(256 / C4) pixels side by side); a thread keeps ONE
// channel quad for its whole life, so the per-channel constants are loaded once into registers instead of 6 scalar loads + 2 double
// loads per element, and walks pixels with a 4-way unrolled batch of independent 16 B loads (the first version was load-instruction
// bound: 34 load instructions per 3 data accesses, ~2 TB/s).  A warp still reads whole contiguous pixel rows (coalesced).
struct BrnQuad { float sa[4], sb[4], mean[4], istd[4]; };
DR_DEVINL void brn_quad_load(BrnQuad& q, const float* __restrict__ aff, const float* __restrict__ bstat, int C, int c) {
#pragma unroll
  for (int e = 0; e < 4; ++e) { q.sa[e] = __ldg(aff + c + e); q.sb[e] = __ldg(aff + C + c + e); q.mean[e] = __ldg(bstat + c + e); q.istd[e] = __ldg(bstat + C + c + e); }
}

// sums[0:C] += sum g, sums[C:2C] += sum g*xhat  (g = dy * [z > 0]); fp32 partial sums over 4 pixels, then double
__global__ void brn_bwd_reduce_v4_kernel(unsigned npix, unsigned C4, const float4* __restrict__ dy, unsigned dy_cs4,
                                         const float4* __restrict__ raw, unsigned raw_cs4, const float* __restrict__ aff,
                                         const float* __restrict__ bstat, int relu, double* __restrict__ sums, int rev) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  extern __shared__ double red[];                       // [2][blockDim.x][4]
  const unsigned lanes = blockDim.x / C4, cq = threadIdx.x % C4, pl = threadIdx.x / C4;
  const int C = (int)C4 * 4, c = (int)cq * 4;
  BrnQuad q; brn_quad_load(q, aff, bstat, C, c);
  double A[4] = {0.0, 0.0, 0.0, 0.0}, Bs[4] = {0.0, 0.0, 0.0, 0.0};
  const unsigned stride = gridDim.x * lanes;
  for (unsigned p0 = blockIdx.x * lanes + pl; p0 < npix; p0 += 4 * stride) {
    float4 x4[4], g4[4]; bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned pp = p0 + u * stride; ok[u] = pp < npix;
      unsigned pc = ok[u] ? pp : p0;
      if (rev) pc = npix - 1 - pc;                      // walk the map from its end (ew_reverse())
      x4[u] = raw[(size_t)pc * raw_cs4 + cq]; g4[u] = dy[(size_t)pc * dy_cs4 + cq];
    }
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const float xs[4] = {x4[u].x, x4[u].y, x4[u].z, x4[u].w}, gs[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float g = gs[e];
        if (relu && !(xs[e] * q.sa[e] + q.sb[e] > 0.f)) g = 0.f;
        const float xh = (xs[e] - q.mean[e]) * q.istd[e];
        a[e] += g; b[e] += g * xh;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { A[e] += (double)a[e]; Bs[e] += (double)b[e]; }
  }
  double* r0 = red; double* r1 = red + (size_t)blockDim.x * 4;
#pragma unroll
  for (int e = 0; e < 4; ++e) { r0[threadIdx.x * 4 + e] = A[e]; r1[threadIdx.x * 4 + e] = Bs[e]; }
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double sa = 0.0, sb = 0.0;
      for (unsigned l = 0; l < lanes; ++l) { sa += r0[(l * C4 + cq) * 4 + e]; sb += r1[(l * C4 + cq) * 4 + e]; }
      atomicAdd(sums + c + e, sa); atomicAdd(sums + C + c + e, sb);
    }
  }
}

// draw = gamma*r*inv_std*(g - sum_g/N - xhat*sum_gx/N); block 0 also accumulates dbeta, dgamma
__global__ void brn_bwd_apply_v4_kernel(unsigned npix, unsigned C4, const float4* __restrict__ dy, unsigned dy_cs4,
                                        const float4* __restrict__ raw, unsigned raw_cs4, const float* __restrict__ aff,
                                        const float* __restrict__ bstat, const float* __restrict__ bg, int relu,
                                        const double* __restrict__ sums, float4* __restrict__ draw, unsigned draw_cs4,
                                        float* __restrict__ gparam) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  const unsigned lanes = blockDim.x / C4, cq = threadIdx.x % C4, pl = threadIdx.x / C4;
  const int C = (int)C4 * 4, c = (int)cq * 4;
  const double inv_n = 1.0 / (double)npix;
  BrnQuad q; brn_quad_load(q, aff, bstat, C, c);
  float K[4], mg[4], mgx[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    K[e] = __ldg(bg + C + c + e) * __ldg(bstat + 2 * C + c + e) * q.istd[e];          // gamma * r * inv_std (same association as before)
    mg[e] = (float)(sums[c + e] * inv_n); mgx[e] = (float)(sums[C + c + e] * inv_n);
  }
  const unsigned stride = gridDim.x * lanes;
  for (unsigned p0 = blockIdx.x * lanes + pl; p0 < npix; p0 += 4 * stride) {
    float4 x4[4], g4[4]; bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned pp = p0 + u * stride; ok[u] = pp < npix;
      const unsigned pc = ok[u] ? pp : p0;
      x4[u] = raw[(size_t)pc * raw_cs4 + cq]; g4[u] = dy[(size_t)pc * dy_cs4 + cq];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const float xs[4] = {x4[u].x, x4[u].y, x4[u].z, x4[u].w}, gs[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float g = gs[e];
        if (relu && !(xs[e] * q.sa[e] + q.sb[e] > 0.f)) g = 0.f;
        const float xh = (xs[e] - q.mean[e]) * q.istd[e];
        o[e] = K[e] * (g - mg[e] - xh * mgx[e]);
      }
      draw[(size_t)(p0 + u * stride) * draw_cs4 + cq] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  if (blockIdx.x == 0) {
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
      float r = bstat[2 * C + cc], d = bstat[3 * C + cc];
      gparam[cc] += (float)sums[cc];
      gparam[C + cc] += (float)((double)r * sums[C + cc] + (double)d * sums[cc]);
    }
  }
}

// ---- BRN backward for SMALL layers: reduce + apply in ONE launch by one thread-block cluster --------------------------------------------
// The <= 8x8 hourglass levels (and everything at a small batch) are a few CTAs' worth of data per layer; with two kernels the layer costs two
// launch / latency floors on the dependency chain of the backward pass.  Here the 8 CTAs of one cluster each reduce their pixel slice
// (same channel-stationary float4 mapping as brn_bwd_reduce_v4), publish per-channel partial sums in shared memory, meet at the hardware cluster
// barrier, add up the 8 partials through distributed shared memory and apply (same arithmetic as brn_bwd_apply_v4) to their own slice, which
// is still in L1 / L2.  Results are identical to the two-kernel path up to the order of the double-precision partial sums.
constexpr int BRN_CL = 8;
DR_DEVINL unsigned brn_cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
DR_DEVINL void brn_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
DR_DEVINL double brn_ld_remote(const double* local, unsigned rank) {
  unsigned la = (unsigned)__cvta_generic_to_shared(local), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
__global__ void brn_bwd_cluster_kernel(unsigned npix, unsigned C4, const float4* __restrict__ dy, unsigned dy_cs4,
                                       const float4* __restrict__ raw, unsigned raw_cs4, const float* __restrict__ aff,
                                       const float* __restrict__ bstat, const float* __restrict__ bg, int relu,
                                       float4* __restrict__ draw, unsigned draw_cs4, float* __restrict__ gparam) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  extern __shared__ double red[];                       // [2][blockDim.x][4] block reduction, then part[2][C] (published), tot[2][C]
  const unsigned lanes = blockDim.x / C4, cq = threadIdx.x % C4, pl = threadIdx.x / C4;
  const int C = (int)C4 * 4, c = (int)cq * 4;
  const unsigned rank = brn_cluster_rank();
  double* part = red + (size_t)blockDim.x * 8;          // [2][C]
  double* tot = part + 2 * C;                           // [2][C]
  BrnQuad q; brn_quad_load(q, aff, bstat, C, c);
  double A[4] = {0.0, 0.0, 0.0, 0.0}, Bs[4] = {0.0, 0.0, 0.0, 0.0};
  const unsigned stride = BRN_CL * lanes;
  for (unsigned p0 = rank * lanes + pl; p0 < npix; p0 += 4 * stride) {
    float4 x4[4], g4[4]; bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned pp = p0 + u * stride; ok[u] = pp < npix;
      const unsigned pc = ok[u] ? pp : p0;
      x4[u] = raw[(size_t)pc * raw_cs4 + cq]; g4[u] = dy[(size_t)pc * dy_cs4 + cq];
    }
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const float xs[4] = {x4[u].x, x4[u].y, x4[u].z, x4[u].w}, gs[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float g = gs[e];
        if (relu && !(xs[e] * q.sa[e] + q.sb[e] > 0.f)) g = 0.f;
        const float xh = (xs[e] - q.mean[e]) * q.istd[e];
        a[e] += g; b[e] += g * xh;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { A[e] += (double)a[e]; Bs[e] += (double)b[e]; }
  }
  double* r0 = red; double* r1 = red + (size_t)blockDim.x * 4;
#pragma unroll
  for (int e = 0; e < 4; ++e) { r0[threadIdx.x * 4 + e] = A[e]; r1[threadIdx.x * 4 + e] = Bs[e]; }
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      double sa = 0.0, sb = 0.0;
      for (unsigned l = 0; l < lanes; ++l) { sa += r0[(l * C4 + cq) * 4 + e]; sb += r1[(l * C4 + cq) * 4 + e]; }
      part[c + e] = sa; part[C + c + e] = sb;
    }
  }
  brn_cluster_sync();                                   // every CTA's partials are published
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    double t = 0.0;
#pragma unroll
    for (unsigned r = 0; r < BRN_CL; ++r) t += brn_ld_remote(part + i, r);
    tot[i] = t;
  }
  brn_cluster_sync();                                   // nobody leaves (or reuses `part`) while a peer may still read it; also orders tot[] for this CTA
  const double inv_n = 1.0 / (double)npix;
  {
    float K[4], mg[4], mgx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      K[e] = __ldg(bg + C + c + e) * __ldg(bstat + 2 * C + c + e) * q.istd[e];
      mg[e] = (float)(tot[c + e] * inv_n); mgx[e] = (float)(tot[C + c + e] * inv_n);
    }
    for (unsigned p0 = rank * lanes + pl; p0 < npix; p0 += 4 * stride) {
      float4 x4[4], g4[4]; bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned pp = p0 + u * stride; ok[u] = pp < npix;
        const unsigned pc = ok[u] ? pp : p0;
        x4[u] = raw[(size_t)pc * raw_cs4 + cq]; g4[u] = dy[(size_t)pc * dy_cs4 + cq];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        const float xs[4] = {x4[u].x, x4[u].y, x4[u].z, x4[u].w}, gs[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float g = gs[e];
          if (relu && !(xs[e] * q.sa[e] + q.sb[e] > 0.f)) g = 0.f;
          const float xh = (xs[e] - q.mean[e]) * q.istd[e];
          o[e] = K[e] * (g - mg[e] - xh * mgx[e]);
        }
        draw[(size_t)(p0 + u * stride) * draw_cs4 + cq] = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  if (rank == 0) {
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
      const float r = bstat[2 * C + cc], d = bstat[3 * C + cc];
      gparam[cc] += (float)tot[cc];
      gparam[C + cc] += (float)((double)r * tot[C + cc] + (double)d * tot[cc]);
    }
  }
}

// float4 copy / accumulate of a view (optional depth mask)
__global__ void copy_view_v4_kernel(unsigned n4, unsigned C4, const float4* __restrict__ src, unsigned src_cs4, float4* __restrict__ dst,
                                    unsigned dst_cs4, int accumulate, const float* __restrict__ tiny_mask) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const unsigned pix = i / C4, c = i - pix * C4;
    float4 v = src[(size_t)pix * src_cs4 + c];
    if (tiny_mask && tiny_mask[pix] < -0.9f) v = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* o = dst + (size_t)pix * dst_cs4 + c;
    if (accumulate) { const float4 w = *o; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    *o = v;
  }
}

__global__ void bias_bwd_kernel(size_t npix, int C, const float* __restrict__ dy, int dy_cs, const float* __restrict__ out, int out_cs,
                                int relu, int dropout, float* __restrict__ dz, int dz_cs, float* __restrict__ gbias) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  __shared__ float s1[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (c < C) {
    for (size_t p = blockIdx.y * 8 + threadIdx.y; p < npix; p += (size_t)gridDim.y * 8) {
      float g = dy[p * dy_cs + c];
      if (relu) g = out[p * out_cs + c] > 0.f ? (dropout ? g * 2.0f : g) : 0.f;
      dz[p * dz_cs + c] = g;
      a += g;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) a += s1[i][threadIdx.x];
    atomicAdd(gbias + c, a);
  }
}

// GT synthesis + 3 l2 losses + dL/d(out) for every stack; one thread per (b, pixel, joint): consecutive threads take consecutive joints of
// one pixel, so the hm / hm3 rows (J floats) and the um row (3J floats) are read and written contiguously.  (The first version looped over
// the joints inside one thread per pixel: 40960 threads for batch 40, 101 us for 52 MB of traffic.)
__global__ void loss_kernel(LossArgs a) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  const int hw = a.hw, J = a.J;
  const size_t n = (size_t)a.B * hw * hw * J;
  double l_hm = 0.0, l_hm3 = 0.0, l_um = 0.0;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / J; const int j = (int)(t - i * J);
    int b = (int)(i / (hw * hw)); int r = (int)(i - (size_t)b * hw * hw);
    int py = r / hw, px = r - py * hw;
    const float* cfg = a.cfgs + b * 6; const float* com = a.coms + b * 3;
    float w_ratio = cfg[4] / (float)hw, h_ratio = cfg[5] / (float)hw;
    float fx = cfg[0] / w_ratio, fy = cfg[1] / h_ratio, cx = cfg[2] / w_ratio, cy = cfg[3] / h_ratio;
    float d = a.tiny[i];
    float z = d < -0.99f ? com[2] + 150.0f : d * 300.0f + (com[2] - 150.0f);
    float X = ((float)px - cx) * (z / fx), Y = ((float)py - cy) * (z / fy);
    float Pn0 = (X - com[0]) / 100.0f, Pn1 = (Y - com[1]) / 100.0f, Pn2 = (z - com[2]) / 100.0f;
    {
      const float* pose = a.poses + (size_t)b * 3 * J + 3 * j;
      float o0 = (pose[0] - com[0]) / 100.0f - Pn0;
      float o1 = (pose[1] - com[1]) / 100.0f - Pn1;
      float o2 = (pose[2] - com[2]) / 100.0f - Pn2;
      float dist = sqrtf(o0 * o0 + o1 * o1 + o2 * o2);
      float g3 = fmaxf((0.8f - dist) / 0.8f, 0.f);                              // _hm_3d :206-208
      float dd = 0.8f - g3 * 0.8f;                                               // _um :260
      bool m = dd < (0.8f - 1e-2f);
      float u0 = m ? o0 / dd : 0.f, u1 = m ? o1 / dd : 0.f, u2 = m ? o2 / dd : 0.f;
      float uu = pose[0] * fx / pose[2] + cx, vv = pose[1] * fy / pose[2] + cy;  // util.py:20
      float e0 = (float)px - uu, e1 = (float)py - vv;
      float g2 = fmaxf(4.0f - sqrtf(e0 * e0 + e1 * e1), 0.f) / 4.0f;             // _hm_2d :243-244
      for (int s = 0; s < a.S; ++s) {
        size_t o = i * a.cs[s], go = i * a.gcs[s];
        float dh = a.hm[s][o + j] - g2;
        float dh3 = a.hm3[s][o + j] - g3;
        float q0 = a.um[s][o + 3 * j] - u0, q1 = a.um[s][o + 3 * j + 1] - u1, q2 = a.um[s][o + 3 * j + 2] - u2;
        a.ghm[s][go + j] = dh; a.ghm3[s][go + j] = dh3;
        a.gum[s][go + 3 * j] = q0; a.gum[s][go + 3 * j + 1] = q1; a.gum[s][go + 3 * j + 2] = q2;
        l_hm += 0.5 * (double)dh * dh; l_hm3 += 0.5 * (double)dh3 * dh3;
        l_um += 0.5 * ((double)q0 * q0 + (double)q1 * q1 + (double)q2 * q2);
      }
    }
  }
  __shared__ double red[3][EW_T / 32];
  for (int off = 16; off > 0; off >>= 1) {
    l_hm += __shfl_xor_sync(0xffffffffu, l_hm, off);
    l_hm3 += __shfl_xor_sync(0xffffffffu, l_hm3, off);
    l_um += __shfl_xor_sync(0xffffffffu, l_um, off);
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = l_hm; red[1][w] = l_hm3; red[2][w] = l_um; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int k = 0; k < EW_T / 32; ++k) s += red[threadIdx.x][k];
    atomicAdd(a.loss_acc + threadIdx.x, s);
  }
}

__global__ void wd_kernel(size_t n, const float* __restrict__ p, const float* __restrict__ wdm, float* __restrict__ g, double* __restrict__ reg) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float w = wdm[i];
    if (w != 0.f) { float v = p[i]; g[i] += w * v; acc += 0.5 * (double)w * v * v; }
  }
  __shared__ double red[EW_T / 32];
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0; for (int k = 0; k < EW_T / 32; ++k) s += red[k]; atomicAdd(reg, s); }
}

__global__ void finish_loss_kernel(const double* __restrict__ acc, float* __restrict__ out5) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  double t = acc[0] + acc[1] + acc[2] + acc[3];
  out5[0] = (float)t; out5[1] = (float)acc[0]; out5[2] = (float)acc[1]; out5[3] = (float)acc[2]; out5[4] = (float)acc[3];
}

// TF ApplyAdam form (oracle/um_v1_torch.py adam_step): m += (g - m)(1 - b1); v += (g*g - v)(1 - b2); p -= (m * alpha) / (sqrt(v) + eps),
// every operation a separate fp32 rounding (no FMA contraction), omb1 / omb2 = the fp32 differences 1.0f - beta.
__global__ void adam_kernel(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            float inv_scale, float clip, float alpha, float omb1, float omb2, float eps) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = __fdiv_rn(g[i], inv_scale);                  // inv_scale carries accum_steps * world: the mean is a division, as in the reference
    gi = fminf(fmaxf(gi, -clip), clip);
    const float mi = __fadd_rn(m[i], __fmul_rn(__fsub_rn(gi, m[i]), omb1));
    const float vi = __fadd_rn(v[i], __fmul_rn(__fsub_rn(__fmul_rn(gi, gi), v[i]), omb2));
    m[i] = mi; v[i] = vi;
    p[i] = __fsub_rn(p[i], __fdiv_rn(__fmul_rn(mi, alpha), __fadd_rn(__fsqrt_rn(vi), eps)));
  }
}


__global__ void init_trunc_normal_kernel(size_t n, float* __restrict__ p, float stddev, uint64_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // Box-Muller on a splitmix64 stream, re-drawn outside +-2 sigma (tf.truncated_normal_initializer, ops.py:272)
    uint64_t ctr = i * 16;
    float z = 0.f;
    for (int t = 0; t < 16; ++t) {
      uint64_t x = (ctr + t) + seed * 0x9E3779B97F4A7C15ull;
      x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
      float u1 = ((uint32_t)(x >> 40) + 1.0f) * (1.0f / 16777217.0f);
      float u2 = (uint32_t)((x >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
      z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
      if (fabsf(z) <= 2.0f) break;
      z = 0.f;
    }
    p[i] = z * stddev;
  }
}

__global__ void gather_outputs_kernel(size_t npix, int C, const float* __restrict__ src, int src_cs, float* __restrict__ dst) {
  size_t n = npix * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C); size_t pix = i / C;
    dst[i] = src[pix * src_cs + c];
  }
}

}  // namespace

int launch_norm_dm(int B, int hw, const float* dm, const float* coms, float* out, cudaStream_t st) {
  size_t n = (size_t)B * hw * hw;
  dr_launch(norm_dm_kernel, dim3(blocks_for(n)), dim3(EW_T), 0, st, B, hw * hw, dm, coms, out);
  return 1;
}
int launch_make_uvd(int B, int in_hw, int out_hw, const float* x0, float* tiny, UvdDst dst, cudaStream_t st) {
  size_t n = (size_t)B * out_hw * out_hw;
  dr_launch(make_uvd_kernel, dim3(blocks_for(n)), dim3(EW_T), 0, st, B, in_hw, out_hw, x0, tiny, dst);
  return 1;
}
int launch_maxpool(int B, int H, int W, int C, int k, const float* x, int x_cs, float* y, int y_cs, cudaStream_t st) {
  size_t n = (size_t)B * (H / 2) * (W / 2) * C;
  dr_launch(maxpool_kernel, dim3(blocks_for(n)), dim3(EW_T), 0, st, B, H, W, C, k, x, x_cs, y, y_cs);
  return 1;
}
int launch_maxpool_bwd(int B, int H, int W, int C, int k, const float* x, int x_cs, const float* dy, int dy_cs,
                       float* dx, int dx_cs, int accumulate, cudaStream_t st) {
  size_t n = (size_t)B * H * W * C;
  static int v4 = -1;
  if (v4 < 0) { const char* e = getenv("DENSEREG_POOL_BWD_V4"); v4 = (e && e[0] == '0') ? 0 : 1; }   // default on: -0.36 ms per micro-batch (profiles/r2_sweep.md)
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (v4 && C % 4 == 0 && x_cs % 4 == 0 && dy_cs % 4 == 0 && dx_cs % 4 == 0 && al(x) && al(dy) && al(dx)) {
    dr_launch(maxpool_bwd_v4_kernel, dim3(blocks_for(n / 4)), dim3(EW_T), 0, st, B, H, W, C / 4, k, (const float4*)x, x_cs / 4, (const float4*)dy, dy_cs / 4,
                                                              (float4*)dx, dx_cs / 4, accumulate);
    return 1;
  }
  dr_launch(maxpool_bwd_kernel, dim3(blocks_for(n)), dim3(EW_T), 0, st, B, H, W, C, k, x, x_cs, dy, dy_cs, dx, dx_cs, accumulate);
  return 1;
}
int launch_upadd(int B, int H, int W, int C, const float* a, int a_cs, const float* lo, int lo_cs, float* y, int y_cs, cudaStream_t st) {
  size_t n = (size_t)B * H * W * C;
  dr_launch(upadd_kernel, dim3(blocks_for(n)), dim3(EW_T), 0, st, B, H, W, C, a, a_cs, lo, lo_cs, y, y_cs);
  return 1;
}
int launch_upadd_bwd_lo(int B, int H, int W, int C, const float* dy, int dy_cs, float* dlo, int dlo_cs, int accumulate, cudaStream_t st) {
  size_t n = (size_t)B * (H / 2) * (W / 2) * C;
  dr_launch(upadd_bwd_lo_kernel, dim3(blocks_for(n)), dim3(EW_T), 0, st, B, H, W, C, dy, dy_cs, dlo, dlo_cs, accumulate);
  return 1;
}
int launch_copy_view(size_t npix, int C, const float* src, int src_cs, float* dst, int dst_cs, int accumulate,
                     const float* tiny_mask, cudaStream_t st) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (C % 4 == 0 && src_cs % 4 == 0 && dst_cs % 4 == 0 && al(src) && al(dst) && npix * (size_t)(C / 4) < 0xFFFFFFFFull) {
    const unsigned n4 = (unsigned)(npix * (C / 4));
    dr_launch(copy_view_v4_kernel, dim3(blocks_for(n4)), dim3(EW_T), 0, st, n4, C / 4, (const float4*)src, src_cs / 4, (float4*)dst, dst_cs / 4, accumulate, tiny_mask);
  } else {
    dr_launch(copy_view_kernel, dim3(blocks_for(npix * C)), dim3(EW_T), 0, st, npix, C, src, src_cs, dst, dst_cs, accumulate, tiny_mask);
  }
  return 1;
}
int launch_fill_view(size_t npix, int C, float* dst, int dst_cs, float v, cudaStream_t st) {
  dr_launch(fill_view_kernel, dim3(blocks_for(npix * C)), dim3(EW_T), 0, st, npix, C, dst, dst_cs, v);
  return 1;
}
static dim3 stats_grid(size_t npix, int C) {
  int gx = (C + 31) / 32;
  size_t gy = (npix + 8 * 32 - 1) / (8 * 32);          // >= 32 pixels per thread before splitting further
  size_t cap = (148 * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  return dim3(gx, (unsigned)gy);
}
int launch_channel_stats_finalize(size_t npix, int C, const float* x, int x_cs, double* sums, unsigned int* counter,
                                  const float* beta_gamma, float* state, float* aff, float* bstat, int update_state, cudaStream_t st) {
  dr_launch(channel_stats_finalize_kernel, dim3(stats_grid(npix, C)), dim3(dim3(32, 8)), 0, st, npix, C, x, x_cs, sums, counter, beta_gamma, state, aff, bstat,
                                                                            update_state);
  return 1;
}
// Streaming order.  The conv kernels walk their tiles from the start of a map to its end, so when a conv finishes, the END of its output is what the
// 126 MB L2 still holds.  The pass that reads that output next (training-mode BRN normalise after the conv; BRN-backward reduce after the dgrad that wrote
// dy) therefore walks from the end to the start -- it hits L2 for the tail of maps that do not fit (84 MB at 512 channels, batch 40), and leaves the START
// of its own output in L2 for the next conv.  DENSEREG_EW_REVERSE=0: everything front to back.
static int ew_reverse() {
  static int r = -1;
  if (r < 0) { const char* e = getenv("DENSEREG_EW_REVERSE"); r = (e && e[0] == '0') ? 0 : 1; }
  return r;
}
int launch_brn_apply(size_t npix, int C, const float* raw, int raw_cs, const float* aff, int relu,
                     const float* res, int res_cs, float* y, int y_cs, cudaStream_t st) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (C % 4 == 0 && raw_cs % 4 == 0 && y_cs % 4 == 0 && (!res || res_cs % 4 == 0) && al(raw) && al(y) && al(aff) && (!res || al(res)) &&
      npix * (size_t)(C / 4) < 0xFFFFFFFFull) {
    const unsigned n4 = (unsigned)(npix * (C / 4));
    dr_launch(brn_apply_v4_kernel, dim3(blocks_for(n4)), dim3(EW_T), 0, st, n4, C / 4, (const float4*)raw, raw_cs / 4, (const float4*)aff, relu,
                                                        (const float4*)res, res_cs / 4, (float4*)y, y_cs / 4, ew_reverse());
  } else {
    dr_launch(brn_apply_kernel, dim3(blocks_for(npix * C)), dim3(EW_T), 0, st, npix, C, raw, raw_cs, aff, relu, res, res_cs, y, y_cs);
  }
  return 1;
}
// block / grid of the channel-stationary float4 kernels: blockDim = C4 * floor(256 / C4); every thread gets >= 8 pixels when there are enough.
// `cap` bounds the grid.  The reduce kernel ends with 2*C double atomics per BLOCK onto 2*C addresses (64 lines at C = 512), so its time grows
// with the block count once the atomics serialise: 1184 blocks x 16 atomics per line = 19 k atomics per line ~ 20 us per launch, twice the
// streaming time (profiles/r1_final.md section 4) -> ONE block per SM for the reduce (round-2 sweep: 148 -> 20.90, 296 -> 20.95, 592 -> 21.40,
// 1184 -> 21.76 ms per micro-batch), 8 per SM for the apply.
// DENSEREG_BRN_BLOCKS overrides both.
static bool brn_v4_shape(size_t npix, int C, unsigned* block, unsigned* grid, size_t cap) {
  const unsigned C4 = (unsigned)C / 4;
  if (C % 4 != 0 || C4 == 0 || C4 > 256 || npix >= 0x7FFFFFFFull) return false;
  const unsigned lanes = 256 / C4;
  *block = C4 * lanes;
  size_t g = (npix + (size_t)lanes * 8 - 1) / ((size_t)lanes * 8);
  static long env_cap = -1;
  if (env_cap < 0) { const char* e = getenv("DENSEREG_BRN_BLOCKS"); env_cap = e ? atol(e) : 0; }
  if (env_cap > 0) cap = (size_t)env_cap;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  *grid = (unsigned)g;
  return true;
}
// one-launch BRN backward for small layers (brn_bwd_cluster_kernel); returns 0 when the layer is too big or not float4-addressable
int launch_brn_bwd_small(size_t npix, int C, const float* dy, int dy_cs, const float* raw, int raw_cs, const float* aff, const float* bstat,
                         const float* beta_gamma, int relu, float* draw, int draw_cs, float* gparam, cudaStream_t st) {
  static long max_elems = -1;
  if (max_elems < 0) { const char* e = getenv("DENSEREG_BRN_SMALL_ELEMS"); max_elems = e ? atol(e) : 96 * 1024; }     // npix * C at or below this: one cluster
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const unsigned C4 = (unsigned)C / 4;
  if (max_elems <= 0 || (long)(npix * (size_t)C) > max_elems || C % 4 != 0 || C4 == 0 || C4 > 128) return 0;
  if (raw_cs % 4 != 0 || dy_cs % 4 != 0 || draw_cs % 4 != 0 || !al(raw) || !al(dy) || !al(draw)) return 0;
  const unsigned lanes = 256 / C4, block = C4 * lanes;
  const size_t smem = ((size_t)block * 8 + 4 * (size_t)C) * sizeof(double);
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(BRN_CL); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = BRN_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = dr_pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 2;
  if (cudaLaunchKernelEx(&cfg, brn_bwd_cluster_kernel, (unsigned)npix, C4, (const float4*)dy, (unsigned)dy_cs / 4, (const float4*)raw, (unsigned)raw_cs / 4,
                         aff, bstat, beta_gamma, relu, (float4*)draw, (unsigned)draw_cs / 4, gparam) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return 1;
}
int launch_brn_bwd_reduce(size_t npix, int C, const float* dy, int dy_cs, const float* raw, int raw_cs,
                          const float* aff, const float* bstat, int relu, double* sums, cudaStream_t st) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  unsigned block, grid;
  if (raw_cs % 4 == 0 && dy_cs % 4 == 0 && al(raw) && al(dy) && brn_v4_shape(npix, C, &block, &grid, 148)) {
    dr_launch(brn_bwd_reduce_v4_kernel, dim3(grid), dim3(block), (size_t)block * 4 * 2 * sizeof(double), st, (unsigned)npix, (unsigned)C / 4, (const float4*)dy, dy_cs / 4,
                                                                                        (const float4*)raw, raw_cs / 4, aff, bstat, relu, sums, ew_reverse());
    return 1;
  }
  dr_launch(brn_bwd_reduce_kernel, dim3(stats_grid(npix, C)), dim3(dim3(32, 8)), 0, st, npix, C, dy, dy_cs, raw, raw_cs, aff, bstat, relu, sums);
  return 1;
}
int launch_brn_bwd_apply(size_t npix, int C, const float* dy, int dy_cs, const float* raw, int raw_cs,
                         const float* aff, const float* bstat, const float* beta_gamma, int relu,
                         const double* sums, float* draw, int draw_cs, float* gparam, cudaStream_t st) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  unsigned block, grid;
  if (raw_cs % 4 == 0 && dy_cs % 4 == 0 && draw_cs % 4 == 0 && al(raw) && al(dy) && al(draw) && brn_v4_shape(npix, C, &block, &grid, 148 * 8)) {
    dr_launch(brn_bwd_apply_v4_kernel, dim3(grid), dim3(block), 0, st, (unsigned)npix, (unsigned)C / 4, (const float4*)dy, dy_cs / 4, (const float4*)raw, raw_cs / 4, aff,
                                                    bstat, beta_gamma, relu, sums, (float4*)draw, draw_cs / 4, gparam);
  } else {
    dr_launch(brn_bwd_apply_kernel, dim3(blocks_for(npix * C)), dim3(EW_T), 0, st, npix, C, dy, dy_cs, raw, raw_cs, aff, bstat, beta_gamma, relu,
                                                                sums, draw, draw_cs, gparam);
  }
  return 1;
}
int launch_bias_bwd(size_t npix, int C, const float* dy, int dy_cs, const float* out, int out_cs, int relu, int dropout,
                    float* dz, int dz_cs, float* gbias, cudaStream_t st) {
  dr_launch(bias_bwd_kernel, dim3(stats_grid(npix, C)), dim3(dim3(32, 8)), 0, st, npix, C, dy, dy_cs, out, out_cs, relu, dropout, dz, dz_cs, gbias);
  return 1;
}
int launch_loss(const LossArgs& a, cudaStream_t st) {
  size_t n = (size_t)a.B * a.hw * a.hw * a.J;
  dr_launch(loss_kernel, dim3(blocks_for(n, EW_T * 2, 148 * 8)), dim3(EW_T), 0, st, a);
  return 1;
}
int launch_wd(size_t n, const float* params, const float* wdmask, float* grads, double* reg_acc, cudaStream_t st) {
  dr_launch(wd_kernel, dim3(blocks_for(n, EW_T * 4, 148 * 4)), dim3(EW_T), 0, st, n, params, wdmask, grads, reg_acc);
  return 1;
}
int launch_finish_loss(const double* acc, float* out5, cudaStream_t st) {
  dr_launch(finish_loss_kernel, dim3(1), dim3(1), 0, st, acc, out5);
  return 1;
}
int launch_adam(size_t n, float* p, const float* g, float* m, float* v, float divisor, float clip,
                float alpha, float omb1, float omb2, float eps, cudaStream_t st) {
  adam_kernel<<<blocks_for(n, EW_T * 4, 148 * 8), EW_T, 0, st>>>(n, p, g, m, v, divisor, clip, alpha, omb1, omb2, eps);
  return 1;
}
int launch_init_trunc_normal(size_t n, float* p, float stddev, uint64_t seed, cudaStream_t st) {
  init_trunc_normal_kernel<<<blocks_for(n), EW_T, 0, st>>>(n, p, stddev, seed);
  return 1;
}
int launch_gather_outputs(size_t npix, int C, const float* src, int src_cs, float* dst, cudaStream_t st) {
  gather_outputs_kernel<<<blocks_for(npix * C), EW_T, 0, st>>>(npix, C, src, src_cs, dst);
  return 1;
}
