// crop.cu -- depth-frame front-end, compiled with --fmad=false so that the bilinear lerps round like the
// reference's un-fused fp32 ops.
//
// Replaces (reference, /root/reference):
//   data/preprocess.py:10-79    crop_from_xyz_pose (bbox from the projected pose +-pad, crop_to_bounding_box, pad_to_bounding_box to a
//                               square, tf.image.resize_images bilinear -> 128x128, depth threshold, rescaled CameraConfig)
//   data/preprocess.py:81-129   crop_from_bbx (NYU test boxes)
//   data/preprocess.py:131-142  center_of_mass
// Three small launches per batch: per-crop parameters (one warp per frame), gather/lerp (one thread per output pixel, HBM-bound:
// 64 KiB written per crop), centre of mass (mean of the positive pixels accumulated by the gather kernel).
#include "common.cuh"
#include <math.h>

struct CropParams { int top, left, bottom, right, L, off_h, off_w; float thr; };

namespace {

__device__ __forceinline__ void finish_params(CropParams& cp, const float* cfg, int out_hw, float* cfg_out) {
  const int bh = cp.bottom - cp.top, bw = cp.right - cp.left;
  cp.L = bh > bw ? bh : bw;
  cp.off_h = (int)((double)(cp.L - cp.bottom + cp.top) / 2.0);          // tf.to_int32(tf.divide(int, 2)) : truediv in double, truncate
  cp.off_w = (int)((double)(cp.L - cp.right + cp.left) / 2.0);
  const float ratio = (float)((double)cp.L / (double)out_hw);             // tf.cast(longer_edge/out_w, tf.float32) :70-71
  cfg_out[0] = cfg[0] / ratio; cfg_out[1] = cfg[1] / ratio;
  cfg_out[2] = ((cfg[2] - (float)cp.left) + (float)cp.off_w) / ratio;     // :75-78
  cfg_out[3] = ((cfg[3] - (float)cp.top) + (float)cp.off_h) / ratio;
  cfg_out[4] = (float)out_hw; cfg_out[5] = (float)out_hw;
}

struct Cfg6 { float v[6]; };

__global__ void crop_params_pose_kernel(int in_h, int in_w, const float* __restrict__ frames, const float* __restrict__ poses, int J,
                                        Cfg6 cfg, int out_hw, float pad, int icvl, CropParams* __restrict__ params,
                                        float* __restrict__ cfg_out) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const float* dm = frames + (size_t)b * in_h * in_w;
  const float* pose = poses + (size_t)b * 3 * J;
  float umin = INFINITY, umax = -INFINITY, vmin = INFINITY, vmax = -INFINITY, dmin = INFINITY;
  for (int j = lane; j < J; j += 32) {
    const float x = pose[3 * j], y = pose[3 * j + 1], z = pose[3 * j + 2];
    const float u = (x * cfg.v[0]) / z + cfg.v[2];                         // data/util.py:20
    const float v = (y * cfg.v[1]) / z + cfg.v[3];
    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
    int uu = (int)u, vv = (int)v;                                          // tf.to_int32 truncates :56-57
    uu = uu < 0 ? 0 : (uu > in_w - 1 ? in_w - 1 : uu);
    vv = vv < 0 ? 0 : (vv > in_h - 1 ? in_h - 1 : vv);
    const float d = dm[(size_t)vv * in_w + uu];
    if (d > 100.0f) dmin = fminf(dmin, d);                                 // boolean_mask(dd > 100) ; reduce_min :59-61
  }
  for (int o = 16; o > 0; o >>= 1) {
    umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o)); umax = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, o));
    vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o)); vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
  }
  if (lane == 0) {
    CropParams cp;
    const float top = fminf(fmaxf(vmin - pad, 0.0f), cfg.v[5] - 2.0f * pad);              // :29-32
    const float left = fminf(fmaxf(umin - pad, 0.0f), cfg.v[4] - 2.0f * pad);
    const float bottom = fmaxf(fminf(vmax + pad, cfg.v[5]), (top + 2.0f * pad) - 1.0f);
    const float right = fmaxf(fminf(umax + pad, cfg.v[4]), (left + 2.0f * pad) - 1.0f);
    cp.top = (int)top; cp.left = (int)left; cp.bottom = (int)bottom; cp.right = (int)right;
    cp.thr = icvl ? 500.0f : dmin + 250.0f;                                               // :61-65
    finish_params(cp, cfg.v, out_hw, cfg_out + (size_t)b * 6);
    params[b] = cp;
  }
}

__global__ void crop_params_bbx_kernel(const float* __restrict__ bbx, Cfg6 cfg, int out_hw, CropParams* __restrict__ params,
                                       float* __restrict__ cfg_out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  CropParams cp;
  cp.top = (int)bbx[b * 5 + 0]; cp.left = (int)bbx[b * 5 + 1]; cp.bottom = (int)bbx[b * 5 + 2]; cp.right = (int)bbx[b * 5 + 3];
  cp.thr = bbx[b * 5 + 4];
  finish_params(cp, cfg.v, out_hw, cfg_out + (size_t)b * 6);
  params[b] = cp;
}

// value of the zero-padded square at (y, x)
__device__ __forceinline__ float sq_at(const float* __restrict__ dm, int in_w, const CropParams& cp, int y, int x) {
  const int cy = y - cp.off_h, cx = x - cp.off_w;
  if (cy < 0 || cx < 0 || cy >= cp.bottom - cp.top || cx >= cp.right - cp.left) return 0.0f;
  return dm[(size_t)(cp.top + cy) * in_w + cp.left + cx];
}

__global__ void crop_resize_kernel(int in_h, int in_w, const float* __restrict__ frames, const CropParams* __restrict__ params, int out_hw,
                                   float* __restrict__ out, double* __restrict__ sums, unsigned int* __restrict__ counts) {
  const int b = blockIdx.x, oy = blockIdx.y;
  const CropParams cp = params[b];
  const float* dm = frames + (size_t)b * in_h * in_w;
  const float scale = (float)cp.L / (float)out_hw;                          // resize_bilinear: in/out, align_corners = false
  const float ys = (float)oy * scale;
  const int y0 = (int)floorf(ys); int y1 = (int)ceilf(ys); if (y1 > cp.L - 1) y1 = cp.L - 1;
  const float yl = ys - (float)y0;
  float psum = 0.f; unsigned int pcnt = 0;
  for (int ox = threadIdx.x; ox < out_hw; ox += blockDim.x) {
    const float xs = (float)ox * scale;
    const int x0 = (int)floorf(xs); int x1 = (int)ceilf(xs); if (x1 > cp.L - 1) x1 = cp.L - 1;
    const float xl = xs - (float)x0;
    const float tl = sq_at(dm, in_w, cp, y0, x0), tr = sq_at(dm, in_w, cp, y0, x1);
    const float bl = sq_at(dm, in_w, cp, y1, x0), br = sq_at(dm, in_w, cp, y1, x1);
    const float top = tl + (tr - tl) * xl;
    const float bot = bl + (br - bl) * xl;
    float v = top + (bot - top) * yl;
    v = v < cp.thr ? v : 0.0f;                                                // :62-65 / :114
    out[((size_t)b * out_hw + oy) * out_hw + ox] = v;
    if (v > 0.0f) { psum += v; ++pcnt; }
  }
  for (int o = 16; o > 0; o >>= 1) { psum += __shfl_xor_sync(0xffffffffu, psum, o); pcnt += __shfl_xor_sync(0xffffffffu, pcnt, o); }
  __shared__ float s_sum[32]; __shared__ unsigned int s_cnt[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s_sum[w] = psum; s_cnt[w] = pcnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0; unsigned int c = 0;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) { s += (double)s_sum[i]; c += s_cnt[i]; }
    if (c) { atomicAdd(sums + b, s); atomicAdd(counts + b, c); }
  }
}

__global__ void crop_com_kernel(int B, int out_hw, const double* __restrict__ sums, const unsigned int* __restrict__ counts,
                                const float* __restrict__ cfg_out, float* __restrict__ com) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float ave_d = counts[b] ? (float)(sums[b] / (double)counts[b]) : 200.0f;     // reduce_mean(boolean_mask(dm, dm > 0)) :135
  ave_d = fmaxf(ave_d, 200.0f);                                               // :137
  const float* c = cfg_out + (size_t)b * 6;
  const float ave_u = (float)((double)out_hw / 2.0), ave_v = ave_u;           // tf.cast(c_w/2, tf.float32)
  com[b * 3 + 0] = ((ave_u - c[2]) * ave_d) / c[0];                           // :139-140
  com[b * 3 + 1] = ((ave_v - c[3]) * ave_d) / c[1];
  com[b * 3 + 2] = ave_d;
}

// ---- data augmentation: data/preprocess.py:234-267 data_aug ------------------------------------------------------
// rotate (tf.contrib.image.rotate, NEAREST) -> nearest resize by edge_ratio -> centred crop/pad back to (h,w), composed into ONE
// gather per output pixel; the random draws (cos/sin of the angle, edge ratios) are inputs so the result is reproducible.
__device__ __forceinline__ int round_half_away(float v) { return (int)(v >= 0.f ? floorf(v + 0.5f) : ceilf(v - 0.5f)); }

__global__ void data_aug_image_kernel(int B, int h, int w, const float* __restrict__ dms, const float* __restrict__ cs /*B,2*/,
                                      const float* __restrict__ er /*B,2*/, float* __restrict__ out) {
  const size_t n = (size_t)B * h * w;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / ((size_t)h * w)); const int r = (int)(i - (size_t)b * h * w); const int y = r / w, x = r - y * w;
    const float cost = cs[b * 2], sint = cs[b * 2 + 1];
    const int th = (int)((float)h * er[b * 2]), tw = (int)((float)w * er[b * 2 + 1]);       // tf.to_int32(tf.to_float(shape)*ratio)
    // resize_image_with_crop_or_pad (floor-division offsets)
    const int wd = w - tw, hd = h - th;
    auto fdiv2 = [](int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); };                      // python floor division by 2
    const int ocw = max(fdiv2(-wd), 0), opw = max(fdiv2(wd), 0), och = max(fdiv2(-hd), 0), oph = max(fdiv2(hd), 0);
    const int yr = y - oph + och, xr = x - opw + ocw;                                         // coordinate in the resized image
    float v = 0.f;
    if (y >= oph && x >= opw && yr < th && xr < tw && (y - oph) < (th < h ? th : h) && (x - opw) < (tw < w ? tw : w)) {
      // ResizeNearestNeighbor, align_corners = false
      int ys = (int)floorf((float)yr * ((float)h / (float)th)); if (ys > h - 1) ys = h - 1;
      int xs = (int)floorf((float)xr * ((float)w / (float)tw)); if (xs > w - 1) xs = w - 1;
      // projective transform of tf.contrib.image.rotate (output -> input), NEAREST
      const float x_off = ((float)(w - 1) - (cost * (float)(w - 1) - sint * (float)(h - 1))) / 2.0f;
      const float y_off = ((float)(h - 1) - (sint * (float)(w - 1) + cost * (float)(h - 1))) / 2.0f;
      const float xi = (cost * (float)xs + (-sint) * (float)ys) + x_off;
      const float yi = (sint * (float)xs + cost * (float)ys) + y_off;
      const int rx = round_half_away(xi), ry = round_half_away(yi);
      if (rx >= 0 && rx < w && ry >= 0 && ry < h) v = dms[((size_t)b * h + ry) * w + rx];
    }
    out[i] = v;
  }
}

__global__ void data_aug_pose_kernel(int B, int J, const float* __restrict__ poses, const float* __restrict__ cfgs, const float* __restrict__ coms,
                                     const float* __restrict__ cs, const float* __restrict__ er, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * J) return;
  const int b = i / J;
  const float* cfg = cfgs + b * 6; const float* com = coms + b * 3; const float* p = poses + (size_t)i * 3;
  const float cost = cs[b * 2], sint = cs[b * 2 + 1];
  const float ucom = (com[0] * cfg[0]) / com[2] + cfg[2], vcom = (com[1] * cfg[1]) / com[2] + cfg[3];   // xyz2uvd_op(com) :241
  const float u = ((p[0] * cfg[0]) / p[2] + cfg[2]) - ucom, v = ((p[1] * cfg[1]) / p[2] + cfg[3]) - vcom, d = p[2] - com[2];
  float ur = u * cost + v * sint, vr = u * (-sint) + v * cost;                                           // uvd_pt @ rot_mat :247
  ur = ur * er[b * 2 + 1] + ucom; vr = vr * er[b * 2] + vcom;                                           // :258-260
  const float dr = d + com[2];
  out[(size_t)i * 3 + 0] = ((ur - cfg[2]) * dr) / cfg[0];                                               // uvd2xyz_op :261 (util.py:21)
  out[(size_t)i * 3 + 1] = ((vr - cfg[3]) * dr) / cfg[1];
  out[(size_t)i * 3 + 2] = dr;
}

}  // namespace

int launch_data_aug(int B, int hw, int J, const float* dms, const float* poses, const float* cfgs, const float* coms, const float* cossin,
                    const float* edge_ratio, float* dms_out, float* poses_out, cudaStream_t st) {
  const size_t n = (size_t)B * hw * hw;
  int blocks = (int)((n + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
  data_aug_image_kernel<<<blocks, 256, 0, st>>>(B, hw, hw, dms, cossin, edge_ratio, dms_out);
  data_aug_pose_kernel<<<(B * J + 127) / 128, 128, 0, st>>>(B, J, poses, cfgs, coms, cossin, edge_ratio, poses_out);
  return 2;
}

int launch_crop(int B, int in_h, int in_w, const float* frames, const float* poses, int J, const float* bbx, const float cfg_host[6],
                int out_hw, float pad, int icvl, void* scratch /* B*(sizeof(CropParams)+16) bytes */, float* dm_out, float* cfg_out,
                float* com_out, cudaStream_t st) {
  Cfg6 cfg; for (int i = 0; i < 6; ++i) cfg.v[i] = cfg_host[i];
  CropParams* params = reinterpret_cast<CropParams*>(scratch);
  double* sums = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + (((size_t)B * sizeof(CropParams) + 15) / 16) * 16);
  unsigned int* counts = reinterpret_cast<unsigned int*>(sums + B);
  cudaMemsetAsync(sums, 0, (size_t)B * (sizeof(double) + sizeof(unsigned int)), st);
  if (poses) crop_params_pose_kernel<<<B, 32, 0, st>>>(in_h, in_w, frames, poses, J, cfg, out_hw, pad, icvl, params, cfg_out);
  else crop_params_bbx_kernel<<<(B + 127) / 128, 128, 0, st>>>(bbx, cfg, out_hw, params, cfg_out, B);
  crop_resize_kernel<<<dim3(B, out_hw), 128, 0, st>>>(in_h, in_w, frames, params, out_hw, dm_out, sums, counts);
  crop_com_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, out_hw, sums, counts, cfg_out, com_out);
  return 3;
}

size_t crop_scratch_bytes(int B) { return (((size_t)B * sizeof(CropParams) + 15) / 16) * 16 + (size_t)B * 16 + 64; }
