// conv_simt.cu -- fp32 (FFMA) implicit-GEMM convolution: forward / dgrad (same kernel, transformed
// weights) and wgrad.  This is the PARITY path (DR_PREC_FP32) and the fallback for layers whose shape
// the tcgen05 path does not take (Cin % 4 != 0 views, the 7x7/s2 stem, tiny spatial levels).
//
// Replaces tf.nn.conv2d NHWC/HWIO SAME (network/slim/ops.py:282) and TF's Conv2DBackpropInput /
// Conv2DBackpropFilter for every conv2d call of network/um_v1.py.
//
// GEMM view: Y[M,N] = A[M,K] * Wm[K,N],  M = B*Ho*Wo pixels, N = Cout, K = k*k*Cin (tap-major, channel
// minor == HWIO flattening).  Tile 128x64x16, 256 threads, 8x4 register tile per thread, double-
// buffered shared memory with register prefetch.  A is gathered with zero fill for SAME padding.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int NT = 256;
constexpr int APAD = 4;

struct PixCoord { int base; int iy0; int ix0; };   // base = b*H*W (pixel units) or -1 if row invalid

template <bool VEC4>
__global__ void __launch_bounds__(NT)
conv_fwd_kernel(ConvProblem p) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int M = p.B * p.Ho * p.Wo;
  const int K = p.k * p.k * p.Cin;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- A-load mapping -------------------------------------------------------------------------
  // VEC4: thread loads float4 along k: kq = tid%4 (k offset 4*kq), rows r = tid/4 + 64*i (i<2)
  // scalar: kk = tid%16, rows r = tid/16 + 16*i (i<8)
  constexpr int A_ROWS = VEC4 ? 2 : 8;
  const int a_k = VEC4 ? (tid & 3) * 4 : (tid & 15);
  const int a_r0 = VEC4 ? (tid >> 2) : (tid >> 4);
  constexpr int A_RSTEP = VEC4 ? 64 : 16;
  PixCoord pc[A_ROWS];
#pragma unroll
  for (int i = 0; i < A_ROWS; ++i) {
    int m = m0 + a_r0 + i * A_RSTEP;
    if (m < M) {
      int b = m / (p.Ho * p.Wo);
      int r = m - b * p.Ho * p.Wo;
      int oy = r / p.Wo, ox = r - oy * p.Wo;
      pc[i].base = b * p.H * p.W;
      pc[i].iy0 = oy * p.stride - p.pad_t;
      pc[i].ix0 = ox * p.stride - p.pad_l;
    } else {
      pc[i].base = -1; pc[i].iy0 = 0; pc[i].ix0 = 0;
    }
  }
  // ---- B-load mapping: float4 along n when Cout%4==0 (checked at run time per element group) ----
  const int b_n = (tid & 15) * 4;       // 16 threads x 4 = 64 columns
  const int b_k = tid >> 4;             // 16 rows
  const int w_ld = p.w_ld ? p.w_ld : p.Cout;
  const bool b_vec = (w_ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.w) & 15) == 0);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float a_reg[A_ROWS][VEC4 ? 4 : 1];
  float b_reg[4];

  auto load_tiles = [&](int k0) {
    // A
    int kg = k0 + a_k;
    int tap = kg / p.Cin;
    int c = kg - tap * p.Cin;
    int dy = tap / p.k, dx = tap - dy * p.k;
    bool kvalid = kg < K;
#pragma unroll
    for (int i = 0; i < A_ROWS; ++i) {
      int iy = pc[i].iy0 + dy, ix = pc[i].ix0 + dx;
      bool ok = kvalid && pc[i].base >= 0 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      if (VEC4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)(pc[i].base + iy * p.W + ix) * p.x_cs + c));
        a_reg[i][0] = v.x; a_reg[i][1 % (VEC4 ? 4 : 1)] = v.y; a_reg[i][2 % (VEC4 ? 4 : 1)] = v.z; a_reg[i][3 % (VEC4 ? 4 : 1)] = v.w;
      } else {
        a_reg[i][0] = ok ? __ldg(p.x + (size_t)(pc[i].base + iy * p.W + ix) * p.x_cs + c) : 0.f;
      }
    }
    // B
    int kb = k0 + b_k;
    if (kb < K) {
      int kbs = kb;
      if (p.flip_taps) { int tp = kb / p.Cin; kbs = (p.k * p.k - 1 - tp) * p.Cin + (kb - tp * p.Cin); }
      const float* wr = p.w + (size_t)kbs * w_ld + n0 + b_n;
      if (b_vec && n0 + b_n + 3 < p.Cout) {
        float4 v = __ldg(reinterpret_cast<const float4*>(wr));
        b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) b_reg[j] = (n0 + b_n + j < p.Cout) ? __ldg(wr + j) : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) b_reg[j] = 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_ROWS; ++i) {
      int r = a_r0 + i * A_RSTEP;
      if (VEC4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) As[buf][a_k + q][r] = a_reg[i][q % (VEC4 ? 4 : 1)];
      } else {
        As[buf][a_k][r] = a_reg[i][0];
      }
    }
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  const int ty = tid >> 4, tx = tid & 15;     // 16 x 16 threads; rows ty*8.., cols tx*4..
  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue -------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    float* yr = p.y + (size_t)m * p.y_cs;
    const float* rr = p.res ? p.res + (size_t)m * p.res_cs : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j];
      if (p.scale) v = v * __ldg(p.scale + n);
      if (p.shift) v = v + __ldg(p.shift + n);
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.dropout) v = dr_hash_keep(p.drop_seed, p.drop_tag, (uint64_t)m * p.Cout + n) ? v * 2.0f : 0.f;
      if (rr) v += rr[n];
      if (p.accumulate) v += yr[n];
      yr[n] = v;
    }
  }
}

// ---- wgrad: dW[K,N] += A[M,K]^T * dY[M,N], split over M across blockIdx.z, fp32 atomics -----------
constexpr int WK = 64, WN = 64, WM = 16;

__global__ void __launch_bounds__(NT)
conv_wgrad_kernel(WgradProblem p, int m_per_block) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  __shared__ __align__(16) float As[2][WM][WK];
  __shared__ __align__(16) float Bs[2][WM][WN];
  const int tid = threadIdx.x;
  const int M = p.B * p.Ho * p.Wo;
  const int K = p.k * p.k * p.Cin;
  const int k0 = blockIdx.x * WK, n0 = blockIdx.y * WN;
  const int m_begin = blockIdx.z * m_per_block;
  const int m_end = min(M, m_begin + m_per_block);

  // A load: kk = tid%64 (fixed tap/channel per thread), rows mm = tid/64 + 4*i (i<4)
  const int a_k = tid & 63;
  const int a_m = tid >> 6;
  const int kg = k0 + a_k;
  const bool kvalid = kg < K;
  const int tap = kvalid ? kg / p.Cin : 0;
  const int c = kg - tap * p.Cin;
  const int dy = tap / p.k, dx = tap - dy * p.k;
  // B load: n = (tid%16)*4, rows mm = tid/16
  const int b_n = (tid & 15) * 4, b_m = tid >> 4;
  const bool b_vec = (p.dy_cs % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dy) & 15) == 0);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float a_reg[4], b_reg[4];
  const int HoWo = p.Ho * p.Wo;

  auto load_tiles = [&](int mb) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = mb + a_m + i * 4;
      float v = 0.f;
      if (kvalid && m < m_end) {
        int b = m / HoWo; int r = m - b * HoWo; int oy = r / p.Wo, ox = r - oy * p.Wo;
        int iy = oy * p.stride - p.pad_t + dy, ix = ox * p.stride - p.pad_l + dx;
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
          v = __ldg(p.x + (size_t)((b * p.H + iy) * p.W + ix) * p.x_cs + c);
      }
      a_reg[i] = v;
    }
    int m = mb + b_m;
    if (m < m_end) {
      const float* r = p.dy + (size_t)m * p.dy_cs + n0 + b_n;
      if (b_vec && n0 + b_n + 3 < p.Cout) {
        float4 v = __ldg(reinterpret_cast<const float4*>(r));
        b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) b_reg[j] = (n0 + b_n + j < p.Cout) ? __ldg(r + j) : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) b_reg[j] = 0.f;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[buf][a_m + i * 4][a_k] = a_reg[i];
    *reinterpret_cast<float4*>(&Bs[buf][b_m][b_n]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
  };

  const int ty = tid >> 4, tx = tid & 15;   // k rows ty*4.., n cols tx*4..
  const int nm = (m_end - m_begin + WM - 1) / WM;
  if (nm <= 0) return;
  load_tiles(m_begin);
  store_tiles(0);
  __syncthreads();
  for (int mt = 0; mt < nm; ++mt) {
    const int buf = mt & 1;
    if (mt + 1 < nm) load_tiles(m_begin + (mt + 1) * WM);
#pragma unroll
    for (int mm = 0; mm < WM; ++mm) {
      float4 av = *reinterpret_cast<const float4*>(&As[buf][mm][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][mm][tx * 4]);
      float a4[4] = {av.x, av.y, av.z, av.w};
      float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    if (mt + 1 < nm) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int kk = k0 + ty * 4 + i;
    if (kk >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < p.Cout) atomicAdd(p.dw + (size_t)kk * p.Cout + n, acc[i][j]);
    }
  }
}

}  // namespace

int launch_conv_simt(const ConvProblem& p, cudaStream_t st) {
  const int M = p.B * p.Ho * p.Wo;
  dim3 grid((M + BM - 1) / BM, (p.Cout + BN - 1) / BN);
  const bool vec = (p.Cin % 4 == 0) && (p.x_cs % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
  if (vec) dr_launch(conv_fwd_kernel<true>, dim3(grid), dim3(NT), 0, st, p);
  else dr_launch(conv_fwd_kernel<false>, dim3(grid), dim3(NT), 0, st, p);
  return 1;
}

int launch_wgrad_simt(const WgradProblem& p, cudaStream_t st) {
  const int M = p.B * p.Ho * p.Wo;
  const int K = p.k * p.k * p.Cin;
  const int gx = (K + WK - 1) / WK, gy = (p.Cout + WN - 1) / WN;
  // enough M-splits to fill the 148 SMs a few times over, at least 64 pixels (4 tiles of WM) per block.  (512 pixels per block left the
  // 4x4 / 2x2 hourglass levels -- 640 / 160 pixels at batch 40 -- on 2 blocks walking 20 tiles each: 35 us per launch, 18 launches per
  // micro-batch; profiles/r2_sweep.md)
  int splits = (148 * 4 + gx * gy - 1) / (gx * gy);
  int max_splits = (M + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int m_per_block = ((M + splits - 1) / splits + WM - 1) / WM * WM;
  splits = (M + m_per_block - 1) / m_per_block;
  dim3 grid(gx, gy, splits);
  dr_launch(conv_wgrad_kernel, dim3(grid), dim3(NT), 0, st, p, m_per_block);
  return 1;
}
