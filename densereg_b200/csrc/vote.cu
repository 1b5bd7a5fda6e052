// vote.cu -- offset-vote post-process as ONE kernel (compiled with --fmad=false so that every fp32
// rounding matches the reference's op-by-op TF graph; see oracle/vote_numpy.py).
//
// Replaces (reference, /root/reference):
//   model/hourglass_um_crop_tiny.py:276-299  _resume_om
//   model/hourglass_um_crop_tiny.py:743-785  _xyz_estimation
//   model/hourglass_um_crop_tiny.py:598-627  _generate_candidates (tf.nn.top_k k=5, ties -> lower index)
//   model/hourglass_um_crop_tiny.py:629-682  _get_candidate_weights
//   model/hourglass_um_crop_tiny.py:684-741  _weighted_mean_shift (4^3 histogram seed, 10 iterations, bw 0.4)
//   data/preprocess.py:189-232 generate_xyzs_from_multi_cfgs, :157-170 unnorm_xyz_pose; data/util.py:20 _pro
//
// Layout: hm/hm3 (B,H,W,J) and um (B,H,W,3J) NHWC with an arbitrary channel stride (so the kernel can
// read the network's concat buffer in place); dm (B,H,W) dense.
// One CTA per sample.  blockDim = J*G: thread (g,j) streams pixels g, g+G, ... of joint j, so a warp
// reads runs of J consecutive floats (coalesced; fully contiguous when the stride equals J) and
// keeps a private sorted top-5; the G partial lists are merged through shared memory and one thread per
// joint finishes candidates -> weights -> histogram seed -> mean shift in registers.
// HBM-bound: algorithmic bytes per sample 4*H*W*(5J+1); the um plane is only gathered
// at the 5 winners per joint, so measured DRAM traffic is ~(2J+1)/(5J+1) of that.
#include "common.cuh"
#include <math.h>

#define VOTE_K 5
#define VOTE_MAX_THREADS 256
#define VOTE_U 8

struct VoteParams {
  int B, H, W, J, G;
  const float* hm; int hm_cs;
  const float* hm3; int hm3_cs;
  const float* um; int um_cs;
  const float* dm; const float* cfgs; const float* coms;
  float* xyz; int32_t* top5; int32_t* clamp_count;
};

DR_DEVINL void top5_insert(float (&s)[VOTE_K], int (&ix)[VOTE_K], float r, int p) {
  // strict '>' : later (higher-index) pixels never displace an equal score -> ties keep the lower index
  if (r > s[VOTE_K - 1]) {
    s[VOTE_K - 1] = r; ix[VOTE_K - 1] = p;
#pragma unroll
    for (int i = VOTE_K - 1; i > 0; --i) {
      if (s[i] > s[i - 1]) {
        float ts = s[i]; s[i] = s[i - 1]; s[i - 1] = ts;
        int ti = ix[i]; ix[i] = ix[i - 1]; ix[i - 1] = ti;
      }
    }
  }
}

DR_DEVINL float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

DR_DEVINL int f2i_trunc_sat(float v) {
  // tf.to_int32 truncates toward zero; NaN -> 0, +-inf saturate (same convention as the oracle)
  if (v != v) return 0;
  if (v >= 2147483520.f) return 2147483647;
  if (v <= -2147483648.f) return (int)0x80000000;
  return (int)v;
}

__global__ void __launch_bounds__(VOTE_MAX_THREADS)
vote_kernel(VoteParams a) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  __shared__ float s_sc[VOTE_MAX_THREADS * VOTE_K];
  __shared__ int s_ix[VOTE_MAX_THREADS * VOTE_K];
  const int b = blockIdx.x;
  const int J = a.J, G = a.G;
  const int t = threadIdx.x;
  const int P = a.H * a.W;
  const bool active = t < J * G;
  const int j = t % J, g = t / J;

  float sc[VOTE_K]; int ix[VOTE_K];
#pragma unroll
  for (int i = 0; i < VOTE_K; ++i) { sc[i] = -INFINITY; ix[i] = 0x7fffffff; }

  if (active) {
    const float* hm = a.hm + (size_t)b * P * a.hm_cs + j;
    const float* hm3 = a.hm3 + (size_t)b * P * a.hm3_cs + j;
    const float* dm = a.dm + (size_t)b * P;
    int p = g;
    // VOTE_U pixels per trip: issue all loads first (memory-level parallelism: ~3*VOTE_U 4-byte loads in flight per thread,
    // streaming / no L1 allocation), then insert in index order
    for (; p + (VOTE_U - 1) * G < P; p += VOTE_U * G) {
      float d[VOTE_U], h[VOTE_U], h3[VOTE_U];
#pragma unroll
      for (int u = 0; u < VOTE_U; ++u) {
        int q = p + u * G;
        d[u] = ld_stream(dm + q);
        h[u] = ld_stream(hm + (size_t)q * a.hm_cs);
        h3[u] = ld_stream(hm3 + (size_t)q * a.hm3_cs);
      }
#pragma unroll
      for (int u = 0; u < VOTE_U; ++u) {
        float r = (h[u] + 1.0f) * h3[u];                 // refined_hms = (hms+1)*hm3s          :764
        r = r * (d[u] < -0.99f ? 0.0f : 1.0f);           // * dms_mask                          :767-768
        top5_insert(sc, ix, r, p + u * G);
      }
    }
    for (; p < P; p += G) {
      float d = __ldg(dm + p);
      float h = __ldg(hm + (size_t)p * a.hm_cs);
      float h3 = __ldg(hm3 + (size_t)p * a.hm3_cs);
      float r = (h + 1.0f) * h3;
      r = r * (d < -0.99f ? 0.0f : 1.0f);
      top5_insert(sc, ix, r, p);
    }
#pragma unroll
    for (int i = 0; i < VOTE_K; ++i) { s_sc[t * VOTE_K + i] = sc[i]; s_ix[t * VOTE_K + i] = ix[i]; }
  }
  __syncthreads();
  if (t >= J) return;

  // ---- merge the G partial lists of joint j: 5 selection passes in (score desc, index asc) order ----
  int win[VOTE_K];
  {
    float ps = INFINITY; int pi = -1;
    for (int sel = 0; sel < VOTE_K; ++sel) {
      float bs = -INFINITY; int bi = 0x7fffffff;
      for (int gg = 0; gg < G; ++gg) {
        const int base = (gg * J + j) * VOTE_K;
        for (int i = 0; i < VOTE_K; ++i) {
          float s = s_sc[base + i]; int id = s_ix[base + i];
          bool after_prev = (s < ps) || (s == ps && id > pi);     // strictly after the previous pick
          bool better = (s > bs) || (s == bs && id < bi);
          if (after_prev && better) { bs = s; bi = id; }
        }
      }
      ps = bs; pi = bi;
      win[sel] = bi < P ? bi : P - 1;
    }
  }
  if (a.top5) {
    for (int i = 0; i < VOTE_K; ++i) a.top5[((size_t)b * J + j) * VOTE_K + i] = win[i];
  }

  // ---- candidates: votes at the 5 winners ---------------------------------------------------------
  const float* cfg = a.cfgs + (size_t)b * 6;
  const float comx = a.coms[b * 3 + 0], comy = a.coms[b * 3 + 1], comz = a.coms[b * 3 + 2];
  const float w_ratio = cfg[4] / (float)a.W;               // preprocess.py:212-216
  const float h_ratio = cfg[5] / (float)a.H;
  const float fx = cfg[0] / w_ratio, fy = cfg[1] / h_ratio, cx = cfg[2] / w_ratio, cy = cfg[3] / h_ratio;
  const float min_depth = comz - 300.0f * 0.5f;            // preprocess.py:203-204
  const float max_depth = comz + 300.0f * 0.5f;
  const float* hmb = a.hm + (size_t)b * P * a.hm_cs;
  const float* hm3b = a.hm3 + (size_t)b * P * a.hm3_cs;
  const float* umb = a.um + (size_t)b * P * a.um_cs;
  const float* dmb = a.dm + (size_t)b * P;

  float can[VOTE_K][3], wt[VOTE_K];
  int cell[VOTE_K];
  int n_clamped = 0;
#pragma unroll
  for (int k = 0; k < VOTE_K; ++k) {
    const int p = win[k];
    const int pi = p / a.W, pj = p % a.W;
    const float d = dmb[p];
    const float z = d < -0.99f ? max_depth : d * 300.0f + min_depth;     // preprocess.py:205-207
    float x = ((float)pj - cx) * (z / fx);                              // :218
    float y = ((float)pi - cy) * (z / fy);                              // :219
    x = (x - comx) / 100.0f; y = (y - comy) / 100.0f;                   // :222-224
    const float zn = (z - comz) / 100.0f;
    const float dd = 0.8f - hm3b[(size_t)p * a.hm3_cs + j] * 0.8f;      // _resume_om :288
    const float* u3 = umb + (size_t)p * a.um_cs + 3 * j;
    can[k][0] = x + u3[0] * dd;                                         // xyzs + oms :760
    can[k][1] = y + u3[1] * dd;
    can[k][2] = zn + u3[2] * dd;
    // _get_candidate_weights :640-664
    const float qx = can[k][0] * 100.0f + comx;
    const float qy = can[k][1] * 100.0f + comy;
    const float qz = can[k][2] * 100.0f + comz;
    const float uf = (qx * fx) / qz + cx;                               // util.py:20 _pro
    const float vf = (qy * fy) / qz + cy;
    int uu = f2i_trunc_sat(uf + 0.5f), vv = f2i_trunc_sat(vf + 0.5f);
    if (uu < 0 || uu >= a.W || vv < 0 || vv >= a.H) {
      ++n_clamped;                                                      // TF-CPU gather_nd would raise (documented deviation)
      uu = uu < 0 ? 0 : (uu >= a.W ? a.W - 1 : uu);
      vv = vv < 0 ? 0 : (vv >= a.H ? a.H - 1 : vv);
    }
    wt[k] = hmb[(size_t)(vv * a.W + uu) * a.hm_cs + j];                 // raw hm :663
    // histogram cell :703-705   clip((p+1)*2, 0, 3.9) -> int
    int cc = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float qv = (can[k][c] + 1.0f) * 2.0f;
      qv = fminf(fmaxf(qv, 0.0f), 3.9f);
      if (qv != qv) qv = 0.0f;
      cc = cc * 4 + (int)qv;
    }
    cell[k] = cc;
  }
  if (a.clamp_count && n_clamped) atomicAdd(a.clamp_count, n_clamped);

  // ---- seed: last (row-major) cell whose summed weight equals the histogram maximum :707-711 ----
  float csum[VOTE_K];
#pragma unroll
  for (int k = 0; k < VOTE_K; ++k) {
    float s = 0.0f;
#pragma unroll
    for (int k2 = 0; k2 < VOTE_K; ++k2) if (cell[k2] == cell[k]) s = s + wt[k2];   // scatter_nd sums in order
    csum[k] = s;
  }
  float hmax = 0.0f;                                 // >= 59 cells are empty (value 0)
#pragma unroll
  for (int k = 0; k < VOTE_K; ++k) hmax = (csum[k] > hmax || csum[k] != csum[k]) ? csum[k] : hmax;
  int seed_cell = 0;
  for (int c = 63; c >= 0; --c) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < VOTE_K; ++k) if (cell[k] == c) v = csum[k];
    if (v == hmax) { seed_cell = c; break; }
  }
  float cur[3];
  cur[0] = (float)(seed_cell >> 4) / 2.0f - 1.0f + 0.25f;              // :711-712
  cur[1] = (float)((seed_cell >> 2) & 3) / 2.0f - 1.0f + 0.25f;
  cur[2] = (float)(seed_cell & 3) / 2.0f - 1.0f + 0.25f;

  // ---- weighted Gaussian mean shift, 10 iterations, bandwidth 0.4 :715-721 -----------------------
  const float inv_sigma = -3.125f;                    // -1/(2*0.4*0.4) rounded to fp32
  for (int it = 0; it < 10; ++it) {
    float num0 = 0.f, num1 = 0.f, num2 = 0.f, den = 0.f;
#pragma unroll
    for (int k = 0; k < VOTE_K; ++k) {
      const float d0 = can[k][0] - cur[0], d1 = can[k][1] - cur[1], d2 = can[k][2] - cur[2];
      float s = (d0 * d0 + d1 * d1) + d2 * d2;
      s = expf(inv_sigma * s);
      s = s * wt[k];
      if (k == 0) { num0 = can[k][0] * s; num1 = can[k][1] * s; num2 = can[k][2] * s; den = s; }
      else { num0 = num0 + can[k][0] * s; num1 = num1 + can[k][1] * s; num2 = num2 + can[k][2] * s; den = den + s; }
    }
    cur[0] = num0 / den; cur[1] = num1 / den; cur[2] = num2 / den;
  }
  float* o = a.xyz + ((size_t)b * J + j) * 3;
  o[0] = cur[0] * 100.0f + comx;                                         // unnorm_xyz_pose preprocess.py:166
  o[1] = cur[1] * 100.0f + comy;
  o[2] = cur[2] * 100.0f + comz;
}

int launch_vote(int B, int H, int W, int J,
                const float* hm, int hm_cs, const float* hm3, int hm3_cs, const float* um, int um_cs,
                const float* dm, const float* cfgs, const float* coms,
                float* xyz, int32_t* top5, int32_t* clamp_count, cudaStream_t st) {
  VoteParams a;
  a.B = B; a.H = H; a.W = W; a.J = J;
  a.G = VOTE_MAX_THREADS / J;
  if (a.G > H * W) a.G = H * W;
  a.hm = hm; a.hm_cs = hm_cs; a.hm3 = hm3; a.hm3_cs = hm3_cs; a.um = um; a.um_cs = um_cs;
  a.dm = dm; a.cfgs = cfgs; a.coms = coms; a.xyz = xyz; a.top5 = top5; a.clamp_count = clamp_count;
  int threads = a.G * J;
  threads = (threads + 31) / 32 * 32;
  if (threads > VOTE_MAX_THREADS) threads = VOTE_MAX_THREADS;
  dr_launch(vote_kernel, dim3(B), dim3(threads), 0, st, a);
  return 1;
}
