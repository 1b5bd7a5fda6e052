// conv_tc_epilogue.cuh -- parameters and the fused per-tile epilogue shared by the tcgen05 conv kernels
// (conv_tc.cu: one CTA per 128-pixel tile; conv_tc_pair.cu: CTA pairs, cta_group::2, 256-pixel tiles).
#pragma once
#include "tc_common.cuh"
#include "brn.cuh"

namespace tcconv {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // fp32 elements per k-block = one 128 B swizzle row
constexpr int A_TILE_BYTES = TC_BM * TC_BK * 4; // 16 KB
constexpr int SPLIT_THREADS = 128;             // 4 splitter warps (3xTF32); 8 measured slower (issue-slot pressure on the MMA thread)

struct TcParams {
  int M;                 // B*H*W output pixels
  int H, W;              // spatial (stride 1: input == output size)
  int Cin, Cout;
  int ksz, pad;          // 1 or 3; pad (before)
  int flip_taps;         // dgrad: weight tap index reversed (180-degree rotation)
  int BN;                // N tile (multiple of 16, <= 256)
  int tiles_m, tiles_n;  // persistent tile walk
  int kblocks_per_tap;   // ceil(Cin / 32)
  int stages;
  int tmem_cols;         // power of two >= BN, >= 32
  float* y; int y_cs;
  const float* scale; const float* shift; int relu;
  const float* res; int res_cs; int accumulate;
  int dropout; unsigned long long drop_seed; unsigned int drop_tag;
  double* stats; unsigned int* stats_counter; const float* bn_bg; float* bn_state; float* bn_aff; float* bn_bstat; int bn_update_state;
};

using namespace tc;

// Epilogue of ONE 128-row accumulator tile: TMEM -> registers -> (BRN statistics) -> scale/shift | bias, ReLU, dropout,
// residual add, accumulate -> NHWC view; then the BRN finalize if this was the last tile of the layer.  Called by the four
// epilogue warps (128 threads, named barrier 1).  tmem_acc = TMEM address of column 0 of this accumulator stage (lane 0);
// release() is invoked by lane 0 of every warp once that warp's TMEM reads are complete (hands the stage back to the MMA issuer).
template <class Release>
DR_DEVINL void tc_epilogue_tile(const TcParams& p, uint32_t tmem_acc, int q, int lane, int row, int et, bool vec_ok, int tile_m, int n0,
                                int total_tiles, float (*s_sum)[256], float (*s_sq)[256], int& s_last, Release release) {
    const int m = tile_m * TC_BM + row;
    const bool mvalid = m < p.M;
    float* yr = p.y + (size_t)m * p.y_cs;
    const float* rr = p.res ? p.res + (size_t)m * p.res_cs : nullptr;
    for (int cb = 0; cb < p.BN; cb += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
      if (p.stats) {
        // column sums over this warp's 32 rows by recursive halving: 31 shuffles, lane l ends with column cb+l
        float a[32], b2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { a[i] = __uint_as_float(v[i]); b2[i] = a[i] * a[i]; }   // rows past M are exact zeros
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) {
          const bool up = (lane & sft) != 0;
#pragma unroll
          for (int j = 0; j < sft; ++j) {
            const float sa = up ? a[j] : a[j + sft], ka = up ? a[j + sft] : a[j];
            const float sb = up ? b2[j] : b2[j + sft], kb2 = up ? b2[j + sft] : b2[j];
            a[j] = ka + __shfl_xor_sync(0xffffffffu, sa, sft);
            b2[j] = kb2 + __shfl_xor_sync(0xffffffffu, sb, sft);
          }
        }
        s_sum[q][cb + lane] = a[0]; s_sq[q][cb + lane] = b2[0];
      }
      if (!mvalid) continue;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int n = n0 + cb + g * 4;
        if (n >= p.Cout) break;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int nn = n + e;
          float x = __uint_as_float(v[g * 4 + e]);
          if (nn < p.Cout) {
            if (p.scale) x = x * __ldg(p.scale + nn);
            if (p.shift) x = x + __ldg(p.shift + nn);
            if (p.relu) x = fmaxf(x, 0.f);
            if (p.dropout) x = dr_hash_keep(p.drop_seed, p.drop_tag, (uint64_t)m * p.Cout + nn) ? x * 2.0f : 0.f;
          }
          o[e] = x;
        }
        if (vec_ok && n + 3 < p.Cout) {
          if (rr) {
            const float4 r4 = *reinterpret_cast<const float4*>(rr + n);
            o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
          }
          if (p.accumulate) {
            const float4 y4 = *reinterpret_cast<const float4*>(yr + n);
            o[0] += y4.x; o[1] += y4.y; o[2] += y4.z; o[3] += y4.w;
          }
          *reinterpret_cast<float4*>(yr + n) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int nn = n + e;
            if (nn < p.Cout) {
              float x = o[e];
              if (rr) x += rr[nn];
              if (p.accumulate) x += yr[nn];
              yr[nn] = x;
            }
          }
        }
      }
    }
    // all of this warp's TMEM reads are complete (tcgen05.wait::ld inside tmem_ld32): hand the stage back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) release();
    if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int col = et; col < p.BN; col += 128) {
        const int n = n0 + col;
        if (n < p.Cout) {
          atomicAdd(p.stats + n, (double)((s_sum[0][col] + s_sum[1][col]) + (s_sum[2][col] + s_sum[3][col])));
          atomicAdd(p.stats + p.Cout + n, (double)((s_sq[0][col] + s_sq[1][col]) + (s_sq[2][col] + s_sq[3][col])));
        }
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) s_last = (atomicAdd(p.stats_counter, 1u) == (unsigned)total_tiles - 1);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (s_last) {                                                     // last TILE of the layer: BRN finalize
        __threadfence();
        brn_finalize_dev<1, 128>(et, p.Cout, (double)p.M, p.stats, p.bn_bg, p.bn_state, p.bn_aff, p.bn_bstat, p.bn_update_state);
      }
    }
}

}  // namespace tcconv
