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
// One-CTA kernels (conv_tc.cu, conv_tc_atmem.cu): three warpgroups so that registers can be moved between roles with setmaxnreg --
//   warpgroup 0: warp 0 TMA producer, warp 1 MMA issuer (+ tensor-memory alloc), warps 2-3 idle;  warpgroup 1: operand splitters;
//   warpgroup 2: epilogue.  A 384-thread CTA gets 168 registers per thread at launch; the control and splitter warpgroups hand most of theirs
//   to the epilogue warps, which then have room for the running sums of the two-level accumulation (128 registers for a 128-column tile).
constexpr int TC1_THREADS = 384;
constexpr int TC1_WARP_SPLIT0 = 4, TC1_WARP_EPI0 = 8;
constexpr int REG_CTRL = 40, REG_SPLIT = 96, REG_EPI = 232;
#define DR_SETMAXNREG_DEC(n) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(n))
#define DR_SETMAXNREG_INC(n) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(n))

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
  int stats_per_cta;     // fused BRN statistics: accumulate per CTA in shared memory, ONE round of atomics + fence + counter per CTA (opt-in)
  int coalesce;          // epilogue: transpose each 32x32 chunk through shared memory so that global stores / residual loads are whole 128 B rows
  int acc_stride;        // tensor-memory columns per accumulator stage: BN rounded up to 32 (whole tcgen05.ld / st groups belong to ONE stage)
  int chunk_kb;          // > 0: two-level accumulation -- the tensor core sums at most chunk_kb k-blocks into a partial accumulator (its fp32 adds
                         // truncate), partials are added into a running sum with round-to-nearest fp32 adds by the epilogue warps (conv_tc.cu)
};
constexpr int TC_STAGE_LD = 36;      // floats per staging row: 16 B aligned, conflict-free for float4 writes (row per lane) and row reads

using namespace tc;

constexpr int TC_MAX_COUT = 768;     // capacity of the shared-memory scale/shift tables (largest Cout on the path: 515, dgrad of um_full1)

// Per-channel epilogue constants, staged ONCE per kernel in shared memory by the 128 epilogue threads (named barrier 1): the tile
// loop then reads them with broadcast LDS.128 instead of two dependent global loads per element.  Channels past Cout get scale 0 /
// shift 0 (they are never stored).
DR_DEVINL void tc_epilogue_stage_affine(const TcParams& p, int et, float* s_scale, float* s_shift) {
  const int ncol = p.tiles_n * p.BN;
  if (p.stats && p.stats_per_cta) {      // raw output (no affine): the two tables serve as this CTA's running column sums / sums of squares
    for (int c = et; c < ncol; c += 128) { s_scale[c] = 0.f; s_shift[c] = 0.f; }
  } else {
    for (int c = et; c < ncol; c += 128) {
      s_scale[c] = (c < p.Cout) ? (p.scale ? __ldg(p.scale + c) : 1.f) : 0.f;
      s_shift[c] = (c < p.Cout && p.shift) ? __ldg(p.shift + c) : 0.f;
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

// Epilogue of ONE 128-row accumulator tile: TMEM -> registers -> (BRN statistics) -> scale/shift | bias, ReLU, dropout,
// residual add, accumulate -> NHWC view; then the BRN finalize if this was the last tile of the layer.  Called by the four
// epilogue warps (128 threads, named barrier 1).  tmem_acc = TMEM address of column 0 of this accumulator stage (lane 0);
// release() is invoked by lane 0 of every warp once that warp's TMEM reads are complete (hands the stage back to the MMA issuer).
// Every warp-uniform decision (scale? shift? relu? dropout? residual? accumulate? whole 32-column chunk inside Cout?) is taken ONCE per
// 32-column chunk, and the residual / accumulate operands of a chunk are fetched as one batch of independent 16 B loads, so the chunk
// is a straight line of independent instructions (the first version interleaved two dependent global loads and four branches per
// element: 0.1 instructions per cycle per warp, 12 us per 128x128 tile -- profiles/r1_epilogue.md).
template <class Release>
DR_DEVINL void tc_epilogue_tile(const TcParams& p, uint32_t tmem_acc, int q, int lane, int row, int et, bool vec_ok, int tile_m, int n0, int bn,
                                int total_tiles, float (*s_sum)[256], float (*s_sq)[256], int& s_last, float* s_scale,
                                float* s_shift, float* stg, Release release) {
    const int m = tile_m * TC_BM + row;
    const bool mvalid = m < p.M;
    float* yr = p.y + (size_t)m * p.y_cs;
    const float* rr = p.res ? p.res + (size_t)m * p.res_cs : nullptr;
    const bool has_scale = p.scale != nullptr, has_shift = p.shift != nullptr;
    for (int cb = 0; cb < bn; cb += 32) {          // bn = columns of this work item (p.BN)
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
      const int nc = n0 + cb;                       // first channel of this chunk
      if (nc >= p.Cout) continue;
      const bool chunk_vec = vec_ok && nc + 32 <= p.Cout;      // whole chunk inside Cout and 16 B aligned rows (warp-uniform)
      const bool transposed = chunk_vec && p.coalesce;
      if (!transposed && !mvalid) continue;
      float x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]);
      // ---- per-channel affine from the shared-memory tables (separate multiply and add, never contracted: same roundings as before)
      if (has_scale) {
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale + nc);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 c4 = sc4[g];
          x[4 * g] = __fmul_rn(x[4 * g], c4.x); x[4 * g + 1] = __fmul_rn(x[4 * g + 1], c4.y);
          x[4 * g + 2] = __fmul_rn(x[4 * g + 2], c4.z); x[4 * g + 3] = __fmul_rn(x[4 * g + 3], c4.w);
        }
      }
      if (has_shift) {
        const float4* sh4 = reinterpret_cast<const float4*>(s_shift + nc);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 c4 = sh4[g];
          x[4 * g] = __fadd_rn(x[4 * g], c4.x); x[4 * g + 1] = __fadd_rn(x[4 * g + 1], c4.y);
          x[4 * g + 2] = __fadd_rn(x[4 * g + 2], c4.z); x[4 * g + 3] = __fadd_rn(x[4 * g + 3], c4.w);
        }
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = fmaxf(x[i], 0.f);
      }
      if (p.dropout) {
        const uint64_t base = (uint64_t)m * p.Cout + nc;
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = dr_hash_keep(p.drop_seed, p.drop_tag, base + i) ? x[i] * 2.0f : 0.f;
      }
      if (transposed) {
        // ---- lane = row  ->  (8 lanes per row) x (4 rows per instruction): every global access of the warp is 4 whole 128 B row segments
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(stg + lane * TC_STAGE_LD + 4 * g) = make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
        __syncwarp();
        const int rsub = lane >> 3, cc = (lane & 7) * 4;
        const int mrow0 = tile_m * TC_BM + q * 32 + rsub;
        float4 o4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o4[j] = *reinterpret_cast<const float4*>(stg + (4 * j + rsub) * TC_STAGE_LD + cc);
        if (p.res) {
          float4 r4[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int mr = mrow0 + 4 * j;
            r4[j] = mr < p.M ? *reinterpret_cast<const float4*>(p.res + (size_t)mr * p.res_cs + nc + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) { o4[j].x += r4[j].x; o4[j].y += r4[j].y; o4[j].z += r4[j].z; o4[j].w += r4[j].w; }
        }
        if (p.accumulate) {
          float4 y4[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int mr = mrow0 + 4 * j;
            y4[j] = mr < p.M ? *reinterpret_cast<const float4*>(p.y + (size_t)mr * p.y_cs + nc + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) { o4[j].x += y4[j].x; o4[j].y += y4[j].y; o4[j].z += y4[j].z; o4[j].w += y4[j].w; }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int mr = mrow0 + 4 * j;
          if (mr < p.M) *reinterpret_cast<float4*>(p.y + (size_t)mr * p.y_cs + nc + cc) = o4[j];
        }
        __syncwarp();                               // the staging tile is rewritten by the next chunk
      } else if (chunk_vec) {
        // ---- whole chunk inside Cout, 16 B aligned rows: batched 16 B loads, then 8 x 16 B stores
        if (rr) {
          float4 r4[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) r4[g] = *reinterpret_cast<const float4*>(rr + nc + 4 * g);
#pragma unroll
          for (int g = 0; g < 8; ++g) { x[4 * g] += r4[g].x; x[4 * g + 1] += r4[g].y; x[4 * g + 2] += r4[g].z; x[4 * g + 3] += r4[g].w; }
        }
        if (p.accumulate) {
          float4 y4[8];
#pragma unroll
          for (int g = 0; g < 8; ++g) y4[g] = *reinterpret_cast<const float4*>(yr + nc + 4 * g);
#pragma unroll
          for (int g = 0; g < 8; ++g) { x[4 * g] += y4[g].x; x[4 * g + 1] += y4[g].y; x[4 * g + 2] += y4[g].z; x[4 * g + 3] += y4[g].w; }
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(yr + nc + 4 * g) = make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
      } else {
        // ---- ragged chunk (channel tail, or rows that are not 16 B aligned): element by element
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int nn = nc + i;
          if (nn < p.Cout) {
            float o = x[i];
            if (rr) o += rr[nn];
            if (p.accumulate) o += yr[nn];
            yr[nn] = o;
          }
        }
      }
    }
    // all of this warp's TMEM reads are complete (tcgen05.wait::ld inside tmem_ld32): hand the stage back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) release();
    if (p.stats && p.stats_per_cta) {
      // per-CTA accumulation: the tile's column sums go into the shared-memory running totals (tc_epilogue_finish publishes them once)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int col = et; col < bn; col += 128) {
        const int n = n0 + col;
        if (n < p.Cout) {
          s_scale[n] += (s_sum[0][col] + s_sum[1][col]) + (s_sum[2][col] + s_sum[3][col]);
          s_shift[n] += (s_sq[0][col] + s_sq[1][col]) + (s_sq[2][col] + s_sq[3][col]);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");                     // s_sum / s_sq are rewritten by the next tile
    } else if (p.stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int col = et; col < bn; col += 128) {
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

// ---- two-level accumulation (TcParams::chunk_kb): running sums of the finished partial accumulators live in REGISTERS of the epilogue
// warps -- lane = accumulator row, run[j][i] = column 32*j + i (BN <= 128) -- so a flush is one tensor-memory read of the partial (~1000
// cycles for 128 x 128 at 64 B/clk, hidden behind the >= 1300 MMA cycles of the next k-block) and round-to-nearest fp32 adds.
template <bool FIRST>
DR_DEVINL void tc_flush_partial(uint32_t tacc_lane, int bn, float (&run)[4][32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j * 32 < bn) {
      uint32_t v[32];
      tmem_ld32(tacc_lane + (uint32_t)(j * 32), v);
#pragma unroll
      for (int i = 0; i < 32; ++i) run[j][i] = FIRST ? __uint_as_float(v[i]) : __fadd_rn(run[j][i], __uint_as_float(v[i]));
    }
  }
}
// last chunk of a tile: partial += running sum, written back into the partial's own tensor-memory columns so that the (unchanged) epilogue
// reads the complete accumulator from there
DR_DEVINL void tc_fold_running(uint32_t tacc_lane, int bn, const float (&run)[4][32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (j * 32 < bn) {
      uint32_t v[32];
      tmem_ld32(tacc_lane + (uint32_t)(j * 32), v);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__fadd_rn(run[j][i], __uint_as_float(v[i])));
      tmem_st32(tacc_lane + (uint32_t)(j * 32), v);
    }
  }
  tmem_wait_st();
}

// After a CTA's last tile (per-CTA statistics mode only): publish the running totals with one round of double atomics, then the usual
// "last arriver finalizes" protocol counted in CTAs instead of tiles.  Called by the 128 epilogue threads.
DR_DEVINL void tc_epilogue_finish(const TcParams& p, int et, const float* acc_sum, const float* acc_sq, int& s_last) {
  if (!(p.stats && p.stats_per_cta)) return;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  for (int n = et; n < p.Cout; n += 128) {
    const float a = acc_sum[n], b = acc_sq[n];
    if (a != 0.f || b != 0.f) { atomicAdd(p.stats + n, (double)a); atomicAdd(p.stats + p.Cout + n, (double)b); }
  }
  __threadfence();
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (et == 0) s_last = (atomicAdd(p.stats_counter, 1u) == gridDim.x - 1);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (s_last) {                                                         // last CTA of the layer: BRN finalize
    __threadfence();
    brn_finalize_dev<1, 128>(et, p.Cout, (double)p.M, p.stats, p.bn_bg, p.bn_state, p.bn_aff, p.bn_bstat, p.bn_update_state);
  }
}

}  // namespace tcconv
