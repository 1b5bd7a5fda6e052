// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory -> tcgen05.mma kind::tf32 -> TMEM (fp32 accumulate)
//   -> tcgen05.ld -> fused epilogue (scale/shift | bias, ReLU, dropout, residual add, accumulate) -> NHWC view.
//
// Replaces tf.nn.conv2d (network/slim/ops.py:282) + the folded BRN/bias/ReLU/add epilogue of every stride-1
// 1x1 / 3x3 conv of network/um_v1.py (97 % of the MACs), and -- with rotated/transposed weights -- TF's
// Conv2DBackpropInput (dgrad).  Not a GEMM library call: the kernel below is the whole op.
//
// GEMM view per CTA: D[128 pixels, BN couts] = sum_{tap} sum_{c-block of 32} A_tap[128,32] * W_tap[32,BN].
//   A tile: ONE 4-D TMA box (32 ch, W, 128/W rows [, images]) of the NHWC activation view at the tap's spatial
//           offset; out-of-image coordinates are zero-filled by TMA, which IS the SAME padding.  Because a tile is
//           whole image rows, smem row r == output pixel tile*128 + r.
//   B tile: 3-D TMA box (32 ch, BN couts, 1 tap) of the K-major weight copy [tap][cout][cin]; channels past Cin
//           and couts past Cout are zero-filled.
//   Both K-major, SWIZZLE_128B (32 fp32 = 128 B rows, 8-row 1024 B atoms): UMMA smem descriptors advance
//   32 B per K=8 MMA.  DR_PREC_TF32X3: four splitter warps rewrite each landed A tile in shared memory into
//   hi = rn_tf32(v) and lo = rn_tf32(v - hi), the weights arrive pre-split (two TMA boxes), and every k-step issues hi*lo + lo*hi + hi*hi into the same TMEM accumulator (fp32-class accuracy; algorithmic FLOPs unchanged).
// Warp roles (384 threads = 3 warpgroups, registers moved to the epilogue warpgroup with setmaxnreg): warp 0 TMA producer, warp 1 TMEM alloc +
//   MMA issuer, warps 2-3 idle, warps 4-7 operand splitter (3xTF32 only), warps 8-11 epilogue (TMEM lane quarter = warp_idx % 4).
//   mbarrier full/empty ring.
#include "conv_tc_epilogue.cuh"
#include <stdlib.h>

namespace {

using namespace tcconv;
using namespace tc;

// dynamic smem layout (1024 B aligned): [stage][A hi 16K | (A lo 16K) | B hi BN*128 | (B lo BN*128)] ... barriers ... tmem ptr
// PERSISTENT: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (n-tile fastest so neighbouring CTAs share the A tile in
// L2); the smem ring runs continuously across tiles and the accumulator is double-buffered in TMEM (2 x BN columns), so the
// epilogue of tile i (TMEM -> registers -> global, BRN statistics) overlaps the main loop of tile i+1.
template <bool SPLIT3>
__global__ void __launch_bounds__(TC1_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_wlo, TcParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.BN * TC_BK * 4;
  const int stage_bytes = (SPLIT3 ? 2 : 1) * (A_TILE_BYTES + b_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* split_bar = empty_bar + p.stages;       // A tile split done (SPLIT3)
  uint64_t* acc_full = split_bar + p.stages;        // [2] accumulator stage complete (MMA -> epilogue)
  uint64_t* acc_empty = acc_full + 2;               // [2] accumulator stage drained (epilogue -> MMA)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.ksz * p.ksz * p.kblocks_per_tap;
  const int total_tiles = p.tiles_m * p.tiles_n;
  // Two-level accumulation (p.chunk_kb > 0): the tensor core's fp32 accumulate truncates, which biases long reductions (measured 1.6e-5 of
  // the output scale at K = 2304 against 3e-7 for FFMA).  The k-blocks of a tile are therefore cut into chunks of `ch`; each chunk is summed
  // by the MMAs into one of the two partial accumulator stages (first MMA of the chunk overwrites), and the epilogue warps add finished
  // partials into a running sum held in their registers with round-to-nearest fp32 adds while the MMAs fill the other stage
  // (conv_tc_epilogue.cuh: tc_flush_partial / tc_fold_running).  Tiles with num_kb <= ch are unchanged.
  const int ch = (p.chunk_kb > 0 && num_kb > p.chunk_kb) ? p.chunk_kb : num_kb;
  const int nchunks = (num_kb + ch - 1) / ch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], SPLIT_THREADS / 32); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                      // everything above (barriers, tensor-memory allocation) overlapped the previous kernel's tail

  if (warp < TC1_WARP_SPLIT0) {
  DR_SETMAXNREG_DEC(REG_CTRL);                           // warpgroup 0 (control): hand registers to the epilogue warpgroup
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(A_TILE_BYTES + (SPLIT3 ? 2 : 1) * b_bytes);   // A fp32 (split in smem); B hi (+ lo, pre-split)
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tile_m = tile / p.tiles_n, n0 = (tile - tile_m * p.tiles_n) * p.BN;
        const int pix0 = tile_m * TC_BM;
        const int img = pix0 / (p.H * p.W);
        const int y0 = (pix0 - img * p.H * p.W) / p.W;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const int tap = kb / p.kblocks_per_tap;
          const int c0 = (kb - tap * p.kblocks_per_tap) * TC_BK;
          const int dy = tap / p.ksz - p.pad, dx = tap % p.ksz - p.pad;
          uint8_t* st = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], tx);
          tma_load_4d(&map_a, &full_bar[s], st, c0, dx, y0 + dy, img);
          uint8_t* bdst = st + (SPLIT3 ? 2 : 1) * A_TILE_BYTES;
          const int wtap = p.flip_taps ? p.ksz * p.ksz - 1 - tap : tap;
          tma_load_3d(&map_w, &full_bar[s], bdst, c0, n0, wtap);
          if (SPLIT3) tma_load_3d(&map_wlo, &full_bar[s], bdst + b_bytes, c0, n0, wtap);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      uint32_t it = 0, ccount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        for (int c = 0; c < nchunks; ++c, ++ccount) {
          const uint32_t as = ccount & 1, aph = (ccount >> 1) & 1;
          mbar_wait(&acc_empty[as], aph ^ 1);            // epilogue has drained this accumulator stage
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + as * (uint32_t)p.acc_stride;
          const int kb0 = c * ch, kb1 = kb0 + ch < num_kb ? kb0 + ch : num_kb;
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            if (SPLIT3) mbar_wait(&split_bar[s], ph); else mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
            const uint32_t b_addr = a_addr + (SPLIT3 ? 2 : 1) * A_TILE_BYTES;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
              const uint64_t ad = make_desc(a_addr + k * 32), bd = make_desc(b_addr + k * 32);
              const uint32_t acc = ((kb - kb0) | k) != 0;                // first MMA of a chunk overwrites the partial accumulator
              if (SPLIT3) {
                const uint64_t ald = make_desc(a_addr + A_TILE_BYTES + k * 32), bld = make_desc(b_addr + b_bytes + k * 32);
                tc_mma_tf32(tmem_d, ad, bld, idesc, acc);                // hi * lo
                tc_mma_tf32(tmem_d, ald, bd, idesc, 1);                  // lo * hi
                tc_mma_tf32(tmem_d, ad, bd, idesc, 1);                   // hi * hi
              } else {
                tc_mma_tf32(tmem_d, ad, bd, idesc, acc);
              }
            }
            tc_commit(&empty_bar[s]);            // frees the smem slot when these MMAs retire
          }
          tc_commit(&acc_full[as]);              // partial accumulator complete -> epilogue warps
        }
      }
    }
  }
  } else if (warp >= TC1_WARP_EPI0) {
    DR_SETMAXNREG_INC(REG_EPI);                           // warpgroup 2 (epilogue)
    // ===================== epilogue: TMEM -> registers -> global =====================
    __shared__ float s_sum[4][256], s_sq[4][256];
    __shared__ __align__(16) float s_scale[TC_MAX_COUT], s_shift[TC_MAX_COUT];
    __shared__ __align__(16) float s_stage[4][32 * TC_STAGE_LD];        // per-warp 32x32 transpose tile (coalesced epilogue)
    __shared__ int s_last;
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int et = q * 32 + lane;                        // 0..127
    const bool vec_ok = ((p.y_cs & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                        (!p.res || (((p.res_cs & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0)));
    tc_epilogue_stage_affine(p, et, s_scale, s_shift);
    uint32_t ccount = 0;
    const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tile_m = tile / p.tiles_n, n0 = (tile - tile_m * p.tiles_n) * p.BN;
      float run[4][32];                                   // running sum of this tile's finished partials (two-level accumulation only)
      for (int c = 0; c + 1 < nchunks; ++c, ++ccount) {
        const uint32_t as = ccount & 1, aph = (ccount >> 1) & 1;
        mbar_wait_sleep(&acc_full[as], aph);
        tc_fence_after();
        const uint32_t tl = tmem_base + as * (uint32_t)p.acc_stride + lane_bits;
        if (c == 0) tc_flush_partial<true>(tl, p.BN, run); else tc_flush_partial<false>(tl, p.BN, run);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);       // the MMA warp may overwrite this partial stage
      }
      {
        const uint32_t as = ccount & 1, aph = (ccount >> 1) & 1;
        mbar_wait_sleep(&acc_full[as], aph);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * (uint32_t)p.acc_stride;
        if (nchunks > 1) tc_fold_running(tacc + lane_bits, p.BN, run);
        tc_epilogue_tile(p, tacc, q, lane, row, et, vec_ok, tile_m, n0, p.BN, total_tiles, s_sum, s_sq, s_last,
                         s_scale, s_shift, s_stage[q], [&]() { mbar_arrive(&acc_empty[as]); });
        ++ccount;
      }
    }
    tc_epilogue_finish(p, et, s_scale, s_shift, s_last);
  } else {
    DR_SETMAXNREG_DEC(REG_SPLIT);                         // warpgroup 1 (splitters; idle for single-pass tf32)
    if (SPLIT3) {
    // ===================== A splitter: hi = rn_tf32(a), lo = rn_tf32(a - hi) =====================
    const int t = threadIdx.x - TC1_WARP_SPLIT0 * 32;     // 0..SPLIT_THREADS-1
    const int na4 = A_TILE_BYTES / 16;                    // weights arrive pre-split (hi, lo) by TMA; only the A tile is split here
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        float4* a_hi = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
        float4* a_lo = a_hi + na4;
        for (int idx = t; idx < na4; idx += SPLIT_THREADS) {   // elementwise: the swizzled layout is irrelevant
          const float4 a = a_hi[idx];
          float4 h, l;
          h.x = tf32_rna(a.x); l.x = tf32_rna(a.x - h.x);
          h.y = tf32_rna(a.y); l.y = tf32_rna(a.y - h.y);
          h.z = tf32_rna(a.z); l.z = tf32_rna(a.z - h.z);
          h.w = tf32_rna(a.w); l.w = tf32_rna(a.w - h.w);
          a_hi[idx] = h; a_lo[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);                       // one arrival per splitter warp
      }
    }
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

}  // namespace

// Weight operands for the tensor-core path are K-major copies [tap][cout][cin] prepared by the engine (p.w points at
// that copy when ConvProblem::w_kmajor is set).  Eligibility: stride 1, k in {1,3}, square power-of-two maps that tile
// into whole rows (W <= 128, 128 % W == 0), input view 16 B aligned with a channel stride that is a multiple of 4 floats (Cin itself may be ragged: 131, 65, 515 ...;
// TMA zero-fills the channel tail of both operands).
const char* tc_last_error() { return tc::last_error(); }

bool conv_tc_eligible(const ConvProblem& p) {
  if (!p.w_kmajor) return false;
  if (p.stride != 1 || (p.k != 1 && p.k != 3)) return false;
  if (p.H != p.W || p.Ho != p.H || p.Wo != p.W) return false;
  if (p.W < 2 || p.W > 128 || (128 % p.W) != 0) return false;
  if ((p.W * p.H) % 128 != 0 && 128 % (p.W * p.H) != 0) return false;
  if (p.Cin < 8 || (p.x_cs % 4) != 0 || (reinterpret_cast<uintptr_t>(p.x) & 15) != 0) return false;   // only STRIDES must be 16 B multiples
  if (p.wk_ld < p.Cin || (p.wk_ld % 4) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p.w_kmajor) & 15) != 0) return false;
  if (p.Cout < 8 || (p.Cout + 255) / 256 * 256 > TC_MAX_COUT) return false;     // shared-memory scale/shift tables
  if (p.pad_t != (p.k - 1) / 2 || p.pad_l != p.pad_t) return false;
  // tiny problems are latency-bound on the TMA/mbarrier pipeline (~18 us floor); below the threshold the FFMA kernel is used
  static double min_mmac = -1.0;
  if (min_mmac < 0) { const char* e = getenv("DENSEREG_TC_MIN_MMAC"); min_mmac = e ? atof(e) : 0.0; }
  if (min_mmac > 0 && (double)p.B * p.H * p.W * p.k * p.k * p.Cin * p.Cout < min_mmac * 1e6) return false;
  return true;
}

int launch_conv_tc(const ConvProblem& p, int split3, cudaStream_t st) {
  static bool attr_set[2] = {false, false};
  // two-level accumulation (kernel comment): ConvProblem::chunk_kb = k-blocks per partial accumulator, 0 = off (the engine sets it: on for
  // inference, off for training -- engine.cu).  3xTF32 only; chunked layers run on this kernel or on the A-in-tensor-memory kernel (a 256-column
  // tile of the CTA-pair kernel would need 256 registers per epilogue thread for its running sum).
  const int chunk_kb = p.chunk_kb > 0 ? p.chunk_kb : 0;
  const int num_kb_all = p.k * p.k * ((p.Cin + TC_BK - 1) / TC_BK);
  const bool chunked = split3 && chunk_kb > 0 && num_kb_all > chunk_kb && num_kb_all > p.chunk_min_kb;   // short reductions stay one-level
  if (split3) {
    const bool pair_w = !chunked && conv_tc_pair_wanted(p);   // the CTA-pair kernel (BN = 256) has no register room for a running sum
    const int am = conv_tc_atmem_mode();                   // opt-in: split A operand in tensor memory (conv_tc_atmem.cu)
    if (am == 2 || (am == 1 && !pair_w)) {
      const int n = launch_conv_tc_atmem(p, st);
      if (n > 0) return n;                                 // 0 = shape not taken by that variant (no launch attempted)
    }
    if (pair_w) return launch_conv_tc_pair(p, st);         // CTA-pair kernel for the big layers; a failed launch is an error, not a reason to run another kernel
  }
  TcParams t; memset(&t, 0, sizeof(t));
  t.M = p.B * p.H * p.W; t.H = p.H; t.W = p.W; t.Cin = p.Cin; t.Cout = p.Cout; t.ksz = p.k; t.pad = p.pad_t; t.flip_taps = p.flip_taps;
  int BN = (p.Cout + 15) / 16 * 16;
  if (BN > 256) BN = 256;
  if (split3 && BN > 128) BN = 128;                        // 3 stages of [A hi|lo, B hi|lo] fit; 2 stages at BN=256 measured slower
  { static int bn_cap = -1; if (bn_cap < 0) { const char* e = getenv("DENSEREG_TC_BN"); bn_cap = e ? atoi(e) : 0; }   // tuning knob
    if (bn_cap >= 16 && BN > bn_cap) BN = bn_cap / 16 * 16; }
  t.BN = BN;
  t.kblocks_per_tap = (p.Cin + TC_BK - 1) / TC_BK;
  t.acc_stride = (BN + 31) / 32 * 32;
  int cols = 32; while (cols < 2 * t.acc_stride) cols <<= 1;   // two accumulator stages
  t.tmem_cols = cols;
  t.chunk_kb = chunked ? chunk_kb : 0;
  t.tiles_m = (t.M + TC_BM - 1) / TC_BM; t.tiles_n = (p.Cout + BN - 1) / BN;
  const int stage_bytes = (split3 ? 2 : 1) * (A_TILE_BYTES + BN * TC_BK * 4);
  // shared-memory budget: 227 KB per CTA minus the kernel's static part (statistics staging + scale/shift tables), barriers, alignment slack
  static int smem_budget = 0;
  if (!smem_budget) {
    cudaFuncAttributes fa;
    const size_t st_bytes = cudaFuncGetAttributes(&fa, conv_tc_kernel<true>) == cudaSuccess ? fa.sharedSizeBytes : 16 * 1024;
    smem_budget = 227 * 1024 - (int)st_bytes - 1536;     // 1536 >= barriers + tmem slot + 1024 B alignment slack
  }
  int stages = smem_budget / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) stages = 2;
  const int num_kb = p.k * p.k * t.kblocks_per_tap;
  if (stages > num_kb) stages = num_kb < 2 ? 2 : num_kb;
  t.stages = stages;
  t.y = p.y; t.y_cs = p.y_cs; t.scale = p.scale; t.shift = p.shift; t.relu = p.relu; t.res = p.res; t.res_cs = p.res_cs;
  t.accumulate = p.accumulate; t.dropout = p.dropout; t.drop_seed = p.drop_seed; t.drop_tag = p.drop_tag;
  t.stats = p.stats; t.stats_counter = p.stats_counter; t.bn_bg = p.bn_bg; t.bn_state = p.bn_state; t.bn_aff = p.bn_aff; t.bn_bstat = p.bn_bstat;
  t.bn_update_state = p.bn_update_state;
  { static int co = -1; if (co < 0) { const char* e = getenv("DENSEREG_TC_EPI_COALESCE"); co = (e && e[0] == '0') ? 0 : 1; } t.coalesce = co; }
  { static int pc = -1; if (pc < 0) { const char* e = getenv("DENSEREG_TC_STATS_PER_CTA"); pc = (e && e[0] == '0') ? 0 : 1; }   // default on: measured -0.12 ms per micro-batch (profiles/r2_sweep.md)
    t.stats_per_cta = (pc && p.stats && !p.scale && !p.shift) ? 1 : 0; }
  const size_t smem_bytes = (size_t)stages * stage_bytes + (3 * stages + 4) * 8 + 16 + 1024 + 64;   // + 8.2 KB static (fused-stats staging)

  // activation map: dims (C, W, H, B)
  CUtensorMap ma, mw, mwlo;
  const int rows = TC_BM / p.W;                            // image rows per tile (may exceed H -> several images)
  const int bh = rows < p.H ? rows : p.H;
  const int bb = rows < p.H ? 1 : rows / p.H;
  cuuint64_t ad[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  cuuint64_t as[3] = {(cuuint64_t)p.x_cs * 4, (cuuint64_t)p.W * p.x_cs * 4, (cuuint64_t)p.H * p.W * p.x_cs * 4};
  cuuint32_t ab[4] = {(cuuint32_t)TC_BK, (cuuint32_t)p.W, (cuuint32_t)bh, (cuuint32_t)bb};
  if (!encode_map(&ma, p.x, 4, ad, as, ab)) return 0;
  // weight map: dims (Cin, Cout, taps) over the K-major copy
  cuuint64_t wd[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)(p.k * p.k)};
  cuuint64_t ws[2] = {(cuuint64_t)p.wk_ld * 4, (cuuint64_t)p.wk_ld * p.Cout * 4};
  cuuint32_t wb[3] = {(cuuint32_t)TC_BK, (cuuint32_t)BN, 1};
  if (!encode_map(&mw, p.w_kmajor, 3, wd, ws, wb)) return 0;
  mwlo = mw;
  if (split3) { if (!encode_map(&mwlo, p.w_kmajor_lo, 3, wd, ws, wb)) return 0; }

  static int num_sms = 0;
  if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
  const int total_tiles = t.tiles_m * t.tiles_n;
  dim3 grid(total_tiles < num_sms ? total_tiles : num_sms);
  if (split3) {
    if (!attr_set[1]) { cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_budget + 1536); attr_set[1] = true; }
    dr_launch(conv_tc_kernel<true>, dim3(grid), dim3(TC1_THREADS), smem_bytes, st, ma, mw, mwlo, t);
    return launch_ok(cudaPeekAtLastError(), "conv_tc_kernel<3xTF32>") ? 1 : 0;
  } else {
    if (!attr_set[0]) { cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_budget + 1536); attr_set[0] = true; }
    dr_launch(conv_tc_kernel<false>, dim3(grid), dim3(TC1_THREADS), smem_bytes, st, ma, mw, mwlo, t);
  }
  return launch_ok(cudaPeekAtLastError(), "conv_tc_kernel<TF32>") ? 1 : 0;
}
