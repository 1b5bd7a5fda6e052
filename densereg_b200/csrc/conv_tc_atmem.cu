// conv_tc_atmem.cu -- 3xTF32 implicit-GEMM convolution with the split A operand held in TENSOR MEMORY (tcgen05.mma, A from TMEM).
//
// Why (profiles/r1_final.md section 1): the one-CTA 3xTF32 kernel of conv_tc.cu is bound by shared-memory bandwidth, and most of the bytes are
// A-side: per 32-channel k-block the 128x32 activation tile costs 16 KB of TMA writes + 16 KB splitter reads + 32 KB splitter writes (hi, lo) +
// 3 MMAs x 4 k-steps x 4 KB = 48 KB of UMMA operand reads = 112 KB of the 192 KB total (BN = 128), independent of the tile width -- which is why
// the narrow layers (Cout 64 / 80) run at a quarter of the tensor rate.  Here the four splitter warps read the landed fp32 tile once from shared
// memory (row per lane, un-swizzling the 16 B chunks), form hi = rn_tf32(a) and lo = rn_tf32(a - hi) in registers and write both with
// tcgen05.st into tensor memory; the MMAs take A from TMEM ([d], [a], b-desc form) and only B from shared memory.  Shared-memory bytes per
// k-block at BN = 128: 48 KB TMA (A fp32 + B hi + B lo) + 16 KB splitter reads + 48 KB B operand reads = 112 KB per 800 MMA cycles = 143 B/clk
// (was 240 B/clk), and the A-lo copy leaves the ring: 48 KB stages, 4 of them.
//
// Tensor-memory plan (512 columns): [0, 2*BN) two accumulator stages (as in conv_tc.cu); [a0 + 64*s, +64), a0 = 2*BN rounded up to 32, for ring stage s: 32 columns of
// A-hi (column = k, lane = pixel row) followed by 32 columns of A-lo.  BN <= 128, stages <= (512 - 2*BN) / 64.
//
// Roles and barriers are those of conv_tc_kernel<3xTF32>; the only differences are the splitter's destination and the MMA's A operand.
// STATUS: opt-in (DENSEREG_TC_A_TMEM=1, or 2 to prefer it over the CTA-pair kernel); written after the round-1 GPU budget was spent, compiled
// for sm_100a but not yet run -- tests/test_gpu_experimental.py and tools/r2_sweep.sh cover it.
#include "conv_tc_epilogue.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

using namespace tcconv;

DR_DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]   (cute::SM100_MMA_TF32_TS: A from tensor memory is always K-major, lane = M row, column = k)
DR_DEVINL void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// dynamic smem (1024 B aligned): [stage][A fp32 16K | B hi BN*128 | B lo BN*128] ... barriers ... tmem ptr
__global__ void __launch_bounds__(TC1_THREADS, 1)
conv_tc_atmem_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_wlo, TcParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.BN * TC_BK * 4;
  const int stage_bytes = A_TILE_BYTES + 2 * b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* split_bar = empty_bar + p.stages;       // A tile split into tensor memory
  uint64_t* acc_full = split_bar + p.stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.ksz * p.ksz * p.kblocks_per_tap;
  const int total_tiles = p.tiles_m * p.tiles_n;
  const uint32_t a_col0 = 2u * (uint32_t)p.acc_stride;             // first tensor-memory column of the A ring (behind the two accumulator stages)
  // two-level accumulation (see conv_tc.cu): chunks of `ch` k-blocks per partial accumulator stage, running sum in the epilogue warps' registers
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
      const uint32_t tx = (uint32_t)(A_TILE_BYTES + 2 * b_bytes);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tile_m = tile / p.tiles_n, n0 = (tile - tile_m * p.tiles_n) * p.BN;
        const int pix0 = tile_m * TC_BM;
        const int img = pix0 / (p.H * p.W);
        const int y0 = (pix0 - img * p.H * p.W) / p.W;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);          // the MMAs that read smem stage s AND tensor-memory A stage s have retired
          const int tap = kb / p.kblocks_per_tap;
          const int c0 = (kb - tap * p.kblocks_per_tap) * TC_BK;
          const int dy = tap / p.ksz - p.pad, dx = tap % p.ksz - p.pad;
          uint8_t* st = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], tx);
          tma_load_4d(&map_a, &full_bar[s], st, c0, dx, y0 + dy, img);
          const int wtap = p.flip_taps ? p.ksz * p.ksz - 1 - tap : tap;
          tma_load_3d(&map_w, &full_bar[s], st + A_TILE_BYTES, c0, n0, wtap);
          tma_load_3d(&map_wlo, &full_bar[s], st + A_TILE_BYTES + b_bytes, c0, n0, wtap);
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
          mbar_wait(&acc_empty[as], aph ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + as * (uint32_t)p.acc_stride;
          const int kb0 = c * ch, kb1 = kb0 + ch < num_kb ? kb0 + ch : num_kb;
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            mbar_wait(&split_bar[s], ph);                // A hi / lo of this stage are in tensor memory, B hi / lo in shared memory
            tc_fence_after();
            const uint32_t b_addr = smem_u32(smem + (size_t)s * stage_bytes) + A_TILE_BYTES;
            const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)s * 64u, a_lo = a_hi + 32u;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
              const uint64_t bd = make_desc(b_addr + k * 32), bld = make_desc(b_addr + b_bytes + k * 32);
              tc_mma_tf32_ts(tmem_d, a_hi + 8u * k, bld, idesc, ((kb - kb0) | k) != 0);   // hi * lo (first MMA of a chunk overwrites)
              tc_mma_tf32_ts(tmem_d, a_lo + 8u * k, bd, idesc, 1);                          // lo * hi
              tc_mma_tf32_ts(tmem_d, a_hi + 8u * k, bd, idesc, 1);                          // hi * hi
            }
            tc_commit(&empty_bar[s]);
          }
          tc_commit(&acc_full[as]);
        }
      }
    }
  }
  } else if (warp >= TC1_WARP_EPI0) {
    DR_SETMAXNREG_INC(REG_EPI);                           // warpgroup 2 (epilogue)
    // ===================== epilogue (shared with conv_tc.cu) =====================
    __shared__ float s_sum[4][256], s_sq[4][256];
    __shared__ __align__(16) float s_scale[TC_MAX_COUT], s_shift[TC_MAX_COUT];
    __shared__ __align__(16) float s_stage[4][32 * TC_STAGE_LD];
    __shared__ int s_last;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = q * 32 + lane;
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
    DR_SETMAXNREG_DEC(REG_SPLIT);                         // warpgroup 1 (splitters)
    // ===================== A splitter: shared memory (fp32, 128B-swizzled rows) -> registers -> tensor memory (hi | lo) =====================
    // warps 4..7: warp % 4 = 0..3 -> each owns one 32-lane quarter of tensor memory = 32 pixel rows of the tile
    const int q = warp & 3;
    const int r = q * 32 + lane;                          // pixel row of the tile == tensor-memory lane
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        const uint8_t* arow = smem + (size_t)s * stage_bytes + (size_t)r * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {                     // SWIZZLE_128B: the 16 B chunk holding k = 4c..4c+3 of row r sits at chunk c ^ (r % 8)
          const float4 a = *reinterpret_cast<const float4*>(arow + ((c ^ (r & 7)) << 4));
          const float h0 = tf32_rna(a.x), h1 = tf32_rna(a.y), h2 = tf32_rna(a.z), h3 = tf32_rna(a.w);
          hi[4 * c] = __float_as_uint(h0); lo[4 * c] = __float_as_uint(tf32_rna(a.x - h0));
          hi[4 * c + 1] = __float_as_uint(h1); lo[4 * c + 1] = __float_as_uint(tf32_rna(a.y - h1));
          hi[4 * c + 2] = __float_as_uint(h2); lo[4 * c + 2] = __float_as_uint(tf32_rna(a.z - h2));
          hi[4 * c + 3] = __float_as_uint(h3); lo[4 * c + 3] = __float_as_uint(tf32_rna(a.w - h3));
        }
        const uint32_t ta = tmem_base + lane_sel + a_col0 + (uint32_t)s * 64u;
        tmem_st32(ta, hi);
        tmem_st32(ta + 32u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();                                // order the tensor-memory stores before the arrival the MMA thread waits on
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
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

// 0 = off (default), 1 = use for 3xTF32 layers the pair kernel does not take, 2 = prefer over the pair kernel
int conv_tc_atmem_mode() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("DENSEREG_TC_A_TMEM"); mode = e ? atoi(e) : 1; }   // default 1 since round 2: -0.43 ms per micro-batch (profiles/r2_sweep.md)
  return mode;
}

// Same contract as launch_conv_tc (conv_tc.cu) with split3 = 1; the caller has checked conv_tc_eligible(p).  Returns 0 if the launch could not be made.
int launch_conv_tc_atmem(const ConvProblem& p, cudaStream_t st) {
  static bool attr_set = false;
  TcParams t; memset(&t, 0, sizeof(t));
  t.M = p.B * p.H * p.W; t.H = p.H; t.W = p.W; t.Cin = p.Cin; t.Cout = p.Cout; t.ksz = p.k; t.pad = p.pad_t; t.flip_taps = p.flip_taps;
  int BN = (p.Cout + 15) / 16 * 16;
  if (BN > 128) BN = 128;                                  // 2*BN accumulator columns + 64 per ring stage <= 512
  t.BN = BN;
  t.kblocks_per_tap = (p.Cin + TC_BK - 1) / TC_BK;
  t.tiles_m = (t.M + TC_BM - 1) / TC_BM; t.tiles_n = (p.Cout + BN - 1) / BN;
  const int stage_bytes = A_TILE_BYTES + 2 * BN * TC_BK * 4;
  static int smem_budget = 0;
  if (!smem_budget) {
    cudaFuncAttributes fa;
    const size_t st_bytes = cudaFuncGetAttributes(&fa, conv_tc_atmem_kernel) == cudaSuccess ? fa.sharedSizeBytes : 34 * 1024;
    smem_budget = 227 * 1024 - (int)st_bytes - 1536;
  }
  int stages = smem_budget / stage_bytes;
  const int acc_stride = (BN + 31) / 32 * 32;
  const int a_col0 = 2 * acc_stride;
  const int tmem_stages = (512 - a_col0) / 64;
  if (stages > tmem_stages) stages = tmem_stages;
  if (stages > 6) stages = 6;
  if (stages < 2) return 0;
  const int num_kb = p.k * p.k * t.kblocks_per_tap;
  if (stages > num_kb) stages = num_kb < 2 ? 2 : num_kb;
  t.stages = stages; t.acc_stride = acc_stride;
  { const int nk = p.k * p.k * t.kblocks_per_tap;
    t.chunk_kb = (p.chunk_kb > 0 && nk > p.chunk_kb && nk > p.chunk_min_kb) ? p.chunk_kb : 0; }
  int cols = 32; while (cols < a_col0 + 64 * stages) cols <<= 1;
  t.tmem_cols = cols;
  t.y = p.y; t.y_cs = p.y_cs; t.scale = p.scale; t.shift = p.shift; t.relu = p.relu; t.res = p.res; t.res_cs = p.res_cs;
  t.accumulate = p.accumulate; t.dropout = p.dropout; t.drop_seed = p.drop_seed; t.drop_tag = p.drop_tag;
  t.stats = p.stats; t.stats_counter = p.stats_counter; t.bn_bg = p.bn_bg; t.bn_state = p.bn_state; t.bn_aff = p.bn_aff; t.bn_bstat = p.bn_bstat;
  t.bn_update_state = p.bn_update_state;
  { static int co = -1; if (co < 0) { const char* e = getenv("DENSEREG_TC_EPI_COALESCE"); co = (e && e[0] == '0') ? 0 : 1; } t.coalesce = co; }
  { static int pc = -1; if (pc < 0) { const char* e = getenv("DENSEREG_TC_STATS_PER_CTA"); pc = (e && e[0] == '0') ? 0 : 1; }   // default on: measured -0.12 ms per micro-batch (profiles/r2_sweep.md)
    t.stats_per_cta = (pc && p.stats && !p.scale && !p.shift) ? 1 : 0; }
  const size_t smem_bytes = (size_t)stages * stage_bytes + (3 * stages + 4) * 8 + 16 + 1024 + 64;

  CUtensorMap ma, mw, mwlo;
  const int rows = TC_BM / p.W;
  const int bh = rows < p.H ? rows : p.H;
  const int bb = rows < p.H ? 1 : rows / p.H;
  cuuint64_t ad[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  cuuint64_t as[3] = {(cuuint64_t)p.x_cs * 4, (cuuint64_t)p.W * p.x_cs * 4, (cuuint64_t)p.H * p.W * p.x_cs * 4};
  cuuint32_t ab[4] = {(cuuint32_t)TC_BK, (cuuint32_t)p.W, (cuuint32_t)bh, (cuuint32_t)bb};
  if (!tc::encode_map(&ma, p.x, 4, ad, as, ab)) return 0;
  cuuint64_t wd[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Cout, (cuuint64_t)(p.k * p.k)};
  cuuint64_t ws[2] = {(cuuint64_t)p.wk_ld * 4, (cuuint64_t)p.wk_ld * p.Cout * 4};
  cuuint32_t wb[3] = {(cuuint32_t)TC_BK, (cuuint32_t)BN, 1};
  if (!tc::encode_map(&mw, p.w_kmajor, 3, wd, ws, wb)) return 0;
  if (!tc::encode_map(&mwlo, p.w_kmajor_lo, 3, wd, ws, wb)) return 0;

  static int num_sms = 0;
  if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
  const int total_tiles = t.tiles_m * t.tiles_n;
  dim3 grid(total_tiles < num_sms ? total_tiles : num_sms);
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_tc_atmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_budget + 1536) != cudaSuccess) return 0;
    attr_set = true;
  }
  dr_launch(conv_tc_atmem_kernel, dim3(grid), dim3(TC1_THREADS), smem_bytes, st, ma, mw, mwlo, t);
  return 1;
}
