// conv_tc_pair.cu -- CTA-pair (cta_group::2) variant of the 3xTF32 implicit-GEMM convolution of conv_tc.cu.
//
// Why: profiles/r1_tensor_core_path.md shows the one-CTA 3xTF32 kernel with the tensor pipe ~52 % active and neither HBM nor L2
// near their limits.  The limiter is SHARED-MEMORY bandwidth: per 32-channel k-block at BN=128 the SM moves 96 KB of UMMA operand
// reads (12 MMAs x (4 KB A + 4 KB B)) + 48 KB of TMA writes + 48 KB of splitter traffic = 192 KB against 800 MMA cycles, i.e.
// 240 B/clk demanded of a 128 B/clk shared memory.  A CTA pair computes D[256 pixels, BN<=256 couts] with ONE tcgen05.mma.cta_group::2
// per k-step: each SM still reads its own 128-row A tile, but only HALF of the B tile (the other half is supplied by the peer SM),
// and TMA fetches each weight tile once per 256 pixels instead of once per 128.  At BN=256: 96 KB operand reads + 48 KB TMA +
// 48 KB splitter per 1600 MMA cycles = 120 B/clk -- inside the budget -- and the L2->SM traffic per FLOP halves.
//
// Protocol (rank 0 = leader, rank 1 = peer; both CTAs run all roles except the MMA issue):
//   producer (warp 0)    waits its LOCAL empty[s], TMA-loads its own A tile (128 pixels x 32 ch) and its own half of the pre-split
//                        weights (BN/2 couts x 32 ch, hi and lo) onto its LOCAL full[s]
//   splitters (4 warps)  wait LOCAL full[s], rewrite A into hi/lo in place, fence.proxy.async, arrive on the LEADER's split[s]
//                        (8 arrivals: 4 warps x 2 CTAs) -- this also tells the leader that the peer's B half landed
//   MMA (leader, 1 thr)  waits split[s], issues 12 x tcgen05.mma.cta_group::2 (M=256, N=BN, K=8: hi*lo, lo*hi, hi*hi),
//                        tcgen05.commit.cta_group::2 .multicast -> empty[s] of BOTH CTAs; after the last k-block -> acc_full[as] of both
//   epilogue (4 warps)   each CTA drains its own 128 TMEM lanes (shared code: tc_epilogue_tile) and arrives on the LEADER's
//                        acc_empty[as] (8 arrivals)
// Accumulators are double-buffered in TMEM (2 x BN columns, allocated with cta_group::2 by warp 1 of both CTAs).
//
// Taken for Cout >= 128 and >= 64 work items (conv_tc_pair_wanted); off with dr_config.reserved[1] < 0 or DENSEREG_TC_PAIR=0.  DESIGN.md 4.1b.
#include "conv_tc_epilogue.cuh"
#include <stdlib.h>

namespace {

using namespace tcconv;

DR_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DR_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// shared::cta address of THIS CTA -> shared::cluster address of the same offset in CTA `rank`
DR_DEVINL uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Remote arrive with the DEFAULT semantics (release at CTA scope), as cutlass::arch::ClusterBarrier::arrive(cta_id) does.  The data the
// arrival publishes is this CTA's own shared memory, made visible to the tensor cores by fence.proxy.async; the arrival itself is only a
// signal.  (A .release.cluster arrive compiles to MEMBAR.ALL.GPU + ERRBAR per call -- measured: ~2400 cycles per k-block floor.)
DR_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
DR_DEVINL void tc_commit_pair(uint64_t* bar) {      // arrives on `bar` at the same offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
DR_DEVINL void tc_mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// dynamic smem (1024 B aligned, identical in both CTAs): [stage][A hi 16K | A lo 16K | B hi (BN/2)*128 | B lo (BN/2)*128] ... barriers ... tmem ptr
// Work item = (pixel-tile PAIR, n-tile); cluster c walks items c, c + #clusters, ...; CTA rank r owns pixel tile 2*pair + r.
__global__ void __launch_bounds__(192 + SPLIT_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_wlo, TcParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int bh_bytes = (p.BN / 2) * TC_BK * 4;                   // this CTA's half of one weight tile (hi or lo)
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * bh_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);   // local: TMA landed
  uint64_t* empty_bar = full_bar + p.stages;        // local: slot free (multicast commit of the leader)
  uint64_t* split_bar = empty_bar + p.stages;       // LEADER's copy is used: both CTAs' A tiles split, both B halves landed
  uint64_t* acc_full = split_bar + p.stages;        // [2] local: accumulator stage complete (multicast commit)
  uint64_t* acc_empty = acc_full + 2;               // [2] LEADER's copy is used: both epilogues drained the stage
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int num_kb = p.ksz * p.ksz * p.kblocks_per_tap;
  const int pairs_m = (p.tiles_m + 1) >> 1;
  const int total_items = pairs_m * p.tiles_n;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  struct PItem { int pair, n0, bn; };
  auto decode = [&](int item) {
    PItem w;
    w.pair = item / p.tiles_n; w.n0 = (item - w.pair * p.tiles_n) * p.BN; w.bn = p.BN;
    return w;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], 2 * (SPLIT_THREADS / 32));
    }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // one warp of EACH CTA (same warp id) allocates with cta_group::2
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();                      // everything above overlapped the previous kernel's tail (common.cuh)

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters) {
        const PItem w = decode(item);
        const uint32_t tx = (uint32_t)(A_TILE_BYTES + 2 * (w.bn / 2) * TC_BK * 4);
        const int pix0 = (w.pair * 2 + (int)rank) * TC_BM;        // may lie past M for the odd tail: TMA zero-fills, the epilogue masks
        const int img = pix0 / (p.H * p.W);
        const int y0 = (pix0 - img * p.H * p.W) / p.W;
        const int nb0 = w.n0 + (int)rank * (w.bn / 2);            // this CTA's half of the cout tile (slice)
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
          uint8_t* bdst = st + 2 * A_TILE_BYTES;
          const int wtap = p.flip_taps ? p.ksz * p.ksz - 1 - tap : tap;
          tma_load_3d(&map_w, &full_bar[s], bdst, c0, nb0, wtap);
          tma_load_3d(&map_wlo, &full_bar[s], bdst + bh_bytes, c0, nb0, wtap);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (rank == 0 && lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int item = cluster_id; item < total_items; item += num_clusters, ++tcount) {
        const PItem w = decode(item);
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(w.bn >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
        const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);           // both epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * (uint32_t)p.BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&split_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
          const uint32_t b_addr = a_addr + 2 * A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t ad = make_desc(a_addr + k * 32), bd = make_desc(b_addr + k * 32);
            const uint64_t ald = make_desc(a_addr + A_TILE_BYTES + k * 32), bld = make_desc(b_addr + bh_bytes + k * 32);
            tc_mma_tf32_pair(tmem_d, ad, bld, idesc, (kb | k) != 0);    // hi * lo
            tc_mma_tf32_pair(tmem_d, ald, bd, idesc, 1);                // lo * hi
            tc_mma_tf32_pair(tmem_d, ad, bd, idesc, 1);                 // hi * hi
          }
          tc_commit_pair(&empty_bar[s]);         // frees the smem slot in both CTAs when these MMAs retire
        }
        tc_commit_pair(&acc_full[as]);           // accumulator complete -> both epilogues
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    __shared__ float s_sum[4][256], s_sq[4][256];
    __shared__ __align__(16) float s_scale[TC_MAX_COUT], s_shift[TC_MAX_COUT];
    __shared__ __align__(16) float s_stage[4][32 * TC_STAGE_LD];        // per-warp 32x32 transpose tile (coalesced epilogue)
    __shared__ int s_last;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = q * 32 + lane;
    const bool vec_ok = ((p.y_cs & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0) &&
                        (!p.res || (((p.res_cs & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.res) & 15) == 0)));
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(&acc_empty[0]), 0);
    const int total_cta_tiles = 2 * total_items;                  // every CTA tile (also the phantom one of an odd tail) counts once
    tc_epilogue_stage_affine(p, et, s_scale, s_shift);
    uint32_t tcount = 0;
    for (int item = cluster_id; item < total_items; item += num_clusters, ++tcount) {
      const PItem w = decode(item);
      const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
      mbar_wait_sleep(&acc_full[as], aph);
      tc_fence_after();
      tc_epilogue_tile(p, tmem_base + as * (uint32_t)p.BN, q, lane, row, et, vec_ok, w.pair * 2 + (int)rank, w.n0, w.bn, total_cta_tiles, s_sum, s_sq,
                       s_last, s_scale, s_shift, s_stage[q], [&]() { mbar_arrive_cluster(acc_empty_leader + as * 8u); });
    }
    tc_epilogue_finish(p, et, s_scale, s_shift, s_last);
  } else {
    // ===================== A splitter (both CTAs): hi = rn_tf32(a), lo = rn_tf32(a - hi) =====================
    const int t = threadIdx.x - 192;
    const int na4 = A_TILE_BYTES / 16;
    const uint32_t split_leader = mapa_u32(smem_u32(&split_bar[0]), 0);
    uint32_t it = 0;
    for (int item = cluster_id; item < total_items; item += num_clusters) {
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        float4* a_hi = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
        float4* a_lo = a_hi + na4;
        for (int idx = t; idx < na4; idx += SPLIT_THREADS) {
          const float4 a = a_hi[idx];
          float4 h, l;
          h.x = tf32_rna(a.x); l.x = tf32_rna(a.x - h.x);
          h.y = tf32_rna(a.y); l.y = tf32_rna(a.y - h.y);
          h.z = tf32_rna(a.z); l.z = tf32_rna(a.z - h.z);
          h.w = tf32_rna(a.w); l.w = tf32_rna(a.w - h.w);
          a_hi[idx] = h; a_lo[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor cores (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(split_leader + (uint32_t)s * 8u);
      }
    }
  }

  // the peer's shared memory and barriers must stay alive until the leader's last MMA / multicast commit has retired
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

}  // namespace

// Pair path is taken for the big layers only: wide enough that the shared B tile pays (Cout >= 128) and enough 256-pixel work items
// to fill the machine with CTA pairs.
bool conv_tc_pair_wanted(const ConvProblem& p) {
  if (!p.pair || !p.w_kmajor_lo) return false;
  if (p.pair >= 2) return p.Cout >= 16;                    // forced (dr_debug_conv flag 0x200): any shape the pair kernel can run
  static int min_cout = -1;       // DENSEREG_TC_PAIR_MINCOUT: narrowest layer that takes the pair kernel (default 128)
  if (min_cout < 0) { const char* e = getenv("DENSEREG_TC_PAIR_MINCOUT"); min_cout = e ? atoi(e) : 128; if (min_cout < 32) min_cout = 32; }
  if (p.Cout < min_cout) return false;
  const int M = p.B * p.H * p.W;
  int BN = (p.Cout + 15) / 16 * 16; if (BN > 256) BN = 256;
  const int items = ((M + 2 * TC_BM - 1) / (2 * TC_BM)) * ((p.Cout + BN - 1) / BN);
  // DENSEREG_TC_PAIR_MINWORK (k-blocks x BN) can keep short reductions on the one-CTA kernel; measured on the whole training step
  // (B=40): threshold 0 -> 1845 crops/s, 4096 -> 1837, 8192 -> 1797, so every Cout >= 128 layer with enough work items takes the pair kernel
  static int min_work = -1;
  if (min_work < 0) { const char* e = getenv("DENSEREG_TC_PAIR_MINWORK"); min_work = e ? atoi(e) : 0; }
  const int num_kb = p.k * p.k * ((p.Cin + TC_BK - 1) / TC_BK);
  return items >= 64 && num_kb * BN >= min_work;
}

// Same contract as launch_conv_tc (conv_tc.cu) with split3 = 1; the caller has checked conv_tc_eligible(p).
int launch_conv_tc_pair(const ConvProblem& p, cudaStream_t st) {
  static bool attr_set = false;
  TcParams t; memset(&t, 0, sizeof(t));
  t.M = p.B * p.H * p.W; t.H = p.H; t.W = p.W; t.Cin = p.Cin; t.Cout = p.Cout; t.ksz = p.k; t.pad = p.pad_t; t.flip_taps = p.flip_taps;
  int BN = (p.Cout + 15) / 16 * 16;
  if (BN > 256) BN = 256;
  t.BN = BN;
  t.kblocks_per_tap = (p.Cin + TC_BK - 1) / TC_BK;
  int cols = 32; while (cols < 2 * BN) cols <<= 1;         // two accumulator stages of BN columns in each SM's TMEM
  t.tmem_cols = cols;
  t.tiles_m = (t.M + TC_BM - 1) / TC_BM; t.tiles_n = (p.Cout + BN - 1) / BN;
  const int stage_bytes = 2 * A_TILE_BYTES + 2 * (BN / 2) * TC_BK * 4;
  static int smem_budget = 0;
  if (!smem_budget) {
    cudaFuncAttributes fa;
    const size_t st_bytes = cudaFuncGetAttributes(&fa, conv_tc_pair_kernel) == cudaSuccess ? fa.sharedSizeBytes : 16 * 1024;
    smem_budget = 227 * 1024 - (int)st_bytes - 1536;     // 1536 >= barriers + tmem slot + 1024 B alignment slack
  }
  int stages = smem_budget / stage_bytes;
  if (stages > 6) stages = 6;
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
  cuuint32_t wb[3] = {(cuuint32_t)TC_BK, (cuuint32_t)(BN / 2), 1};       // each CTA of the pair loads half of the cout tile
  if (!tc::encode_map(&mw, p.w_kmajor, 3, wd, ws, wb)) return 0;
  if (!tc::encode_map(&mwlo, p.w_kmajor_lo, 3, wd, ws, wb)) return 0;

  static int num_sms = 0;
  if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); if (num_sms <= 0) num_sms = 148; }
  const int items = ((t.tiles_m + 1) / 2) * t.tiles_n;
  const int clusters = items < num_sms / 2 ? items : num_sms / 2;
  if (!attr_set) {
    if (!tc::launch_ok(cudaFuncSetAttribute(conv_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_budget + 1536), "conv_tc_pair_kernel smem attribute")) return 0;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(192 + SPLIT_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = dr_pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 2;
  return tc::launch_ok(cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel, ma, mw, mwlo, t), "conv_tc_pair_kernel") ? 1 : 0;
}
