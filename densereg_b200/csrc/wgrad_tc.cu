// wgrad_tc.cu -- filter gradient on the tensor cores (sm_100a): dW[tap][cin][cout] += sum_pixels X_tap[pix][cin]^T * dY[pix][cout]
// as a tcgen05 GEMM with BOTH operands MN-major (the reduction dimension -- pixels -- is the slow-moving index of the
// NHWC tensors, so no transposition pass is needed).
//
// Replaces TF's Conv2DBackpropFilter for the stride-1 1x1 / 3x3 convs of network/um_v1.py (one third of the training FLOPs).
//
// Per CTA: D[128 cin, BN couts] (fp32 in TMEM) accumulates over a contiguous range of 32-pixel k-blocks of ONE filter tap:
//   A = X^T : 4 TMA boxes (32 ch, bw, bh, bb) -- one per 32-channel chunk -- of the forward input at the tap's spatial offset
//             (out-of-image pixels are TMA zero-fill = SAME padding); smem [chunk][32 pixels][32 floats], 128B-span swizzle with 32 B atoms
//             (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B == UMMA SWIZZLE_128B_BASE32B, the only MN-major layout for 32-bit operands).
//   B = dY  : BN/32 boxes of d(raw conv output), unshifted, same layout.
//   4 MMAs (M=128, N=BN, K=8 pixels, kind::tf32, a_major = b_major = MN) per k-block; UMMA descriptors: LBO = 4096 B between
//   32-wide chunks, SBO = 512 B between 4-pixel groups, start advanced by 1024 B per K=8 step.
//   3xTF32 (DR_PREC_TF32X3): 3 MMAs per k-step (hi*lo, lo*hi, hi*hi).  The tensor core reads sign, exponent and the top 10 mantissa bits of a
//   32-bit kind::tf32 operand, i.e. it TRUNCATES: the landed fp32 stage itself is the hi operand, and the eight splitter warps only write
//   lo = rn_tf32(v - trunc(v)) (exact difference, |lo| < 2^-10 |v|, representation error <= 2^-21 |v|, unbiased) next to it -- half the
//   shared-memory writes of a split that also rewrites hi (DENSEREG_SPLIT_TRUNC=0 restores hi = rn_tf32(v), lo = rn_tf32(v - hi) in place).
//   What the 3-product scheme drops is lo*lo <= 2^-20 |a b| with the sign of a*b (2^-22 on average: a uniform relative shrink of 2.4e-7).
// Epilogue: tcgen05.ld -> red.global.add.f32 into the flat gradient (split over pixel ranges across blockIdx.z, and micro-batch
// accumulation, are both just "+=").
#include "tc_common.cuh"
#include <stdlib.h>

using namespace tc;

namespace {

constexpr int WG_KB = 32;                         // pixels per k-block
constexpr int WG_CHUNK_BYTES = WG_KB * 128;       // one 32-channel chunk of one k-block: 4 KB
constexpr int WG_A_BYTES = 4 * WG_CHUNK_BYTES;    // 128 cin
constexpr int WG_SPLIT_THREADS = 256;             // 8 splitter warps (3xTF32)

struct WgParams {
  int H, W, Cin, Cout, ksz, pad;
  int BN, nchunks_b;
  int total_kb, kb_per_split;
  int cin_tiles;
  int stages, tmem_cols;
  int bw, bh, bb;
  float* dw;
  int n_tiles, splits;   // persistent kernel: work item = (split, tap, m tile, n tile), n fastest
  int swap;              // 1: roles exchanged -- M = cout tile (A = dY, unshifted), N = cin tile (B = X at the tap's offset); see launch_wgrad_tc
  int trunc_hi;          // 3xTF32: 1 = the landed fp32 tile IS the hi operand (the tensor core ignores the low 13 mantissa bits of a kind::tf32
                         // operand = truncation), the splitters only write lo = rn_tf32(v - trunc(v)): half the splitter's shared-memory writes
};

template <bool SPLIT3>
__global__ void __launch_bounds__(SPLIT3 ? 192 + WG_SPLIT_THREADS : 192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, WgParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.nchunks_b * WG_CHUNK_BYTES;
  const int stage_bytes = (SPLIT3 ? 2 : 1) * (WG_A_BYTES + b_bytes);      // [A | B] (+ [A_lo | B_lo])
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* split_bar = empty_bar + p.stages;
  uint64_t* accum_bar = split_bar + p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = blockIdx.x / p.cin_tiles;                          // cin_tiles = number of 128-row M tiles (cin tiles, or cout tiles when swapped)
  const int c0 = (blockIdx.x - tap * p.cin_tiles) * 128;             // first M row: cin (cout when swapped)
  const int n0 = blockIdx.y * p.BN;                                  // first N column: cout (cin when swapped)
  const int kb_begin = blockIdx.z * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.total_kb) kb_end = p.total_kb;
  const int num_kb = kb_end - kb_begin;           // >= 1 by construction

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], WG_SPLIT_THREADS / 32); }
    mbar_init(accum_bar, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      const int dy = tap / p.ksz - p.pad, dx = tap % p.ksz - p.pad;
      const uint32_t tx = (uint32_t)(WG_A_BYTES + b_bytes);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int pix = (kb_begin + i) * WG_KB;
        const int img = pix / (p.H * p.W);
        const int rem = pix - img * p.H * p.W;
        const int y = rem / p.W, x = rem - y * p.W;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], tx);
        if (!p.swap) {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(&map_x, &full_bar[s], st + j * WG_CHUNK_BYTES, c0 + 32 * j, x + dx, y + dy, img);
          for (int j = 0; j < p.nchunks_b; ++j)
            tma_load_4d(&map_dy, &full_bar[s], st + WG_A_BYTES + j * WG_CHUNK_BYTES, n0 + 32 * j, x, y, img);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(&map_dy, &full_bar[s], st + j * WG_CHUNK_BYTES, c0 + 32 * j, x, y, img);
          for (int j = 0; j < p.nchunks_b; ++j)
            tma_load_4d(&map_x, &full_bar[s], st + WG_A_BYTES + j * WG_CHUNK_BYTES, n0 + 32 * j, x + dx, y + dy, img);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // a_format = b_format = TF32, c = F32, a_major = b_major = MN (bits 15, 16)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1;
        if (SPLIT3) mbar_wait(&split_bar[s], ph); else mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + WG_A_BYTES;
        const uint32_t lo_off = (uint32_t)(WG_A_BYTES + b_bytes);
#pragma unroll
        for (int k = 0; k < WG_KB / 8; ++k) {
          const uint64_t ad = make_desc_mn(a_addr + k * 1024, WG_CHUNK_BYTES, 512);
          const uint64_t bd = make_desc_mn(b_addr + k * 1024, WG_CHUNK_BYTES, 512);
          if (SPLIT3) {
            const uint64_t ald = make_desc_mn(a_addr + lo_off + k * 1024, WG_CHUNK_BYTES, 512);
            const uint64_t bld = make_desc_mn(b_addr + lo_off + k * 1024, WG_CHUNK_BYTES, 512);
            tc_mma_tf32(tmem_base, ad, bld, idesc, (i | k) != 0);
            tc_mma_tf32(tmem_base, ald, bd, idesc, 1);
            tc_mma_tf32(tmem_base, ad, bd, idesc, 1);
          } else {
            tc_mma_tf32(tmem_base, ad, bd, idesc, (i | k) != 0);
          }
        }
        tc_commit(&empty_bar[s]);
      }
      tc_commit(accum_bar);
    }
  } else if (warp < 6) {
    mbar_wait_sleep(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int c = c0 + q * 32 + lane;                       // TMEM lane == M row of the tile: cin (cout when swapped)
    if (!p.swap) {
      const bool cvalid = c < p.Cin;
      float* row = p.dw + ((size_t)tap * p.Cin + c) * p.Cout;
      for (int cb = 0; cb < p.BN; cb += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
        if (!cvalid) continue;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int n = n0 + cb + e;
          if (n < p.Cout) atomicAdd(row + n, __uint_as_float(v[e]));
        }
      }
    } else {
      // lane = cout, column = cin: dw[(tap*Cin + cin)*Cout + cout] -> the 32 lanes of one reduction hit 32 consecutive floats (one 128 B line)
      const bool cvalid = c < p.Cout;
      float* col = p.dw + (size_t)tap * p.Cin * p.Cout + c;
      for (int cb = 0; cb < p.BN; cb += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
        if (!cvalid) continue;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int n = n0 + cb + e;
          if (n < p.Cin) atomicAdd(col + (size_t)n * p.Cout, __uint_as_float(v[e]));
        }
      }
    }
    tc_fence_before();
  } else if (SPLIT3) {
    const int t = threadIdx.x - 192;
    const int n16 = (WG_A_BYTES + b_bytes) / 16;             // float4 elements of [A | B]
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % p.stages;
      const uint32_t ph = (uint32_t)(i / p.stages) & 1;
      mbar_wait(&full_bar[s], ph);
      float4* hi = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes);
      float4* lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + WG_A_BYTES + b_bytes);
      if (p.trunc_hi) {
        for (int idx = t; idx < n16; idx += WG_SPLIT_THREADS) {
          const float4 a = hi[idx];
          float4 l;
          l.x = tf32_rna(a.x - tf32_trunc(a.x)); l.y = tf32_rna(a.y - tf32_trunc(a.y));
          l.z = tf32_rna(a.z - tf32_trunc(a.z)); l.w = tf32_rna(a.w - tf32_trunc(a.w));
          lo[idx] = l;
        }
      } else {
        for (int idx = t; idx < n16; idx += WG_SPLIT_THREADS) {
          float4 a = hi[idx], h, l;
          h.x = tf32_rna(a.x); l.x = tf32_rna(a.x - h.x);
          h.y = tf32_rna(a.y); l.y = tf32_rna(a.y - h.y);
          h.z = tf32_rna(a.z); l.z = tf32_rna(a.z - h.z);
          h.w = tf32_rna(a.w); l.w = tf32_rna(a.w - h.w);
          hi[idx] = h; lo[idx] = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&split_bar[s]);
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}


// ---- persistent 3xTF32 variant with the split A operand in TENSOR MEMORY (opt-in, DENSEREG_WGRAD_A_TMEM=1) ----------------------------------
// Persistent (one CTA per SM walks (split, tap, m tile, n tile) work items, shared-memory ring across items, two accumulator stages in tensor
// memory so that the reductions of item i overlap the main loop of item i+1); the A tile (X^T, or dY^T when swapped) is split into tensor memory instead of shared memory: splitter
// warp w owns tensor-memory lane quarter w % 4 = 32-channel chunk w % 4 of the tile (smem [chunk][pixel][32 channels]); for every pixel the
// warp reads that pixel's 128 B row (conflict-free), so lane i collects channel i over the pixels = one K-major A row; hi / lo go out with
// tcgen05.st and the MMAs use the [d], [a-tmem], b-desc form (A from tensor memory is K-major; B = the other operand stays MN-major in
// shared memory and is split in place as before).  Two warps share a quarter and take 16 pixels (columns) each.
// Shared-memory bytes per k-block at BN = 128: 32 KB TMA + 16 KB A reads + 48 KB B split + 48 KB B operand reads = 144 KB (224 KB before).
DR_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
DR_DEVINL void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(192 + WG_SPLIT_THREADS, 1)
wgrad_tc_atmem_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, WgParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = p.nchunks_b * WG_CHUNK_BYTES;
  const int stage_bytes = WG_A_BYTES + 2 * b_bytes;                   // [A fp32 | B hi | B lo]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* split_bar = empty_bar + p.stages;
  uint64_t* acc_full = split_bar + p.stages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.ksz * p.ksz * p.cin_tiles * p.n_tiles;
  const int total_items = tiles * p.splits;
  const uint32_t a_col0 = (2u * (uint32_t)p.BN + 31u) & ~31u;         // tensor-memory A ring behind the two accumulator stages

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], WG_SPLIT_THREADS / 32); }
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

  struct Item { int tap, c0, n0, kb_begin, num_kb; };
  auto decode = [&](int item) {
    Item it;
    const int split = item / tiles, tile = item - split * tiles;
    const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
    it.tap = mt / p.cin_tiles;
    it.c0 = (mt - it.tap * p.cin_tiles) * 128;
    it.n0 = nt * p.BN;
    it.kb_begin = split * p.kb_per_split;
    int kb_end = it.kb_begin + p.kb_per_split;
    if (kb_end > p.total_kb) kb_end = p.total_kb;
    it.num_kb = kb_end - it.kb_begin;
    return it;
  };

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = (uint32_t)(WG_A_BYTES + b_bytes);
      uint32_t ring = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const Item w = decode(item);
        const int dy = w.tap / p.ksz - p.pad, dx = w.tap % p.ksz - p.pad;
        for (int i = 0; i < w.num_kb; ++i, ++ring) {
          const int s = ring % p.stages;
          const uint32_t ph = (ring / p.stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const int pix = (w.kb_begin + i) * WG_KB;
          const int img = pix / (p.H * p.W);
          const int rem = pix - img * p.H * p.W;
          const int y = rem / p.W, x = rem - y * p.W;
          uint8_t* st = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], tx);
          if (!p.swap) {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_4d(&map_x, &full_bar[s], st + j * WG_CHUNK_BYTES, w.c0 + 32 * j, x + dx, y + dy, img);
            for (int j = 0; j < p.nchunks_b; ++j)
              tma_load_4d(&map_dy, &full_bar[s], st + WG_A_BYTES + j * WG_CHUNK_BYTES, w.n0 + 32 * j, x, y, img);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_4d(&map_dy, &full_bar[s], st + j * WG_CHUNK_BYTES, w.c0 + 32 * j, x, y, img);
            for (int j = 0; j < p.nchunks_b; ++j)
              tma_load_4d(&map_x, &full_bar[s], st + WG_A_BYTES + j * WG_CHUNK_BYTES, w.n0 + 32 * j, x + dx, y + dy, img);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // a_format = b_format = TF32, c = F32, a_major = K (A from tensor memory), b_major = MN (bit 16)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t ring = 0, tcount = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++tcount) {
        const Item w = decode(item);
        const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&acc_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * (uint32_t)p.BN;
        for (int i = 0; i < w.num_kb; ++i, ++ring) {
          const int s = ring % p.stages;
          const uint32_t ph = (ring / p.stages) & 1;
          mbar_wait(&split_bar[s], ph);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(smem + (size_t)s * stage_bytes) + WG_A_BYTES;
          const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)s * 64u, a_lo = a_hi + 32u;
#pragma unroll
          for (int k = 0; k < WG_KB / 8; ++k) {
            const uint64_t bd = make_desc_mn(b_addr + k * 1024, WG_CHUNK_BYTES, 512);
            const uint64_t bld = make_desc_mn(b_addr + b_bytes + k * 1024, WG_CHUNK_BYTES, 512);
            tc_mma_tf32_ts(tmem_d, a_hi + 8u * k, bld, idesc, (i | k) != 0);
            tc_mma_tf32_ts(tmem_d, a_lo + 8u * k, bd, idesc, 1);
            tc_mma_tf32_ts(tmem_d, a_hi + 8u * k, bd, idesc, 1);
          }
          tc_commit(&empty_bar[s]);
        }
        tc_commit(&acc_full[as]);
      }
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++tcount) {
      const Item w = decode(item);
      const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
      mbar_wait_sleep(&acc_full[as], aph);
      tc_fence_after();
      const int c = w.c0 + q * 32 + lane;
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + as * (uint32_t)p.BN;
      if (!p.swap) {
        const bool cvalid = c < p.Cin;
        float* row = p.dw + ((size_t)w.tap * p.Cin + c) * p.Cout;
        for (int cb = 0; cb < p.BN; cb += 32) {
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)cb, v);
          if (!cvalid) continue;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int n = w.n0 + cb + e;
            if (n < p.Cout) atomicAdd(row + n, __uint_as_float(v[e]));
          }
        }
      } else {
        const bool cvalid = c < p.Cout;
        float* col = p.dw + (size_t)w.tap * p.Cin * p.Cout + c;
        for (int cb = 0; cb < p.BN; cb += 32) {
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)cb, v);
          if (!cvalid) continue;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int n = w.n0 + cb + e;
            if (n < p.Cin) atomicAdd(col + (size_t)n * p.Cout, __uint_as_float(v[e]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  } else {
    // splitter warps 6..13: A -> tensor memory (quarter = warp % 4, pixel half = (warp - 6) / 4), B -> hi / lo in shared memory (all 256 threads)
    const int t = threadIdx.x - 192;
    const int q = warp & 3, half = (warp - 6) >> 2;
    const int nb16 = b_bytes / 16;
    uint32_t ring = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const Item w = decode(item);
      for (int i = 0; i < w.num_kb; ++i, ++ring) {
        const int s = ring % p.stages;
        const uint32_t ph = (ring / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        // A: chunk q = 32 channels (lanes), pixels 16*half .. +15 (columns).  SWIZZLE_128B_ATOM_32B: within a 128 B pixel row the 32 B
        // segment index is XORed with ((pixel >> 1) & 3)  [cute Layout_MN_SW128_32B_Atom: Swizzle<2,5,2> on bytes]
        const uint8_t* achunk = st + (size_t)q * WG_CHUNK_BYTES;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) {
          const int pix = half * 16 + pp;
          const uint32_t off = (uint32_t)pix * 128u + (uint32_t)lane * 4u;
          const uint32_t swz = off ^ ((off >> 2) & 0x60u);              // byte bits [5,6] ^= byte bits [7,8]
          const float a = *reinterpret_cast<const float*>(achunk + swz);
          const float h = tf32_rna(a);
          hi[pp] = __float_as_uint(h); lo[pp] = __float_as_uint(tf32_rna(a - h));
        }
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + a_col0 + (uint32_t)s * 64u + (uint32_t)half * 16u;
        tmem_st16(ta, hi);
        tmem_st16(ta + 32u, lo);
        // B: elementwise split in place (hi) and into the lo copy
        float4* bhi = reinterpret_cast<float4*>(st + WG_A_BYTES);
        float4* blo = reinterpret_cast<float4*>(st + WG_A_BYTES + b_bytes);
        for (int idx = t; idx < nb16; idx += WG_SPLIT_THREADS) {
          float4 a = bhi[idx], h, l;
          h.x = tf32_rna(a.x); l.x = tf32_rna(a.x - h.x);
          h.y = tf32_rna(a.y); l.y = tf32_rna(a.y - h.y);
          h.z = tf32_rna(a.z); l.z = tf32_rna(a.z - h.z);
          h.w = tf32_rna(a.w); l.w = tf32_rna(a.w - h.w);
          bhi[idx] = h; blo[idx] = l;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
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

// ---- CTA-pair (cta_group::2) 3xTF32 variant for the wide layers (M = max(cin, cout) >= 256) ---------------------------------------------
// Why: with both operands split in shared memory a one-CTA 128 x 128 tile moves 192 KB of shared-memory traffic per 32-pixel k-block
// (32 KB TMA, 32 KB read + 32 KB written by the splitters, 12 MMAs x 8 KB operand reads) against 768 MMA cycles = 250 B/clk asked of a
// 128 B/clk shared memory (ncu: tensor pipe 40 % active on um_comb/c2).  A CTA pair computes D[256 rows, BN <= 256 columns] with ONE
// tcgen05.mma.cta_group::2 per k-step: each SM supplies its own 128-row A tile and only HALF of the B tile, and holds 128 x BN of the
// accumulator -- at BN = 256 the same 192 KB per k-block now feed 1536 MMA cycles (125 B/clk).
// Protocol = conv_tc_pair.cu (rank 0 = leader): both CTAs run producer / splitters / epilogue on their own shared memory and tensor
// memory lanes; the splitters of both CTAs arrive on the LEADER's split[s]; the leader's single MMA thread issues the 12 MMAs per
// k-block and frees stage s in both CTAs with a multicast commit; after the last k-block a multicast commit releases both epilogues.
// Splitters always use the truncation form (the landed fp32 tile is the hi operand, only lo is written): stage = [A | Bh | A_lo | Bh_lo].
DR_DEVINL uint32_t wg_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
DR_DEVINL void wg_cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
DR_DEVINL uint32_t wg_mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
DR_DEVINL void wg_mbar_arrive_cluster(uint32_t cluster_addr) {      // default (CTA-scope) semantics: see conv_tc_pair.cu
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
DR_DEVINL void wg_commit_pair(uint64_t* bar) {                      // arrives on `bar` at the same offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
DR_DEVINL void wg_mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// grid (2 * taps * m_pairs, n_tiles, splits), cluster (2,1,1); p.cin_tiles = number of 256-row M pairs, p.nchunks_b = 32-wide chunks of HALF a B tile
__global__ void __launch_bounds__(192 + WG_SPLIT_THREADS, 1)
wgrad_tc_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, WgParams p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int bh_bytes = p.nchunks_b * WG_CHUNK_BYTES;                  // this CTA's half of the B tile
  const int half_bytes = WG_A_BYTES + bh_bytes;                       // [A | Bh] as landed (= the hi operands)
  const int stage_bytes = 2 * half_bytes;                             // + [A_lo | Bh_lo]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);   // local: TMA landed
  uint64_t* empty_bar = full_bar + p.stages;                          // local: slot free (multicast commit of the leader)
  uint64_t* split_bar = empty_bar + p.stages;                         // LEADER's copy is used: both CTAs' stages split
  uint64_t* accum_bar = split_bar + p.stages;                         // local: accumulator complete (multicast commit)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = wg_cluster_ctarank();
  const int pair_idx = blockIdx.x >> 1;
  const int tap = pair_idx / p.cin_tiles;
  const int c0 = (pair_idx - tap * p.cin_tiles) * 256 + (int)rank * 128;   // first M row of THIS CTA: cin (cout when swapped)
  const int n_tile0 = blockIdx.y * p.BN;                                    // first N column of the pair's tile
  const int n0 = n_tile0 + (int)rank * (p.BN / 2);                          // first N column this CTA LOADS (its half of B)
  const int kb_begin = blockIdx.z * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.total_kb) kb_end = p.total_kb;
  const int num_kb = kb_end - kb_begin;           // >= 1 by construction, identical in both CTAs

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&split_bar[s], 2 * (WG_SPLIT_THREADS / 32)); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {      // one warp of EACH CTA (same warp id) allocates with cta_group::2
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  wg_cluster_sync_all();           // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const int dy = tap / p.ksz - p.pad, dx = tap % p.ksz - p.pad;
      const uint32_t tx = (uint32_t)half_bytes;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int pix = (kb_begin + i) * WG_KB;
        const int img = pix / (p.H * p.W);
        const int rem = pix - img * p.H * p.W;
        const int y = rem / p.W, x = rem - y * p.W;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], tx);
        if (!p.swap) {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(&map_x, &full_bar[s], st + j * WG_CHUNK_BYTES, c0 + 32 * j, x + dx, y + dy, img);
          for (int j = 0; j < p.nchunks_b; ++j)
            tma_load_4d(&map_dy, &full_bar[s], st + WG_A_BYTES + j * WG_CHUNK_BYTES, n0 + 32 * j, x, y, img);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(&map_dy, &full_bar[s], st + j * WG_CHUNK_BYTES, c0 + 32 * j, x, y, img);
          for (int j = 0; j < p.nchunks_b; ++j)
            tma_load_4d(&map_x, &full_bar[s], st + WG_A_BYTES + j * WG_CHUNK_BYTES, n0 + 32 * j, x + dx, y + dy, img);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      // a_format = b_format = TF32, c = F32, a_major = b_major = MN (bits 15, 16), N = BN, M = 256 (both CTAs)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (uint32_t)(i / p.stages) & 1;
        mbar_wait(&split_bar[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t b_addr = a_addr + WG_A_BYTES;
        const uint32_t lo_off = (uint32_t)half_bytes;
#pragma unroll
        for (int k = 0; k < WG_KB / 8; ++k) {
          const uint64_t ad = make_desc_mn(a_addr + k * 1024, WG_CHUNK_BYTES, 512);
          const uint64_t bd = make_desc_mn(b_addr + k * 1024, WG_CHUNK_BYTES, 512);
          const uint64_t ald = make_desc_mn(a_addr + lo_off + k * 1024, WG_CHUNK_BYTES, 512);
          const uint64_t bld = make_desc_mn(b_addr + lo_off + k * 1024, WG_CHUNK_BYTES, 512);
          wg_mma_tf32_pair(tmem_base, ad, bld, idesc, (i | k) != 0);
          wg_mma_tf32_pair(tmem_base, ald, bd, idesc, 1);
          wg_mma_tf32_pair(tmem_base, ad, bd, idesc, 1);
        }
        wg_commit_pair(&empty_bar[s]);           // frees the stage in both CTAs when these MMAs retire
      }
      wg_commit_pair(accum_bar);                 // accumulator complete -> both epilogues
    }
  } else if (warp < 6) {
    mbar_wait_sleep(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int c = c0 + q * 32 + lane;                       // TMEM lane == M row of this CTA's half of the tile
    if (!p.swap) {
      const bool cvalid = c < p.Cin;
      float* row = p.dw + ((size_t)tap * p.Cin + c) * p.Cout;
      for (int cb = 0; cb < p.BN; cb += 32) {
        if (n_tile0 + cb >= p.Cout) break;                  // warp-uniform: nothing but padding from here on
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
        if (!cvalid) continue;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int n = n_tile0 + cb + e;
          if (n < p.Cout) atomicAdd(row + n, __uint_as_float(v[e]));
        }
      }
    } else {
      const bool cvalid = c < p.Cout;
      float* col = p.dw + (size_t)tap * p.Cin * p.Cout + c;
      for (int cb = 0; cb < p.BN; cb += 32) {
        if (n_tile0 + cb >= p.Cin) break;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
        if (!cvalid) continue;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int n = n_tile0 + cb + e;
          if (n < p.Cin) atomicAdd(col + (size_t)n * p.Cout, __uint_as_float(v[e]));
        }
      }
    }
    tc_fence_before();
  } else {
    const int t = threadIdx.x - 192;
    const int n16 = half_bytes / 16;                          // float4 elements of [A | Bh]
    const uint32_t split_leader = wg_mapa_u32(smem_u32(&split_bar[0]), 0);
    for (int i = 0; i < num_kb; ++i) {
      const int s = i % p.stages;
      const uint32_t ph = (uint32_t)(i / p.stages) & 1;
      mbar_wait(&full_bar[s], ph);
      const float4* hi = reinterpret_cast<const float4*>(smem + (size_t)s * stage_bytes);
      float4* lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + half_bytes);
      for (int idx = t; idx < n16; idx += WG_SPLIT_THREADS) {
        const float4 a = hi[idx];
        float4 l;
        l.x = tf32_rna(a.x - tf32_trunc(a.x)); l.y = tf32_rna(a.y - tf32_trunc(a.y));
        l.z = tf32_rna(a.z - tf32_trunc(a.z)); l.w = tf32_rna(a.w - tf32_trunc(a.w));
        lo[idx] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) wg_mbar_arrive_cluster(split_leader + (uint32_t)s * 8u);
    }
  }

  // the peer's shared memory and barriers must stay alive until the leader's last MMA / multicast commit has retired
  tc_fence_before();
  __syncthreads();
  wg_cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

}  // namespace

bool wgrad_tc_eligible(const WgradProblem& p) {
  if (p.stride != 1 || (p.k != 1 && p.k != 3)) return false;
  if (p.H != p.W || p.Ho != p.H || p.Wo != p.W) return false;
  if (p.W < 8 || p.W > 128 || (p.W & (p.W - 1)) != 0) return false;
  if (p.Cin < 32 || p.Cout < 32) return false;            // 16-channel stem layers: a 128 x 32 MMA tile would be 1.5 % full (0.26 ms for 0.75 GFLOP)
  if ((p.x_cs % 4) != 0 || (p.dy_cs % 4) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p.x) & 15) != 0 || (reinterpret_cast<uintptr_t>(p.dy) & 15) != 0) return false;
  if (p.pad_t != (p.k - 1) / 2 || p.pad_l != p.pad_t) return false;
  if ((long long)p.B * p.H * p.W < 2048) return false;       // too few pixels to amortise a tensor-core tile
  return true;
}

int launch_wgrad_tc(const WgradProblem& p, int split3, cudaStream_t st) {
  static bool attr_set[2] = {false, false};
  WgParams t;
  t.H = p.H; t.W = p.W; t.Cin = p.Cin; t.Cout = p.Cout; t.ksz = p.k; t.pad = p.pad_t;
  // Orientation.  Default: M = cin (128-row tiles), N = cout.  DENSEREG_WGRAD_SWAP=1 exchanges the roles when that does not increase the padded
  // MMA work pad128(M) x pad32(N): Cin = 64 / 80 layers with a wider Cout stop wasting 37-50 % of the MMA rows, and the epilogue's reductions
  // become one 128 B line per instruction (lane = cout) instead of 32 lines (lane = cin row).  Opt-in until measured on the GPU.
  static int swap_mode = -1;
  if (swap_mode < 0) { const char* e = getenv("DENSEREG_WGRAD_SWAP"); swap_mode = e ? atoi(e) : 1; }   // default 1 since round 2: -0.86 ms per micro-batch
  auto pad = [](int v, int m) { return (v + m - 1) / m * m; };
  const long long work_std = (long long)pad(p.Cin, 128) * pad(p.Cout, 32), work_swp = (long long)pad(p.Cout, 128) * pad(p.Cin, 32);
  t.swap = (swap_mode == 1 && work_swp <= work_std) || swap_mode == 2;
  const int Mdim = t.swap ? p.Cout : p.Cin, Ndim = t.swap ? p.Cin : p.Cout;
  int BN = (Ndim + 31) / 32 * 32;
  if (BN > 256) BN = 256;
  if (split3 && BN > 128) BN = 128;
  t.BN = BN; t.nchunks_b = BN / 32;
  const long long M = (long long)p.B * p.H * p.W;
  t.total_kb = (int)((M + WG_KB - 1) / WG_KB);
  t.cin_tiles = (Mdim + 127) / 128;
  const int cout_tiles = (Ndim + BN - 1) / BN;
  const int tiles = t.cin_tiles * p.k * p.k * cout_tiles;
  static int waves = 0;
  if (!waves) { const char* e = getenv("DENSEREG_WGRAD_WAVES"); waves = e && atoi(e) > 0 ? atoi(e) : 1; }   // round-2 sweep: 1 wave 19.94, 2 waves 20.95, 3 waves 21.58 ms per micro-batch
  int splits = (waves * 148) / tiles;                       // floor: keep the CTA count just under `waves` full waves of 148 SMs
  if (splits > t.total_kb / 8) splits = t.total_kb / 8;
  if (splits < 1) splits = 1;
  t.kb_per_split = (t.total_kb + splits - 1) / splits;
  splits = (t.total_kb + t.kb_per_split - 1) / t.kb_per_split;
  t.n_tiles = cout_tiles; t.splits = splits;
  static int trunc_hi = -1;
  if (trunc_hi < 0) { const char* e = getenv("DENSEREG_SPLIT_TRUNC"); trunc_hi = e ? (atoi(e) & 1) : 1; }   // default since round 2: wgrad 5.54 -> 4.99 ms per micro-batch, dW error 4e-6 .. 8e-6 (bar 2e-5)
  t.trunc_hi = trunc_hi;
  int cols = 32; while (cols < BN) cols <<= 1;
  t.tmem_cols = cols;
  const int stage_bytes = (split3 ? 2 : 1) * (WG_A_BYTES + t.nchunks_b * WG_CHUNK_BYTES);
  int stages = (208 * 1024) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) stages = 2;
  t.stages = stages;
  t.bw = p.W < 32 ? p.W : 32;
  t.bh = 32 / t.bw < p.H ? 32 / t.bw : p.H;
  t.bb = 32 / (t.bw * t.bh);
  t.dw = p.dw;
  const size_t smem_bytes = (size_t)stages * stage_bytes + (4 * stages + 2) * 8 + 1024 + 64;

  CUtensorMap mx, mdy;
  cuuint32_t box[4] = {32, (cuuint32_t)t.bw, (cuuint32_t)t.bh, (cuuint32_t)t.bb};
  cuuint64_t xd[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  cuuint64_t xs[3] = {(cuuint64_t)p.x_cs * 4, (cuuint64_t)p.W * p.x_cs * 4, (cuuint64_t)p.H * p.W * p.x_cs * 4};
  if (!encode_map(&mx, p.x, 4, xd, xs, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 0;
  cuuint64_t yd[4] = {(cuuint64_t)p.Cout, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  cuuint64_t ys[3] = {(cuuint64_t)p.dy_cs * 4, (cuuint64_t)p.W * p.dy_cs * 4, (cuuint64_t)p.H * p.W * p.dy_cs * 4};
  if (!encode_map(&mdy, p.dy, 4, yd, ys, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return 0;

  // CTA pairs for the wide layers (kernel comment above): M = max side >= 256 (DENSEREG_WGRAD_PAIR_MINM; 0 = off), 3xTF32 only
  static int pair_min_m = -1;
  if (pair_min_m < 0) { const char* e = getenv("DENSEREG_WGRAD_PAIR_MINM"); pair_min_m = e ? atoi(e) : 256; }
  if (split3 && pair_min_m > 0 && Mdim >= pair_min_m && Mdim > 128) {
    WgParams tp = t;
    int PBN = (Ndim + 63) / 64 * 64;                        // each CTA loads half of the tile in whole 32-channel chunks
    if (PBN > 256) PBN = 256;
    tp.BN = PBN; tp.nchunks_b = PBN / 64;
    tp.cin_tiles = (Mdim + 255) / 256;                      // 256-row M pairs
    const int n_tiles = (Ndim + PBN - 1) / PBN;
    const int ptiles = tp.cin_tiles * p.k * p.k * n_tiles;
    int psplits = (waves * 74) / ptiles;                    // one wave of CTA pairs
    if (psplits > tp.total_kb / 8) psplits = tp.total_kb / 8;
    if (psplits < 1) psplits = 1;
    tp.kb_per_split = (tp.total_kb + psplits - 1) / psplits;
    psplits = (tp.total_kb + tp.kb_per_split - 1) / tp.kb_per_split;
    tp.n_tiles = n_tiles; tp.splits = psplits; tp.trunc_hi = 1;
    int pcols = 32; while (pcols < PBN) pcols <<= 1;
    tp.tmem_cols = pcols;
    const int pstage = 2 * (WG_A_BYTES + tp.nchunks_b * WG_CHUNK_BYTES);
    int pstages = (208 * 1024) / pstage;
    if (pstages > 6) pstages = 6;
    tp.stages = pstages;
    const size_t psmem = (size_t)pstages * pstage + (3 * pstages + 1) * 8 + 16 + 1024 + 64;
    static bool pattr = false;
    if (!pattr) {
      if (!launch_ok(cudaFuncSetAttribute(wgrad_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024), "wgrad_tc_pair_kernel smem attribute")) return 0;
      pattr = true;
    }
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * tp.cin_tiles * p.k * p.k, n_tiles, psplits); cfg.blockDim = dim3(192 + WG_SPLIT_THREADS);
    cfg.dynamicSmemBytes = psmem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = dr_pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    return launch_ok(cudaLaunchKernelEx(&cfg, wgrad_tc_pair_kernel, mx, mdy, tp), "wgrad_tc_pair_kernel") ? 1 : 0;
  }

  static int atmem = -1;
  if (atmem < 0) { const char* e = getenv("DENSEREG_WGRAD_A_TMEM"); atmem = (e && e[0] == '1') ? 1 : 0; }
  if (atmem && split3 && BN <= 128) {
    // A split into tensor memory: smem stage = A fp32 + B hi + B lo; tensor memory = 2*BN accumulator columns + 64 per ring stage
    static bool aattr = false;
    static int num_sms_a = 0;
    if (!num_sms_a) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms_a, cudaDevAttrMultiProcessorCount, dev); if (num_sms_a <= 0) num_sms_a = 148; }
    const int a_stage = WG_A_BYTES + 2 * t.nchunks_b * WG_CHUNK_BYTES;
    const int a_col0 = (2 * BN + 31) / 32 * 32;
    int a_stages = (208 * 1024) / a_stage;
    if (a_stages > (512 - a_col0) / 64) a_stages = (512 - a_col0) / 64;
    if (a_stages > 6) a_stages = 6;
    if (a_stages >= 2) {
      WgParams ta = t;
      ta.stages = a_stages;
      int acols = 32; while (acols < a_col0 + 64 * a_stages) acols <<= 1;
      ta.tmem_cols = acols;
      const int total_items = tiles * splits;
      const size_t asmem = (size_t)a_stages * a_stage + (3 * a_stages + 4) * 8 + 16 + 1024 + 64;
      if (!aattr) { cudaFuncSetAttribute(wgrad_tc_atmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024); aattr = true; }
      dr_launch(wgrad_tc_atmem_kernel, dim3(dim3(total_items < num_sms_a ? total_items : num_sms_a)), dim3(192 + WG_SPLIT_THREADS), asmem, st, mx, mdy, ta);
      return launch_ok(cudaPeekAtLastError(), "wgrad_tc_atmem_kernel") ? 1 : 0;
    }
  }
  dim3 grid(t.cin_tiles * p.k * p.k, cout_tiles, splits);
  if (split3) {
    if (!attr_set[1]) { cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024); attr_set[1] = true; }
    dr_launch(wgrad_tc_kernel<true>, dim3(grid), dim3(192 + WG_SPLIT_THREADS), smem_bytes, st, mx, mdy, t);
  } else {
    if (!attr_set[0]) { cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024); attr_set[0] = true; }
    dr_launch(wgrad_tc_kernel<false>, dim3(grid), dim3(192), smem_bytes, st, mx, mdy, t);
  }
  return launch_ok(cudaPeekAtLastError(), "wgrad_tc_kernel") ? 1 : 0;
}
