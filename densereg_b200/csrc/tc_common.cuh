// tc_common.cuh -- inline-PTX wrappers shared by the tcgen05 kernels (conv_tc.cu, wgrad_tc.cu): mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05.mma / commit / ld / fences, UMMA shared-memory descriptors, tensor-map encoding.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <unordered_map>

namespace tc {

DR_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DR_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DR_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DR_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DR_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// same, with back-off: for warps that wait for the whole main loop (epilogue) so that their polling does not steal issue
// slots from the single MMA-issuing thread
DR_DEVINL void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok;
  while (true) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(64);
  }
}
DR_DEVINL void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
DR_DEVINL void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
DR_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DR_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
DR_DEVINL void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DR_DEVINL void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// round-to-nearest fp32 -> tf32 (exactly representable, so the tensor core's own conversion is the identity and the
// split error is unbiased; plain truncation gives a bias that grows linearly with K)
DR_DEVINL float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// what the tensor core itself reads from a 32-bit kind::tf32 operand: sign, exponent and the top 10 mantissa bits
DR_DEVINL float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO=1024B | version 1 | layout 2
DR_DEVINL uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
DR_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


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
DR_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MN-major TF32 operand descriptor.  For 32-bit MN-major operands the only layout the tensor core accepts is
// SWIZZLE_128B_BASE32B (cute Layout_MN_SW128_32B_Atom: 32-float = 128 B runs along M/N, 4 k-rows per 512 B atom, 32 B chunks
// XOR-swizzled with the row index; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  LBO = byte stride between 32-wide M/N
// chunks, SBO = byte stride between 4-row k groups.
DR_DEVINL uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
         (1ull << 61);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled is a driver entry point; resolve it through the runtime so that the library does not link
// libcuda.so (and therefore still loads on a CPU-only machine for the symbol / layer-table checks).
inline EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
    else
      fprintf(stderr, "densereg: cuTensorMapEncodeTiled not available; tensor-core path disabled\n");
  }
  return fn;
}

// why the last tensor-core launch of this thread failed (run_conv / run_wgrad report it through dr_last_error; there is no FFMA fallback)
inline char* last_error_buf() { static thread_local char buf[192] = "ok"; return buf; }
inline const char* last_error() { return last_error_buf(); }
inline void set_error(const char* what, int code) { snprintf(last_error_buf(), 192, "%s (code %d)", what, code); }

// Tensor maps are pure functions of (base, rank, dims, strides, box, swizzle).  The arena pointers, layer shapes and batch sizes of a
// handle repeat every micro-batch, so each map is encoded ONCE and looked up afterwards: 2-5 driver calls per tensor-core launch
// (~600 launches per micro-batch) become hash lookups.  DENSEREG_TMAP_CACHE=0 disables the cache.
struct MapKey {
  const void* base; int rank; int swizzle; cuuint64_t dims[5]; cuuint64_t strides[4]; cuuint32_t box[5];
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t hsh = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { hsh ^= w[i] + 0x9E3779B97F4A7C15ull + (hsh << 6) + (hsh >> 2); }
    return (size_t)hsh;
  }
};
struct MapCache {
  std::mutex mu;
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
  int enabled = -1;
};
inline MapCache& map_cache() { static MapCache c; return c; }

inline bool encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                const cuuint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable", -1); return false; }
  MapCache& mc = map_cache();
  MapKey key; memset(&key, 0, sizeof(key));
  key.base = base; key.rank = rank; key.swizzle = (int)swizzle;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  std::lock_guard<std::mutex> lock(mc.mu);
  if (mc.enabled < 0) { const char* e = getenv("DENSEREG_TMAP_CACHE"); mc.enabled = (e && e[0] == '0') ? 0 : 1; }
  if (mc.enabled) {
    auto it = mc.maps.find(key);
    if (it != mc.maps.end()) { memcpy(m, &it->second, sizeof(CUtensorMap)); return true; }
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed", (int)r);
    return false;
  }
  if (mc.enabled) {
    if (mc.maps.size() > 65536) mc.maps.clear();            // debug entry points with ever-changing pointers must not grow it without bound
    mc.maps.emplace(key, *m);
  }
  return true;
}

// launch status of a tensor-core kernel: records the reason when it is not cudaSuccess
inline bool launch_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  snprintf(last_error_buf(), 192, "%s: %s", what, cudaGetErrorString(e));
  cudaGetLastError();
  return false;
}


}  // namespace tc
