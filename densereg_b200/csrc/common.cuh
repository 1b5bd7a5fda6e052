// common.cuh -- shared declarations for libdensereg_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <utility>

#define DR_DEVINL __device__ __forceinline__

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------------------
// Consecutive kernels of one stream are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may start
// (launch latency, barrier init, tensor-memory allocation, tensor-map prefetch) while the previous kernel drains, and block in
// griddepcontrol.wait until it has completed and its memory is visible.  EVERY kernel launched through dr_launch() executes pdl_wait() in all
// of its threads before it touches global memory; pdl_trigger() lets its own dependents start launching as soon as all of its CTAs are
// resident.  The step is ~950 launches per micro-batch, most of them 5-20 us on small grids (profiles/r2_final.md).
#ifdef __CUDACC__
DR_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
DR_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool dr_pdl_enabled();                 // engine.cu: DENSEREG_PDL != 0 and not suspended (CUDA-graph capture)
void dr_pdl_suspend(int on);

template <class... KArgs, class... Args>
inline cudaError_t dr_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = dr_pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// counter-based keep/drop decision for dropout (keep prob 0.5, network/slim/ops.py:711).
// Same splitmix64 finaliser as oracle/um_v1_torch.py:dropout_mask so masks are reproducible.
__host__ __device__ inline uint32_t dr_hash_keep(uint64_t seed, uint32_t tag, uint64_t i) {
  uint64_t x = i + seed * 0x9E3779B97F4A7C15ull + (uint64_t)tag * 0xD1B54A32D192ED03ull;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 63);
}

// ---- conv problem description shared by the SIMT and tcgen05 paths --------------------------
struct ConvProblem {
  // input activation view: element (b,y,x,c) at x[((b*H + y)*W + x)*x_cs + c]
  const float* x; int x_cs;
  int B, H, W, Cin;
  int Ho, Wo, Cout;
  int k, stride, pad_t, pad_l;
  const float* w;          // [k*k*Cin][Cout] row-major (TF HWIO flattened)
  int w_ld;                // row length of w in floats (0 = Cout); the dgrad copy has rows padded to 4 floats
  int flip_taps;           // read weight taps in reverse order (dgrad == conv with the 180-degree rotated filter)
  const float* w_kmajor;   // tensor-core path: 16 B aligned K-major copy [tap][Cout][Cin] (hi part for 3xTF32), or null
  const float* w_kmajor_lo;// 3xTF32: lo part
  int wk_ld;               // row length (floats, multiple of 4) of the K-major copy: Cin rounded up to 4
  int pair;                // 3xTF32 tensor-core path: allow the CTA-pair (cta_group::2) kernel for big layers (conv_tc_pair.cu)
  int chunk_kb;            // 3xTF32 tensor-core path: two-level accumulation, k-blocks (of 32 channels) per partial accumulator; 0 = one level
  int chunk_min_kb;        // ... only for reductions longer than this many k-blocks
  float* y; int y_cs;      // output view
  // epilogue: v = acc*scale[n] + shift[n]; relu; dropout; + res; + beta*y_old
  const float* scale;      // may be null (=1)
  const float* shift;      // may be null (=0)
  int relu;
  const float* res; int res_cs;   // residual view at output resolution (may be null)
  int accumulate;          // y = v + y_old
  int dropout;             // apply keep/2x after relu
  uint64_t drop_seed; uint32_t drop_tag;
  // tensor-core path only: fused BRN batch statistics of the RAW conv output + finalize by the last CTA (all null = off)
  double* stats;           // [2*Cout] sum / sum of squares (pre-zeroed)
  unsigned int* stats_counter;   // pre-zeroed
  const float* bn_bg; float* bn_state; float* bn_aff; float* bn_bstat; int bn_update_state;
};

struct WgradProblem {
  const float* x; int x_cs;       // forward input view (B,H,W,Cin)
  const float* dy; int dy_cs;     // grad of raw conv output (B,Ho,Wo,Cout)
  int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad_t, pad_l;
  float* dw;                      // [k*k*Cin][Cout], ACCUMULATED with atomics
};

// launchers (each returns the number of kernels launched)
int launch_conv_simt(const ConvProblem& p, cudaStream_t st);
int launch_wgrad_simt(const WgradProblem& p, cudaStream_t st);

// tcgen05 path (conv_tc.cu). Returns 0 launches if the problem shape is not eligible.
const char* tc_last_error();     // why the last tensor-core launch of this thread returned 0 launches
bool conv_tc_eligible(const ConvProblem& p);
int launch_conv_tc(const ConvProblem& p, int split3, cudaStream_t st);
// CTA-pair (cta_group::2) 3xTF32 variant (conv_tc_pair.cu), opt-in (ConvProblem::pair); same contract as launch_conv_tc
// A-operand-in-tensor-memory 3xTF32 variant (conv_tc_atmem.cu), opt-in (DENSEREG_TC_A_TMEM); same contract
int conv_tc_atmem_mode();
int launch_conv_tc_atmem(const ConvProblem& p, cudaStream_t st);
bool conv_tc_pair_wanted(const ConvProblem& p);
int launch_conv_tc_pair(const ConvProblem& p, cudaStream_t st);
bool wgrad_tc_eligible(const WgradProblem& p);
int launch_wgrad_tc(const WgradProblem& p, int split3, cudaStream_t st);

// ---- vote -----------------------------------------------------------------------------------
int launch_vote(int B, int H, int W, int J,
                const float* hm, int hm_cs, const float* hm3, int hm3_cs, const float* um, int um_cs,
                const float* dm, const float* cfgs, const float* coms,
                float* xyz, int32_t* top5, int32_t* clamp_count, cudaStream_t st);

// ---- elementwise / reductions (ew.cu) ---------------------------------------------------------
int launch_norm_dm(int B, int hw, const float* dm, const float* coms, float* out, cudaStream_t st);
// writes uu,vv,tiny_dm (3 channels) for every 32x32 pixel into up to 8 destination views
struct UvdDst { float* p[8]; int cs[8]; int n; };
int launch_make_uvd(int B, int in_hw, int out_hw, const float* x0, float* tiny, UvdDst dst, cudaStream_t st);
int launch_maxpool(int B, int H, int W, int C, int k, const float* x, int x_cs, float* y, int y_cs, cudaStream_t st);
int launch_maxpool_bwd(int B, int H, int W, int C, int k, const float* x, int x_cs, const float* dy, int dy_cs,
                       float* dx, int dx_cs, int accumulate, cudaStream_t st);
int launch_upadd(int B, int H, int W, int C, const float* a, int a_cs, const float* lo, int lo_cs, float* y, int y_cs, cudaStream_t st);
// dlo (H/2,W/2) (+)= 2x2 sum of dy
int launch_upadd_bwd_lo(int B, int H, int W, int C, const float* dy, int dy_cs, float* dlo, int dlo_cs, int accumulate, cudaStream_t st);
// dst (+)= src (views, same shape); optional depth mask (zero where tiny<-0.9)
int launch_copy_view(size_t npix, int C, const float* src, int src_cs, float* dst, int dst_cs, int accumulate,
                     const float* tiny_mask, cudaStream_t st);
int launch_fill_view(size_t npix, int C, float* dst, int dst_cs, float v, cudaStream_t st);
// stats + finalize in ONE launch (last block finalizes); counter must be zero on entry
int launch_channel_stats_finalize(size_t npix, int C, const float* x, int x_cs, double* sums, unsigned int* counter,
                                  const float* beta_gamma, float* state, float* aff, float* bstat, int update_state, cudaStream_t st);
// y = act(raw*a+b) (+res)
int launch_brn_apply(size_t npix, int C, const float* raw, int raw_cs, const float* aff, int relu,
                     const float* res, int res_cs, float* y, int y_cs, cudaStream_t st);
// backward reductions: g = dy*(z>0); sums[0:C]=sum g, sums[C:2C]=sum g*xhat   (z = raw*a+b, xhat=(raw-mean)*inv_std)
int launch_brn_bwd_small(size_t npix, int C, const float* dy, int dy_cs, const float* raw, int raw_cs, const float* aff, const float* bstat,
                         const float* beta_gamma, int relu, float* draw, int draw_cs, float* gparam, cudaStream_t st);   // 1 = done in one launch, 0 = not applicable
int launch_brn_bwd_reduce(size_t npix, int C, const float* dy, int dy_cs, const float* raw, int raw_cs,
                          const float* aff, const float* bstat, int relu, double* sums, cudaStream_t st);
// draw = gamma*r*inv_std*(g - sum_g/N - xhat*sum_gx/N); also dbeta,dgamma accumulated into gparam
int launch_brn_bwd_apply(size_t npix, int C, const float* dy, int dy_cs, const float* raw, int raw_cs,
                         const float* aff, const float* bstat, const float* beta_gamma, int relu,
                         const double* sums, float* draw, int draw_cs, float* gparam, cudaStream_t st);
// bias conv backward: dz = dy * (relu? (out>0)*(dropout?2:1) : 1); dbias accumulated
int launch_bias_bwd(size_t npix, int C, const float* dy, int dy_cs, const float* out, int out_cs, int relu, int dropout,
                    float* dz, int dz_cs, float* gbias, cudaStream_t st);
// loss + GT synthesis + output grads for all stacks
struct LossArgs {
  int B, hw, J, S;
  const float* tiny;          // (B,hw,hw) normalised depth
  const float* poses; const float* cfgs; const float* coms;
  const float* hm[4]; const float* hm3[4]; const float* um[4]; int cs[4];
  float* ghm[4]; float* ghm3[4]; float* gum[4]; int gcs[4];
  double* loss_acc;           // 3 doubles (hm, hm3, um), pre-zeroed
};
int launch_loss(const LossArgs& a, cudaStream_t st);
int launch_wd(size_t n, const float* params, const float* wdmask, float* grads, double* reg_acc, cudaStream_t st);
int launch_finish_loss(const double* acc /*hm,hm3,um,reg*/, float* out5, cudaStream_t st);
// g / divisor, clip, TF ApplyAdam: alpha = lr*sqrt(1-b2^t)/(1-b1^t) (fp32), omb = 1.0f - beta
int launch_adam(size_t n, float* p, const float* g, float* m, float* v, float divisor, float clip,
                float alpha, float omb1, float omb2, float eps, cudaStream_t st);
int launch_init_trunc_normal(size_t n, float* p, float stddev, uint64_t seed, cudaStream_t st);
int launch_gather_outputs(size_t npix, int C, const float* src, int src_cs, float* dst, cudaStream_t st);

// ---- crop + centre-of-mass front-end (crop.cu) ---------------------------------------------------
size_t crop_scratch_bytes(int B);
int launch_crop(int B, int in_h, int in_w, const float* frames, const float* poses, int J, const float* bbx, const float cfg_host[6],
                int out_hw, float pad, int icvl, void* scratch, float* dm_out, float* cfg_out, float* com_out, cudaStream_t st);
int launch_data_aug(int B, int hw, int J, const float* dms, const float* poses, const float* cfgs, const float* coms, const float* cossin,
                    const float* edge_ratio, float* dms_out, float* poses_out, cudaStream_t st);
