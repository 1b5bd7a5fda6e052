/* densereg.h -- C-ABI of libdensereg_sm100.so, the B200-native replacement for the
 * denseReg hot path (depth crop -> um_v1 stacked hourglass -> {hm,hm3,um} -> offset vote
 * -> joint xyz in mm; plus the data-parallel training step).
 *
 * The reference (melonwan/denseReg) has no FFI: its seams are Python call signatures.  Each
 * entry point below names the reference interface (file:line under /root/reference) that a
 * binding would replace.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - every tensor pointer is a DEVICE pointer owned by the caller (PyTorch tensors'
 *     data_ptr()); layout NHWC fp32 exactly like the reference's TF tensors; the library owns
 *     only its handle, an activation workspace and scratch.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     enqueued on it and the call returns without synchronising unless stated.
 *   - return 0 on success, a negative dr_status otherwise; dr_last_error() gives the message.
 *     Nothing throws across the boundary.  One handle per device per process; not re-entrant.
 */
#ifndef DENSEREG_H_
#define DENSEREG_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DR_VERSION 100
#define DR_API __attribute__((visibility("default")))

typedef enum {
  DR_OK = 0,
  DR_ERR_ARG = -1,      /* bad argument / shape                 */
  DR_ERR_CUDA = -2,     /* CUDA runtime error                   */
  DR_ERR_STATE = -3,    /* call order (e.g. buffers not bound)  */
  DR_ERR_NOMEM = -4,
  DR_ERR_UNSUPPORTED = -5
} dr_status;

/* conv arithmetic */
typedef enum {
  DR_PREC_FP32 = 0,     /* SIMT FFMA, fp32 accumulate (parity path)                        */
  DR_PREC_TF32 = 1,     /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM                 */
  DR_PREC_TF32X3 = 2    /* tcgen05 3-pass split TF32 (fp32-class accuracy)                 */
} dr_precision;

/* Mirrors the reference flag surface that shapes the graph:
 * model/hourglass_um_crop_tiny.py:29-62 (--num_stack --num_fea --kernel_size), dataset jnt_num
 * (data/icvl.py:17, nyu.py:40-45, msra.py:17), input/output size (:82-87). */
typedef struct {
  int32_t num_stack;     /* FLAGS.num_stack, default 2   */
  int32_t num_fea;       /* FLAGS.num_fea, default 128   */
  int32_t kernel_size;   /* FLAGS.kernel_size, must be 3 */
  int32_t num_jnt;       /* 16 icvl / 14 nyu / 21 msra   */
  int32_t in_hw;         /* 128                          */
  int32_t out_hw;        /* 32                           */
  int32_t max_batch;     /* workspace is sized for this  */
  int32_t precision;     /* dr_precision                 */
  int32_t device;        /* CUDA ordinal                 */
  int32_t reserved[7];   /* reserved[0] != 0: dr_infer replays a CUDA graph captured per (batch, pointer) key;
                            reserved[1] < 0: do NOT run the 3xTF32 convs of the big layers on CTA pairs (tcgen05 cta_group::2; default on);
                            reserved[2] == 2: micro-batch pipeline for training (see dr_pipeline_join) */
} dr_config;

typedef struct dr_handle dr_handle;

/* one row of the conv table, in TF variable-creation order of network/um_v1.py:71-185 */
typedef struct {
  char name[48];
  int32_t k, stride, cin, cout, brn, relu;
  float wd;              /* l2 weight decay (network/slim/losses.py:56-72), 0 for inter-stack convs */
  int64_t w_off;         /* offset of HWIO weights in the flat parameter buffer                    */
  int64_t p_off;         /* beta[cout],gamma[cout] (brn) or biases[cout]                           */
  int64_t s_off;         /* BRN state: mov_mean, mov_var, biased_mean, biased_var [cout each],
                            r_max, d_max, curr_t, local_step                                       */
  int32_t in_hw, out_hw;
} dr_layer_info;

DR_API int dr_version(void);

/* Graph construction == network/um_v1.py:detect_net (:71-185) + JointDetectionModel.__init__
 * (model/hourglass_um_crop_tiny.py:92-127). */
DR_API int dr_create(dr_handle** out, const dr_config* cfg);
DR_API int dr_destroy(dr_handle* h);
DR_API const char* dr_last_error(const dr_handle* h);

DR_API size_t dr_param_count(const dr_handle* h);   /* trainable fp32 scalars (tf.trainable_variables)       */
DR_API size_t dr_state_count(const dr_handle* h);   /* non-trainable BRN state (network/slim/ops.py:100-128) */
DR_API int dr_num_layers(const dr_handle* h);
DR_API int dr_get_layer(const dr_handle* h, int idx, dr_layer_info* out);

/* Caller-owned flat fp32 device buffers (replaces tf.Variable storage / tf.train.Saver contents,
 * model/train_single_gpu.py:108).  grads/adam_m/adam_v may be NULL for inference-only use. */
DR_API int dr_bind(dr_handle* h, float* params, float* state, float* grads, float* adam_m, float* adam_v);

/* The caller wrote into the bound `params` buffer itself (checkpoint restore = saver.restore, model/test_model.py:31-35): the
 * library rebuilds its tensor-core weight copies on the next forward.  dr_init_params / dr_optimizer_step / dr_bind do this
 * implicitly. */
DR_API int dr_params_changed(dr_handle* h);

/* ops.py:272 truncated_normal(stddev), ops.py:86-128 BRN initial values.  Needs dr_bind first. */
DR_API int dr_init_params(dr_handle* h, uint64_t seed, float stddev, void* stream);

/* data/preprocess.py:176-187 norm_dm.  dm_mm (B,HW,HW), coms (B,3) -> out (B,HW,HW). */
DR_API int dr_norm_dm(dr_handle* h, int B, int hw, const float* dm_mm, const float* coms, float* out, void* stream);

/* JointDetectionModel.inference -> detect_net (hourglass_um_crop_tiny.py:186-191, um_v1.py:71-185),
 * including norm_dm.  dm_mm (B,128,128,1) raw depth in mm, coms (B,3).
 * hm/hm3/um: arrays of num_stack device pointers (B,32,32,J)/(B,32,32,J)/(B,32,32,3J); an entry or
 * the array itself may be NULL to skip the copy-out.  is_training selects BRN batch statistics +
 * dropout (ops.py:130-171, :726); update_state applies the BRN UPDATE_OPS (ops.py:134-153). */
DR_API int dr_forward(dr_handle* h, int B, const float* dm_mm, const float* coms,
               float* const* hm, float* const* hm3, float* const* um,
               int is_training, int update_state, uint64_t dropout_seed, void* stream);

/* JointDetectionModel._resume_om + _xyz_estimation + unnorm_xyz_pose
 * (hourglass_um_crop_tiny.py:276-299, :743-785; data/preprocess.py:157-170,189-232).
 * hm,hm3 (B,H,W,J), um (B,H,W,3J) dense; dm_norm (B,H,W) normalised depth at H x W; cfgs (B,6);
 * coms (B,3).  xyz_mm (B,3J).  top5_idx (B,J,5) int32 and clamp_count (1 int32, re-projections
 * that fell outside the map and were clamped) are optional. */
DR_API int dr_vote(dr_handle* h, int B, int H, int W, int J,
            const float* hm, const float* hm3, const float* um, const float* dm_norm,
            const float* cfgs, const float* coms,
            float* xyz_mm, int32_t* top5_idx, int32_t* clamp_count, void* stream);

/* JointDetectionModel.test (hourglass_um_crop_tiny.py:442-462): norm_dm -> inference(eval) ->
 * vote -> xyz in mm.  dm_mm (B,128,128,1), cfgs (B,6), coms (B,3) -> xyz_mm (B,3J). */
DR_API int dr_infer(dr_handle* h, int B, const float* dm_mm, const float* cfgs, const float* coms,
             float* xyz_mm, int32_t* top5_idx, void* stream);

/* JointDetectionModel.loss (hourglass_um_crop_tiny.py:323-371, no data_aug) + TF autodiff +
 * accum_op (model/train_single_gpu.py:69-84): one micro-batch forward+backward; gradients are
 * ADDED into the bound grads buffer.  poses_mm (B,3J).  loss_out: 5 floats on the DEVICE
 * {total, hm, hm3, um, reg}. */
DR_API int dr_loss_backward(dr_handle* h, int B, const float* dm_mm, const float* poses_mm,
                     const float* cfgs, const float* coms, float* loss_out,
                     uint64_t dropout_seed, int update_state, void* stream);

/* Micro-batch pipeline (dr_config.reserved[2] == 2; doubles the activation workspace).  The sub_batch micro-batches of one optimiser
 * step (model/train_single_gpu.py:140-148: `for sub in range(sub_batch): sess.run([loss, accum_op])`) only meet in the BRN moving
 * statistics (written by the forward pass, network/slim/ops.py:141-162) and in the gradient accumulators (written by the backward pass,
 * train_single_gpu.py:69-84).  With the pipeline on, consecutive dr_loss_backward calls alternate between two activation arenas on two
 * internal streams: the FORWARD pass of micro-batch i+1 runs next to the BACKWARD pass of micro-batch i; forward passes stay in call order
 * (same BRN state sequence as the reference), backward passes never overlap each other.  On return the caller's stream is ordered after
 * the forward pass and the loss of THIS micro-batch (loss_out valid, inputs reusable) but not after its backward pass: dr_zero_grads,
 * dr_optimizer_step, dr_forward, dr_infer and the debug entry points join the pipeline themselves; a caller that reads the bound grads
 * buffer directly calls dr_pipeline_join(h, stream) first.  dr_pipeline_depth: 1 (off) or 2. */
DR_API int dr_pipeline_join(dr_handle* h, void* stream);
DR_API int dr_pipeline_depth(const dr_handle* h);
/* Debug / CPU test (tests/test_pipeline_plan.py): the pipeline's stream operations as data -- the very lists dr_loss_backward / dr_pipeline_join
 * execute.  Dry run on a handle that was never bound (no CUDA call): what = 0 the next dr_loss_backward (advances the slot bookkeeping like the real
 * call), 1 dr_pipeline_join / dr_zero_grads, 2 dr_optimizer_step (join + parameters changed), 3 dr_comm_overlap_next_backward (arms the next pass).
 * kind: 0 record `event` on `stream`, 1 `stream` waits for `event`, 2 forward pass on `stream`, 3 backward pass on `stream` (records `event` right after
 * its loss kernels); stream: -1 the caller's, 0 / 1 the slot's internal stream; event: 0 in, 1 / 2 forward done [slot], 3 / 4 backward done [slot],
 * 5 / 6 loss done [slot].  Returns the number of ops written (cap >= 12) or a negative status. */
typedef struct dr_pipe_op { int32_t kind, stream, event; } dr_pipe_op;
DR_API int dr_debug_pipeline_plan(dr_handle* h, int what, dr_pipe_op* out, int cap);

/* Data-parallel communicator: replaces model/train_multi_gpu.py:16-39 (_average_gradients: per-variable concat + mean through host
 * memory) and :63-64,73-92 (towers) with one process per GPU and ONE NCCL all-reduce(sum) of the flat gradient per optimiser step.
 *   dr_comm_unique_id   rank 0 fills 128 bytes (ncclUniqueId); the caller ships them to the other ranks (any transport).
 *   dr_comm_init        every rank, after dr_create: creates the communicator on the handle's device.  world == 1 is a no-op.
 *   dr_comm_overlap_next_backward   call before the LAST dr_loss_backward of an optimiser step: that backward pass all-reduces the
 *                       gradient in buckets (last layers first) on a communication stream as soon as each bucket is final, so the
 *                       transfer overlaps the rest of the backward pass; dr_optimizer_step then only waits for it.  Without this call
 *                       dr_optimizer_step all-reduces the whole buffer itself before the update.
 * NCCL is resolved with dlopen("libnccl.so.2") at dr_comm_init time (inside a PyTorch process: the copy PyTorch already loaded). */
DR_API int dr_comm_unique_id(void* out128);
DR_API int dr_comm_init(dr_handle* h, int rank, int world, const void* nccl_unique_id128);
DR_API int dr_comm_overlap_next_backward(dr_handle* h);
DR_API int64_t dr_comm_allreduce_count(const dr_handle* h);

/* reset_op (train_single_gpu.py:83) */
DR_API int dr_zero_grads(dr_handle* h, void* stream);

/* ave_grad + clip + Adam apply (train_single_gpu.py:86-88, hourglass_um_crop_tiny.py:436-439):
 * g = clip(grads / (accum_steps*world), +-0.2); Adam(beta1 .5, beta2 .999, eps 1e-8) with TF's
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t).  With a communicator (dr_comm_init, world > 1) the sum over ranks happens here (or was
 * overlapped with the last backward pass, dr_comm_overlap_next_backward); without one, `grads` must already hold the sum over
 * micro-batches and ranks.  step is the 1-based optimiser step. */
DR_API int dr_optimizer_step(dr_handle* h, int accum_steps, int world, float lr, int64_t step, void* stream);

/* Depth-frame front-end ("next" row 8f-1): data/preprocess.py:10-79 crop_from_xyz_pose + :131-142 center_of_mass.
 * frames (B,in_h,in_w) raw depth mm, poses (B,3J) xyz mm (GT or estimated), cfg: 6 floats on the HOST [fx,fy,cx,cy,w,h] of the
 * full frame (data/icvl.py:12, nyu.py:13, msra.py:13), pad = 20, icvl != 0 selects the fixed 500 mm threshold (:62-63).
 * Outputs: dm_out (B,out_hw,out_hw,1), cfg_out (B,6) crop intrinsics, com_out (B,3) mm -- exactly the (dms, cfgs, coms) batch
 * that dr_infer / dr_loss_backward consume. */
DR_API int dr_crop_from_xyz_pose(dr_handle* h, int B, int in_h, int in_w, const float* frames, const float* poses, int J,
                                 const float* cfg_host6, int out_hw, float pad, int icvl,
                                 float* dm_out, float* cfg_out, float* com_out, void* stream);
/* data/preprocess.py:81-129 crop_from_bbx: bbx (B,5) device [top,left,bottom,right,d_th] (data/nyu_bbx.pkl rows). */
DR_API int dr_crop_from_bbx(dr_handle* h, int B, int in_h, int in_w, const float* frames, const float* bbx, const float* cfg_host6,
                            int out_hw, float* dm_out, float* cfg_out, float* com_out, void* stream);

/* Training-time augmentation ("next" row 8f-3): data/preprocess.py:234-267 data_aug on a batch of crops.
 * dms (B,hw,hw,1), poses (B,3J), cfgs (B,6), coms (B,3); cossin (B,2) = cos/sin of the rotation angle ~ U(-pi,pi) and
 * edge_ratio (B,2) = clip(N(1,0.2),0.9,1.1) [height, width] are drawn by the CALLER (TF's RNG stream is not reproducible).
 * Outputs dms_out (B,hw,hw,1), poses_out (B,3J). */
DR_API int dr_data_aug(dr_handle* h, int B, int hw, int J, const float* dms, const float* poses, const float* cfgs, const float* coms,
                       const float* cossin, const float* edge_ratio, float* dms_out, float* poses_out, void* stream);

/* per-conv debug entry used by the parity tests: runs ONE conv of the table on caller data.
 * x (B,H,W,cin) dense -> y (B,Ho,Wo,cout) = conv(x, W[idx]) (no BRN/bias/activation). */
DR_API int dr_debug_conv(dr_handle* h, int layer, int B, const float* x, float* y, int precision, void* stream);
/* dgrad / wgrad of the same conv: dy (B,Ho,Wo,cout) -> dx (B,H,W,cin) (overwritten),
 * dw (k*k*cin*cout) (overwritten). Either output may be NULL. */
DR_API int dr_debug_conv_bwd(dr_handle* h, int layer, int B, const float* x, const float* dy,
                      float* dx, float* dw, int precision, void* stream);

/* debug: copy the activation (grad=0) or its gradient (grad=1) that conv `layer` wrote in the last
 * forward/backward into dst (B,Ho,Wo,cout) dense.  For the last conv of a residual block this is the block
 * output (post-activation conv + skip), as in network/um_v1.py:48. */
DR_API int dr_debug_get_output(dr_handle* h, int layer, int B, float* dst, int grad, void* stream);

/* Per-launch timing of the conv-type kernels (measurement only; no reference counterpart).  dr_trace(h, 1) clears the record list and
 * makes every conv / dgrad / wgrad launch time itself with CUDA events on its own stream (this serialises the launches: use it on a
 * separate pass, never inside a timed region); dr_trace(h, 0) stops.  bench.py derives the per-class roofline from these records. */
typedef struct {
  int32_t kind;          /* 0 forward conv, 1 dgrad, 2 wgrad            */
  int32_t B, hw, cin, cout, k;
  int32_t kernel;        /* 0 FFMA, 1 tcgen05 one-CTA, 2 tcgen05 CTA pair */
  float ms;
} dr_trace_rec;
DR_API int dr_trace(dr_handle* h, int on);
DR_API int dr_trace_count(const dr_handle* h);
DR_API int dr_trace_get(const dr_handle* h, int idx, dr_trace_rec* out);

/* One op of the execution graph with its lane plan (debug / tests; works without a GPU).  The library runs independent branches of
 * um_v1 -- the `upper1` block of every hourglass level (network/um_v1.py:54-65), the masked um branch (:143-149), projection skips
 * (:31-47) -- on separate CUDA streams ("lanes"); a build-time hazard analysis over the buffer views each op reads / writes decides
 * which events an op waits for.  tests/test_oracle_net.py re-derives every hazard by brute force from this table. */
typedef struct {
  int32_t kind;             /* 0 conv, 1 max-pool, 2 upsample+add, 3 masked copy */
  int32_t lane, layer, need_dgrad, raw_buf;
  int32_t in_buf, in_c0, in_c;      /* views: buffer id, first channel, channels */
  int32_t out_buf, out_c0, out_c;
  int32_t res_buf, res_c0, res_c;   /* residual operand (conv) / low-resolution operand (upsample+add); buf < 0: none */
  int32_t nwait[2], wait_op[2][3], record[2];   /* [0] forward pass, [1] backward pass: ops (of other lanes) whose event this op waits for */
  /* backward-pass gradient aliasing (the gradient of a residual sum is read in place, never copied): */
  int32_t gsrc_buf, gsrc_c0, gsrc_c;            /* buf >= 0: d(out) is read from this view                                  */
  int32_t dres_buf, dres_c0, dres_c;            /* buf >= 0: this view's gradient is added in the conv's dgrad epilogue     */
  int32_t res_grad_fused, in_grad_fused;        /* no copy of d(out) into d(res) / d(in)                                    */
} dr_op_info;
DR_API int dr_num_ops(const dr_handle* h);
DR_API int dr_debug_op(const dr_handle* h, int idx, dr_op_info* out);

/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
DR_API int64_t dr_launch_count(const dr_handle* h);
/* how many of those were tcgen05 (tensor-core) conv kernels */
DR_API int64_t dr_tc_launch_count(const dr_handle* h);

/* activation workspace bytes currently allocated */
DR_API size_t dr_workspace_bytes(const dr_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* DENSEREG_H_ */
