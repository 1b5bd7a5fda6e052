// engine.cu -- graph construction, forward / backward schedules and the C-ABI of
// libdensereg_sm100.so (include/densereg.h).
//
// The graph is built once per handle by walking network/um_v1.py:detect_net (:71-185) in TF variable-
// creation order; tensors are NHWC fp32 "views" into one activation arena so every tf.concat on the
// path is a channel-offset write.  Backward is an explicit reverse schedule (no tape): BRN backward
// -> dgrad (the forward conv kernel with rotated/transposed weights) -> wgrad, with a build-time
// tracker deciding overwrite-vs-accumulate for every gradient write.
#include "../../include/densereg.h"
#include "common.cuh"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

// programmatic dependent launch switch (common.cuh): DENSEREG_PDL=0 turns it off; suspended while a CUDA graph is being captured
static int g_pdl_suspend = 0;
bool dr_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("DENSEREG_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1 && g_pdl_suspend == 0;
}
void dr_pdl_suspend(int on) { g_pdl_suspend = on; }

namespace {

struct Layer {
  char name[48];
  int k, stride, cin, cout, brn, relu;
  float wd;
  int64_t w_off, p_off, s_off;
  int64_t aff_off, bstat_off, sum_off, wk_off, wa_off;   // wk: [tap][cout][cin_p], wa: [tap][cin][cout_p] (inner dims padded to 4)
  int in_hw, out_hw;
};

struct LayerDev { int C, brn, kk, cin; long long w_off, p_off, s_off, aff_off, wk_off, wa_off; };

struct Buf { int H, W, C, Cs; size_t off; int raw; };   // Cs = padded channel stride; off in per-crop elements
struct View { int buf = -1; int coff = 0; int C = 0; };

enum OpKind { OP_CONV = 0, OP_POOL = 1, OP_UPADD = 2, OP_MASKCOPY = 3 };

struct GradWrite { int acc = 0; int nfill = 0; View fill[4]; };

// Lanes: the op list is a DAG, not a chain -- the `upper1` branch of every hourglass level is independent of the whole lower path
// (um_v1.py:54-65), the hm3 head of the um head (:137-144), the masked um branch of the unmasked one (:143-149), a projection skip of the
// block's c1->c2 chain (:31-47).  Every op is assigned a lane (= CUDA stream) when the graph is built; a build-time hazard analysis over
// the buffer views each op reads / writes (forward: activations; backward: gradients) yields, per op, the events of other lanes it has to
// wait for and whether it must record one.  Small-grid kernels of different lanes then share the GPU.
constexpr int kLanes = 3;
struct OpPlan { int nwait = 0; int wait_op[kLanes]; int record = 0; };

struct Op {
  OpKind kind;
  int lane = 0;
  int layer = -1;
  View in, out, res;
  int raw = -1;
  int accumulate = 0;
  int dropout_tag = -1;
  int k = 0;
  int need_dgrad = 1;
  GradWrite gw_in, gw_res;
  // Gradient aliasing (backward): the gradient of a residual sum flows unchanged into both addends, so it is never copied --
  //   gsrc            this op reads d(out) from that view instead of its own output's gradient (projection-skip conv: the block output's
  //                   gradient; `upper1` block: the gradient of the hourglass level's upsample+add output)
  //   dres            conv: d(dres) is ADDED in the dgrad epilogue as its residual operand (identity skip: d(block input) = dgrad(c1) + d(block output))
  //   res_grad_fused  conv with a residual: no copy of d(out) into d(res) (the consumers above read d(out) directly)
  //   in_grad_fused   upsample+add: no copy of d(out) into d(in)
  View gsrc, dres;
  int res_grad_fused = 0, in_grad_fused = 0;
};

}  // namespace

struct dr_handle {
  dr_config cfg;
  std::vector<Layer> layers;
  std::vector<Buf> bufs;
  std::vector<Op> ops;
  size_t n_params = 0, n_state = 0, n_aff = 0, n_bstat = 0, n_sums = 0;
  size_t act_per_crop = 0, raw_per_crop = 0, scratch_per_crop = 0;
  int buf_x0 = -1, buf_tiny = -1;
  std::vector<View> uvd_dst;
  View hg_ins0;
  std::vector<View> v_hm, v_hm3, v_um;
  // bound buffers
  float *params = nullptr, *state = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  // owned device memory
  float *act = nullptr, *gact = nullptr, *rawa = nullptr, *scratch = nullptr;
  float *aff = nullptr, *bstat = nullptr, *wdmask = nullptr;
  // 16 B aligned weight copies: wk = K-major [tap][cout][cin], wa = plain [tap][cin][cout]; *_hi/_lo = exact-TF32 split (3xTF32)
  float *wk = nullptr, *wa = nullptr, *wk_hi = nullptr, *wk_lo = nullptr, *wa_hi = nullptr, *wa_lo = nullptr;
  size_t n_wk = 0, n_wa = 0;
  double *sums = nullptr, *sums_bw = nullptr, *loss_acc = nullptr;
  unsigned int* counters = nullptr; unsigned int* counters_bw = nullptr;      // one per layer: last-block-done counters of the fused stats+finalize kernel
  int32_t* clamp_dev = nullptr;
  void* crop_scratch = nullptr; size_t crop_scratch_cap = 0;
  LayerDev* ltab = nullptr;
  int cap_B = 0; bool cap_train = false;
  size_t ws_bytes = 0;
  int64_t launches = 0;
  std::string err;
  int precision = 0;
  int64_t tc_launches = 0;
  // aligned / split weight copies are rebuilt only when the parameters changed (dr_optimizer_step, dr_init_params, dr_bind,
  // dr_params_changed) or another arithmetic mode asks for them -- not once per micro-batch
  bool weights_dirty = true; int prepped_precision = -1; bool prep_once = true;
  // two-level accumulation of the 3xTF32 convs (conv_tc.cu): k-blocks per partial accumulator.  Inference (the path whose joint positions
  // are compared with the reference in mm) sums ONE k-block = 32 input channels (12 MMAs) inside the tensor core; training keeps one level
  // (its gradients sit on the fp32 noise floor of the graph either way, and the CTA-pair kernel has no room for a running sum).
  // DENSEREG_TC_CHUNK_EVAL overrides the inference setting (0 = one level, "throughput mode").
  int chunk_eval = 1, chunk_min_kb = 0;
  // backward: filter gradients run on a side stream (they only feed the optimiser) so that they fill the SMs the small
  // BRN / low-resolution kernels of the main stream leave idle; d(raw) scratch is triple-buffered for that
  static const int kSlotsPerLane = 3;
  static const int kScratchSlots = kSlotsPerLane * kLanes;
  cudaStream_t wgrad_stream = nullptr, wgrad_stream2 = nullptr;   // filter gradients alternate between two side streams (DENSEREG_WGRAD_STREAMS=1: one)
  int wgrad_streams = 2;
  cudaEvent_t ev_ready[kSlotsPerLane * kLanes] = {}, ev_wdone[kSlotsPerLane * kLanes] = {}, ev_join = nullptr, ev_join2 = nullptr;
  // lanes (see OpPlan): lane 0 is the caller's stream
  bool lanes_on = true;
  cudaStream_t lane_stream[kLanes] = {};
  std::vector<OpPlan> plan_fwd, plan_bwd;
  std::vector<cudaEvent_t> ev_fwd, ev_bwd;
  cudaEvent_t ev_pass_start = nullptr, ev_lane_done[kLanes] = {};
  bool side_stream = true;
  bool tc_pair = true;       // CTA-pair (cta_group::2) 3xTF32 conv kernel for the big layers (dr_config.reserved[1] < 0 or DENSEREG_TC_PAIR=0: off)
  // inference CUDA graph (dr_config.reserved[0] != 0): the ~170 launches of dr_infer are captured once per (batch, pointers) key
  // on an internal stream and replayed with cudaGraphLaunch on the caller's stream -> B=1 latency is no longer launch-bound
  struct InferGraph { int B = 0; const void *dm = nullptr, *cfg = nullptr, *com = nullptr; void *xyz = nullptr, *top5 = nullptr;
                      cudaGraphExec_t exec = nullptr; int warm = 0; };
  InferGraph infer_graph;
  cudaStream_t capture_stream = nullptr;
  // in-library data-parallel communicator (dr_comm_init): ONE NCCL all-reduce(sum) of the flat gradient per optimiser step, issued in
  // buckets on `comm_stream` while the backward pass of the step's LAST micro-batch is still running (dr_comm_overlap_next_backward)
  void* nccl_comm = nullptr; int comm_rank = 0, comm_world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_bucket_main = nullptr, ev_bucket_side = nullptr, ev_bucket_side2 = nullptr, ev_comm_done = nullptr;
  struct Bucket { int64_t lo, hi; int first_op; };       // flat range [lo,hi) is final once the reverse walk has finished op `first_op`
  std::vector<Bucket> buckets;
  bool overlap_armed = false, reduced_in_backward = false;
  int64_t allreduce_calls = 0;
  // per-launch timing of the conv-type kernels (dr_trace): bench.py's per-class roofline
  bool trace_on = false;
  std::vector<dr_trace_rec> trace;
  float b1p = 1.0f, b2p = 1.0f; int64_t pow_step = 0;      // Adam beta powers (fp32, advanced once per step like TF's variables)
  // micro-batch pipeline (dr_config.reserved[2] == 2): the sub_batch micro-batches of an optimiser step (train_single_gpu.py:140-148) only
  // meet in the BRN moving statistics (written by the forward pass) and in the gradient buffer (written by the backward pass), so the
  // FORWARD pass of micro-batch i+1 may run next to the BACKWARD pass of micro-batch i.  `twin` is a second handle (own activation /
  // gradient arenas, BRN batch statistics, lanes, filter-gradient streams) bound to the SAME parameter / state / gradient buffers and
  // sharing this handle's tensor-core weight copies; micro-batches alternate between the two on two internal streams, ordered by events:
  // forward(i+1) after forward(i) (BRN state order = the reference's), backward(i+1) after backward(i) (exclusive gradient accumulation).
  dr_handle* twin = nullptr; bool is_twin = false;
  cudaStream_t pipe_stream[2] = {};
  cudaEvent_t ev_pipe[7] = {};           // indexed by the PE_* ids of the pipeline plan (pipe_plan_micro_batch): in, fwd[2], bwd[2], loss[2]
  int pipe_next = 0; bool pipe_pending = false;
  int64_t pipe_twin_runs = 0;
  // DENSEREG_CHAIN_PRIO=1: the streams of the dependency chain (pipeline streams, lanes > 0) get the greatest stream priority, so that when SMs
  // free up the block scheduler serves the chain before the filter-gradient side streams (which only feed the optimiser)
  int chain_prio = 0;
};

namespace {

#define CUDA_TRY(h, expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__);                  \
      return DR_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

inline int fail(dr_handle* h, int code, const std::string& msg) { h->err = msg; return code; }

const int kTwinMark = -7777;     // dr_config.reserved[2] of the pipeline's second handle (never creates a twin of its own)

// add the launch count of a run_conv / run_wgrad call, or propagate its (negative) error code
#define RUN_TRY(acc, expr)                  \
  do {                                      \
    const int n__ = (expr);                 \
    if (n__ < 0) return n__;                \
    (acc) += n__;                           \
  } while (0)

int same_pad_before(int n, int k, int s) {
  int out = (n + s - 1) / s;
  int total = (out - 1) * s + k - n;
  if (total < 0) total = 0;
  return total / 2;
}

// ---------------------------------------------------------------------------------------------
// graph builder
// ---------------------------------------------------------------------------------------------
struct Builder {
  dr_handle* h;
  int F, J, S;
  bool plan_overflow = false;
  int cur_lane = 0;          // lane of the ops being appended
  int new_buf(int hw, int C, int raw = 0) {
    Buf b; b.H = hw; b.W = hw; b.C = C; b.Cs = C == 1 ? 1 : (C + 3) / 4 * 4; b.raw = raw;   // 1-channel maps stay dense
    size_t& total = raw ? h->raw_per_crop : h->act_per_crop;
    b.off = total; total += (size_t)hw * hw * b.Cs;
    h->bufs.push_back(b);
    return (int)h->bufs.size() - 1;
  }
  View V(int buf) { View v; v.buf = buf; v.coff = 0; v.C = h->bufs[buf].C; return v; }
  View sub(int buf, int coff, int C) { View v; v.buf = buf; v.coff = coff; v.C = C; return v; }
  int add_layer(const std::string& name, int k, int stride, int cin, int cout, int brn, int relu, float wd, int in_hw) {
    Layer L; memset(&L, 0, sizeof(L));
    snprintf(L.name, sizeof(L.name), "%s", name.c_str());
    L.k = k; L.stride = stride; L.cin = cin; L.cout = cout; L.brn = brn; L.relu = relu; L.wd = wd;
    L.in_hw = in_hw; L.out_hw = (in_hw + stride - 1) / stride;
    L.w_off = (int64_t)h->n_params; h->n_params += (size_t)k * k * cin * cout;
    L.p_off = (int64_t)h->n_params; h->n_params += brn ? 2 * cout : cout;
    L.s_off = (int64_t)h->n_state; if (brn) h->n_state += 4 * cout + 4;
    L.aff_off = (int64_t)h->n_aff; h->n_aff += 2 * cout;
    L.bstat_off = (int64_t)h->n_bstat; h->n_bstat += 4 * cout;
    L.sum_off = (int64_t)h->n_sums; h->n_sums += 2 * cout;
    L.wk_off = (int64_t)h->n_wk; h->n_wk += (size_t)k * k * cout * ((cin + 3) / 4 * 4);
    L.wa_off = (int64_t)h->n_wa; h->n_wa += (size_t)k * k * cin * ((cout + 3) / 4 * 4);
    h->layers.push_back(L);
    return (int)h->layers.size() - 1;
  }
  void conv_op(int layer, View in, View out, View res = View(), int accumulate = 0, int dropout_tag = -1) {
    Op o; o.kind = OP_CONV; o.layer = layer; o.in = in; o.out = out; o.res = res; o.lane = cur_lane;
    o.accumulate = accumulate; o.dropout_tag = dropout_tag;
    const Layer& L = h->layers[layer];
    if (L.brn) o.raw = new_buf(L.out_hw, L.cout, 1);
    size_t sc = (size_t)L.out_hw * L.out_hw * ((L.cout + 3) / 4 * 4);
    if (sc > h->scratch_per_crop) h->scratch_per_crop = sc;
    h->ops.push_back(o);
  }
  // skip_lane >= 0: the projection skip (if any) runs on that lane, next to the c1 -> c2 chain
  struct ResOps { int c1, c3, skip; };
  bool alias_grads = true;   // DENSEREG_GRAD_ALIAS=0: round-1 behaviour (copy the gradient of every residual sum)
  ResOps residual(const std::string& name, View in, int cin, int cout, View dest, int hw, int skip_lane = -1) {
    int hc = cin / 2;
    int L1 = add_layer(name + "/c1", 1, 1, cin, hc, 1, 1, 0.0005f, hw);
    int L2 = add_layer(name + "/c2", 3, 1, hc, hc, 1, 1, 0.0005f, hw);
    int L3 = add_layer(name + "/c3", 1, 1, hc, cout, 1, 1, 0.0005f, hw);
    int Ls = cout != cin ? add_layer(name + "/skip", 1, 1, cin, cout, 1, 1, 0.0005f, hw) : -1;
    int t1 = new_buf(hw, hc), t2 = new_buf(hw, hc);
    ResOps r{-1, -1, -1};
    conv_op(L1, in, V(t1)); r.c1 = (int)h->ops.size() - 1;
    conv_op(L2, V(t1), V(t2));
    View sk = in;
    if (Ls >= 0) {
      int sb = new_buf(hw, cout);
      const int keep = cur_lane;
      if (skip_lane >= 0) cur_lane = skip_lane;
      conv_op(Ls, in, V(sb)); sk = V(sb); r.skip = (int)h->ops.size() - 1;
      cur_lane = keep;
    }
    conv_op(L3, V(t2), dest, sk); r.c3 = (int)h->ops.size() - 1;
    if (alias_grads) {
      h->ops[r.c3].res_grad_fused = 1;
      if (r.skip >= 0) h->ops[r.skip].gsrc = dest;          // the skip conv back-propagates the block output's gradient directly
      else h->ops[r.c1].dres = dest;                        // identity skip: added in c1's dgrad epilogue
    }
    return r;
  }
  void hourglass(const std::string& name, int n, View x, View dest, int hw) {
    char tag[16]; snprintf(tag, sizeof(tag), "/n%d", n);
    std::string p = name + tag;
    int up1 = new_buf(hw, F);
    ResOps up;
    { const int keep = cur_lane; cur_lane = 1 + (n & 1);                      // the upper branch of level n overlaps the whole lower path
      up = residual(p + "/upper1", x, F, F, V(up1), hw);
      cur_lane = keep; }
    int pl = new_buf(hw / 2, F);
    { Op o; o.kind = OP_POOL; o.lane = cur_lane; o.in = x; o.out = V(pl); o.k = 3; h->ops.push_back(o); }
    int low1 = new_buf(hw / 2, F);
    residual(p + "/lower1", V(pl), F, F, V(low1), hw / 2);
    int low2 = low1;
    if (n > 1) { low2 = new_buf(hw / 2, F); hourglass(name, n - 1, V(low1), V(low2), hw / 2); }
    int low3 = new_buf(hw / 2, F);
    residual(p + "/lower3", V(low2), F, F, V(low3), hw / 2);
    { Op o; o.kind = OP_UPADD; o.lane = cur_lane; o.in = V(up1); o.res = V(low3); o.out = dest;
      if (alias_grads) {                                                      // d(up1) == d(dest): the upper1 block reads it in place
        o.in_grad_fused = 1;
        h->ops[up.c3].gsrc = dest;
        h->ops[up.c1].dres = dest;                                            // (identity skip: F -> F)
      }
      h->ops.push_back(o); }
  }
  void build() {
    { const char* e = getenv("DENSEREG_GRAD_ALIAS"); alias_grads = !(e && e[0] == '0'); }
    F = h->cfg.num_fea; J = h->cfg.num_jnt; S = h->cfg.num_stack;
    const int IN = h->cfg.in_hw, OUT = h->cfg.out_hw;
    h->buf_x0 = new_buf(IN, 1);
    h->buf_tiny = new_buf(OUT, 1);
    // stem: um_v1.py:84-97
    int Lc1 = add_layer("stem/conv_1", 7, 2, 1, 32, 1, 1, 0.0005f, IN);
    int c1 = new_buf(IN / 2, 32);
    conv_op(Lc1, V(h->buf_x0), V(c1));
    h->ops.back().need_dgrad = 0;
    int c2 = new_buf(IN / 2, 64);
    residual("stem/conv_2", V(c1), 32, 64, V(c2), IN / 2, 2);
    int p1 = new_buf(OUT, 64);
    { Op o; o.kind = OP_POOL; o.lane = cur_lane; o.in = V(c2); o.out = V(p1); o.k = 2; h->ops.push_back(o); }
    int c3 = new_buf(OUT, 64);
    residual("stem/conv_3", V(p1), 64, 64, V(c3), OUT);
    int c4 = new_buf(OUT, F);
    residual("stem/conv_4", V(c3), 64, F, V(c4), OUT, 2);
    View hg_ins = V(c4);
    h->hg_ins0 = hg_ins;
    for (int s = 0; s < S; ++s) {
      char ps[16]; snprintf(ps, sizeof(ps), "s%d", s);
      std::string p = ps;
      int big = new_buf(OUT, F + 5 * J);
      View hg_outs = sub(big, 0, F), hm = sub(big, F, J), hm3 = sub(big, F + J, J), um = sub(big, F + 2 * J, 3 * J);
      View cat = sub(big, 0, F + 2 * J), tmp_out = sub(big, F, 5 * J);
      h->v_hm.push_back(hm); h->v_hm3.push_back(hm3); h->v_um.push_back(um);
      hourglass(p + "/hg", 4, hg_ins, hg_outs, OUT);                         // um_v1.py:125
      int llr = new_buf(OUT, F);
      residual(p + "/ll_res", hg_outs, F, F, V(llr), OUT);                   // :127
      int lluvd = new_buf(OUT, F + 3);
      View ll = sub(lluvd, 0, F);
      h->uvd_dst.push_back(sub(lluvd, F, 3));
      conv_op(add_layer(p + "/ll", 1, 1, F, F, 1, 1, 0.0005f, OUT), V(llr), ll);              // :128-131
      conv_op(add_layer(p + "/hm_out", 1, 1, F, J, 0, 0, 0.0005f, OUT), ll, hm);               // :133-135
      int h3 = new_buf(OUT, 128);
      residual(p + "/hm3_res", V(lluvd), F + 3, 128, V(h3), OUT, 2);         // :137-138
      conv_op(add_layer(p + "/hm3_out", 1, 1, 128, J, 0, 0, 0.0005f, OUT), V(h3), hm3);        // :139-141
      int u1 = new_buf(OUT, 256);
      residual(p + "/um_res1", cat, F + 2 * J, 256, V(u1), OUT, 2);          // :143-144
      int combin = new_buf(OUT, 512);
      residual(p + "/um_res2", V(u1), 256, 256, sub(combin, 0, 256), OUT);
      int catm = new_buf(OUT, F + 2 * J);
      cur_lane = 1;                                                           // the masked branch runs next to the unmasked one
      { Op o; o.kind = OP_MASKCOPY; o.lane = cur_lane; o.in = cat; o.out = V(catm); h->ops.push_back(o); }      // :146-148
      int m1 = new_buf(OUT, 256);
      residual(p + "/um_mask_res1", V(catm), F + 2 * J, 256, V(m1), OUT, 2);  // :149
      residual(p + "/um_mask_res2", V(m1), 256, 256, sub(combin, 256, 256), OUT);
      cur_lane = 0;
      int combuvd = new_buf(OUT, 515);
      residual(p + "/um_comb", V(combin), 512, 512, sub(combuvd, 0, 512), OUT);   // :151-152
      h->uvd_dst.push_back(sub(combuvd, 512, 3));                              // :153
      int f1 = new_buf(OUT, 512), f2 = new_buf(OUT, 512);
      conv_op(add_layer(p + "/um_full1", 1, 1, 515, 512, 0, 1, 0.0005f, OUT), V(combuvd), V(f1), View(), 0, 2 * s);      // :155-159
      conv_op(add_layer(p + "/um_full2", 1, 1, 512, 512, 0, 1, 0.0005f, OUT), V(f1), V(f2), View(), 0, 2 * s + 1);    // :160-164
      conv_op(add_layer(p + "/um_out", 1, 1, 512, 3 * J, 0, 0, 0.0005f, OUT), V(f2), um);                              // :166-169
      if (s < S - 1) {                                                          // :174-183
        int nxt = new_buf(OUT, F);
        conv_op(add_layer(p + "/inter_out", 1, 1, 5 * J, F, 0, 0, 0.f, OUT), tmp_out, V(nxt), hg_ins);
        conv_op(add_layer(p + "/inter_ll", 1, 1, F, F, 0, 0, 0.f, OUT), ll, V(nxt), View(), 1);
        hg_ins = V(nxt);
      }
    }
    plan_backward();
    plan_lanes();
  }

  // ---- lane hazard analysis ------------------------------------------------------------------------------------------------
  struct Access { int pos, lane, buf, c0, c1; bool write; };
  // order[pos] = op index; accesses(op, out): the views the op touches in this pass.  For every op: wait for the LATEST conflicting
  // op (read-after-write, write-after-read, write-after-write on overlapping channel ranges of one buffer) of each other lane, unless
  // this lane already waited for that op or a later one of the same lane (events of one stream are ordered).
  template <class F>
  void plan_pass(const std::vector<int>& order, F accesses, std::vector<OpPlan>& plan) {
    plan.assign(h->ops.size(), OpPlan());
    std::vector<std::vector<Access>> hist(2 * h->bufs.size());
    int synced[kLanes][kLanes];
    for (int a = 0; a < kLanes; ++a) for (int b = 0; b < kLanes; ++b) synced[a][b] = -1;
    for (int pos = 0; pos < (int)order.size(); ++pos) {
      const int oi = order[pos];
      const int lane = h->ops[oi].lane;
      std::vector<Access> acc;
      accesses(h->ops[oi], acc);
      int need[kLanes]; for (int l = 0; l < kLanes; ++l) need[l] = -1;
      for (Access& a : acc) {
        a.pos = pos; a.lane = lane;
        for (const Access& q : hist[a.buf])
          if (q.lane != lane && (a.write || q.write) && a.c0 < q.c1 && q.c0 < a.c1 && q.pos > need[q.lane]) need[q.lane] = q.pos;
      }
      OpPlan& pl = plan[oi];
      for (int l = 0; l < kLanes; ++l)
        if (l != lane && need[l] > synced[lane][l]) {
          pl.wait_op[pl.nwait++] = order[need[l]];
          plan[order[need[l]]].record = 1;
          synced[lane][l] = need[l];
        }
      for (const Access& a : acc) hist[a.buf].push_back(a);
    }
  }
  void plan_lanes() {
    const int nb = (int)h->bufs.size();
    auto rd = [](std::vector<Access>& v, int buf, const View& w) { if (w.buf >= 0) v.push_back(Access{0, 0, buf, w.coff, w.coff + w.C, false}); };
    auto wr = [](std::vector<Access>& v, int buf, const View& w) { if (w.buf >= 0) v.push_back(Access{0, 0, buf, w.coff, w.coff + w.C, true}); };
    std::vector<int> fo(h->ops.size()), bo(h->ops.size());
    for (size_t i = 0; i < h->ops.size(); ++i) { fo[i] = (int)i; bo[i] = (int)(h->ops.size() - 1 - i); }
    // forward: activation arena (buffer ids as they are)
    plan_pass(fo, [&](const Op& o, std::vector<Access>& v) {
      rd(v, o.in.buf, o.in);
      if (o.res.buf >= 0) rd(v, o.res.buf, o.res);
      wr(v, o.out.buf, o.out);                                   // also covers `accumulate` (read-modify-write)
      if (o.raw >= 0) { View r; r.buf = o.raw; r.coff = 0; r.C = h->bufs[o.raw].Cs; wr(v, o.raw, r); }
    }, h->plan_fwd);
    // backward: gradient arena (buffer id + nb); forward activations are read-only here and complete before the pass starts
    plan_pass(bo, [&](const Op& o, std::vector<Access>& v) {
      const View& gv = o.gsrc.buf >= 0 ? o.gsrc : o.out;
      rd(v, nb + gv.buf, gv);
      if (o.kind == OP_CONV) {
        if (o.res.buf >= 0 && !o.res_grad_fused) wr(v, nb + o.res.buf, o.res);
        if (o.dres.buf >= 0) rd(v, nb + o.dres.buf, o.dres);
        if (o.need_dgrad) wr(v, nb + o.in.buf, o.in);
      } else {
        if (!(o.kind == OP_UPADD && o.in_grad_fused)) wr(v, nb + o.in.buf, o.in);
        if (o.kind == OP_UPADD) wr(v, nb + o.res.buf, o.res);
      }
    }, h->plan_bwd);
  }

  // decide overwrite / accumulate for each gradient write of the reverse schedule
  void plan_backward() {
    std::vector<std::vector<char>> written(h->bufs.size());
    for (size_t i = 0; i < h->bufs.size(); ++i) written[i].assign(h->bufs[i].Cs, 0);
    auto mark = [&](const View& v) { for (int c = 0; c < v.C; ++c) written[v.buf][v.coff + c] = 1; };
    auto plan = [&](const View& v) {
      GradWrite g;
      int nw = 0;
      for (int c = 0; c < v.C; ++c) nw += written[v.buf][v.coff + c];
      if (nw == 0) { g.acc = 0; }
      else if (nw == v.C) { g.acc = 1; }
      else {
        g.acc = 1;
        int c = 0;
        while (c < v.C) {
          if (written[v.buf][v.coff + c]) { ++c; continue; }
          int c0 = c;
          while (c < v.C && !written[v.buf][v.coff + c]) ++c;
          if (g.nfill < 4) { View f; f.buf = v.buf; f.coff = v.coff + c0; f.C = c - c0; g.fill[g.nfill++] = f; }
          else plan_overflow = true;               // more unwritten channel gaps than GradWrite can zero-fill: dr_create fails
        }
      }
      mark(v);
      return g;
    };
    for (size_t s = 0; s < h->v_hm.size(); ++s) { mark(h->v_hm[s]); mark(h->v_hm3[s]); mark(h->v_um[s]); }   // loss kernel
    for (int i = (int)h->ops.size() - 1; i >= 0; --i) {
      Op& o = h->ops[i];
      switch (o.kind) {
        case OP_CONV:
          if (o.res.buf >= 0 && !o.res_grad_fused) o.gw_res = plan(o.res);
          if (o.need_dgrad) o.gw_in = plan(o.in);
          break;
        case OP_POOL: o.gw_in = plan(o.in); break;
        case OP_UPADD: if (!o.in_grad_fused) o.gw_in = plan(o.in); o.gw_res = plan(o.res); break;
        case OP_MASKCOPY: o.gw_in = plan(o.in); break;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// table-driven helper kernels
// ---------------------------------------------------------------------------------------------
__global__ void fold_all_kernel(const LayerDev* __restrict__ t, const float* __restrict__ params, const float* __restrict__ state,
                                float* __restrict__ aff) {
  pdl_trigger(); pdl_wait();      // programmatic dependent launch (common.cuh)
  const LayerDev L = t[blockIdx.x];
  const float* pb = params + L.p_off;
  float* a = aff + L.aff_off;
  for (int c = threadIdx.x; c < L.C; c += blockDim.x) {
    if (L.brn) {                                           // network/slim/ops.py:173-180
      const float* st = state + L.s_off;
      float inv = (1.0f / sqrtf(st[L.C + c] + 0.001f)) * pb[L.C + c];
      a[c] = inv; a[L.C + c] = pb[c] - st[c] * inv;
    } else {
      a[c] = 1.0f; a[L.C + c] = pb[c];
    }
  }
}

// per-layer aligned weight copies: wa[tap][c][n] = w, wk[tap][n][c] = w^T (K-major for the tensor-core path and the
// SIMT dgrad); with split != 0 also the exact-TF32 hi / lo = w - hi parts of both.
__global__ void prep_weights_kernel(const LayerDev* __restrict__ t, const float* __restrict__ params, float* __restrict__ wk,
                                    float* __restrict__ wa, float* __restrict__ wk_hi, float* __restrict__ wk_lo,
                                    float* __restrict__ wa_hi, float* __restrict__ wa_lo, int split) {
  const LayerDev L = t[blockIdx.x];
  const float* w = params + L.w_off;
  const size_t n = (size_t)L.kk * L.cin * L.C;
  const int cin_p = (L.cin + 3) / 4 * 4, cout_p = (L.C + 3) / 4 * 4;
  for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.y * blockDim.x) {
    const int n_o = (int)(i % L.C); const size_t r = i / L.C; const int c = (int)(r % L.cin); const int tap = (int)(r / L.cin);
    const size_t ik = L.wk_off + ((size_t)tap * L.C + n_o) * cin_p + c;
    const size_t ia = L.wa_off + ((size_t)tap * L.cin + c) * cout_p + n_o;
    const float v = w[i];
    wa[ia] = v; wk[ik] = v;
    if (split) {
      uint32_t hb, lb;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
      const float hi = __uint_as_float(hb);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(v - hi));
      const float lo = __uint_as_float(lb);
      wa_hi[ia] = hi; wa_lo[ia] = lo; wk_hi[ik] = hi; wk_lo[ik] = lo;
    }
  }
}

__global__ void init_state_kernel(const LayerDev* __restrict__ t, float* __restrict__ params, float* __restrict__ state) {
  const LayerDev L = t[blockIdx.x];
  float* pb = params + L.p_off;
  for (int c = threadIdx.x; c < L.C; c += blockDim.x) {
    pb[c] = 0.f;                                           // beta / biases = 0
    if (L.brn) {
      pb[L.C + c] = 1.f;                                   // gamma = 1
      float* st = state + L.s_off;
      st[c] = 0.f; st[L.C + c] = 1.f; st[2 * L.C + c] = 0.f; st[3 * L.C + c] = 0.f;
    }
  }
  if (L.brn && threadIdx.x == 0) {
    float* st = state + L.s_off + 4 * L.C;
    st[0] = 1.f; st[1] = 0.f; st[2] = 0.f; st[3] = 0.f;    // r_max, d_max, curr_t, local_step
  }
}

// ---------------------------------------------------------------------------------------------
// execution helpers
// ---------------------------------------------------------------------------------------------
struct Exec {
  dr_handle* h; int B; cudaStream_t st;
  float* ptr(const View& v) const {
    const Buf& b = h->bufs[v.buf];
    float* base = b.raw ? h->rawa : h->act;
    return base + b.off * h->cap_B + v.coff;
  }
  float* gptr(const View& v) const {
    const Buf& b = h->bufs[v.buf];
    return h->gact + b.off * h->cap_B + v.coff;
  }
  int cs(const View& v) const { return h->bufs[v.buf].Cs; }
  size_t npix(const View& v) const { const Buf& b = h->bufs[v.buf]; return (size_t)B * b.H * b.W; }
  View whole(int buf) const { View v; v.buf = buf; v.coff = 0; v.C = h->bufs[buf].C; return v; }
};

// DENSEREG_TRACE=1 or dr_trace(h, 1): every conv / wgrad launch is timed on its own (events + sync, so launches are serialised) and reported on stderr as
//   TRACE <conv|dgrad|wgrad> B H Cin Cout k <kernel: tc|pair|simt> <ms>      -- the per-layer time table of a step (tools/layer_times.py)
static bool trace_env() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("DENSEREG_TRACE"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}
struct TraceScope {
  cudaEvent_t e0 = nullptr, e1 = nullptr; cudaStream_t st; bool on; dr_handle* h;
  TraceScope(dr_handle* hh, cudaStream_t s) : st(s), on(trace_env() || hh->trace_on), h(hh) {
    if (on) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
  }
  void done(const char* what, int B, int H, int Cin, int Cout, int k, const char* kern) {
    if (!on) return;
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    if (trace_env()) fprintf(stderr, "TRACE %s %d %d %d %d %d %s %.4f\n", what, B, H, Cin, Cout, k, kern, ms);
    if (h->trace_on) {
      dr_trace_rec r; memset(&r, 0, sizeof(r));
      r.kind = what[0] == 'c' ? 0 : (what[0] == 'd' ? 1 : 2);
      r.B = B; r.hw = H; r.cin = Cin; r.cout = Cout; r.k = k;
      r.kernel = kern[0] == 's' ? 0 : (kern[0] == 't' ? 1 : 2);
      r.ms = ms;
      h->trace.push_back(r);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
};

// A problem the tensor-core path accepts (conv_tc_eligible / wgrad_tc_eligible) MUST run there: a failed tensor-map encode or launch is an
// error (negative return, message in dr_last_error), never a quiet switch to the FFMA kernels.
int run_conv(dr_handle* h, const ConvProblem& p, int precision, cudaStream_t st) {
  TraceScope tr(h, st);
  if (precision != DR_PREC_FP32 && conv_tc_eligible(p)) {
    ConvProblem q = p;
    q.pair = p.pair ? p.pair : (h->tc_pair ? 1 : 0);
    int n = launch_conv_tc(q, precision == DR_PREC_TF32X3, st);
    if (n <= 0) {
      char msg[256];
      snprintf(msg, sizeof(msg), "tensor-core conv launch failed (B=%d HW=%d Cin=%d Cout=%d k=%d): %s", p.B, p.H, p.Cin, p.Cout, p.k, tc_last_error());
      return fail(h, DR_ERR_CUDA, msg);
    }
    h->tc_launches += n;
    tr.done(p.flip_taps ? "dgrad" : "conv", p.B, p.H, p.Cin, p.Cout, p.k, (precision == DR_PREC_TF32X3 && conv_tc_pair_wanted(q)) ? "pair" : "tc");
    return n;
  }
  int n = launch_conv_simt(p, st);
  tr.done(p.flip_taps ? "dgrad" : "conv", p.B, p.H, p.Cin, p.Cout, p.k, "simt");
  return n;
}

int run_wgrad(dr_handle* h, const WgradProblem& p, int precision, cudaStream_t st) {
  TraceScope tr(h, st);
  if (precision != DR_PREC_FP32 && wgrad_tc_eligible(p)) {
    int n = launch_wgrad_tc(p, precision == DR_PREC_TF32X3, st);
    if (n <= 0) {
      char msg[256];
      snprintf(msg, sizeof(msg), "tensor-core wgrad launch failed (B=%d HW=%d Cin=%d Cout=%d k=%d): %s", p.B, p.H, p.Cin, p.Cout, p.k, tc_last_error());
      return fail(h, DR_ERR_CUDA, msg);
    }
    h->tc_launches += n; tr.done("wgrad", p.B, p.H, p.Cin, p.Cout, p.k, "tc");
    return n;
  }
  int n = launch_wgrad_simt(p, st);
  tr.done("wgrad", p.B, p.H, p.Cin, p.Cout, p.k, "simt");
  return n;
}

// (re)build the aligned weight copies from the bound parameters; called once per forward pass
int prep_weights(dr_handle* h, int precision, cudaStream_t st) {
  const size_t bk = h->n_wk * sizeof(float), ba = h->n_wa * sizeof(float);
  if (!h->wk) {      // pad elements are zeroed once here and never written again
    CUDA_TRY(h, cudaMalloc(&h->wk, bk)); CUDA_TRY(h, cudaMalloc(&h->wa, ba)); h->ws_bytes += bk + ba;
    CUDA_TRY(h, cudaMemsetAsync(h->wk, 0, bk, st)); CUDA_TRY(h, cudaMemsetAsync(h->wa, 0, ba, st));
  }
  const int split = precision == DR_PREC_TF32X3;   // weights are pre-split (hi/lo) here; activations are split in shared memory
  if (split && !h->wk_hi) {
    CUDA_TRY(h, cudaMalloc(&h->wk_hi, bk)); CUDA_TRY(h, cudaMalloc(&h->wk_lo, bk));
    CUDA_TRY(h, cudaMalloc(&h->wa_hi, ba)); CUDA_TRY(h, cudaMalloc(&h->wa_lo, ba)); h->ws_bytes += 2 * (bk + ba);
    CUDA_TRY(h, cudaMemsetAsync(h->wk_hi, 0, bk, st)); CUDA_TRY(h, cudaMemsetAsync(h->wk_lo, 0, bk, st));
    CUDA_TRY(h, cudaMemsetAsync(h->wa_hi, 0, ba, st)); CUDA_TRY(h, cudaMemsetAsync(h->wa_lo, 0, ba, st));
  }
  prep_weights_kernel<<<dim3((unsigned)h->layers.size(), 48), 256, 0, st>>>(h->ltab, h->params, h->wk, h->wa, h->wk_hi, h->wk_lo,
                                                                           h->wa_hi, h->wa_lo, split);
  ++h->launches;
  return DR_OK;
}

// prep_weights only when needed: parameters changed since the last build, or the mode needs the hi / lo split and only the plain copies exist
int ensure_prepped(dr_handle* h, int precision, cudaStream_t st) {
  const int need = precision == DR_PREC_TF32X3 ? 2 : 1;
  const int have = h->prepped_precision == DR_PREC_TF32X3 ? 2 : (h->prepped_precision >= 0 ? 1 : 0);
  if (h->prep_once && !h->weights_dirty && have >= need) return DR_OK;
  const int prec = (have > need && h->prep_once) ? h->prepped_precision : precision;     // keep the split copies current once they exist
  int rc = prep_weights(h, prec, st);
  if (rc) return rc;
  h->weights_dirty = false; h->prepped_precision = prec;
  return DR_OK;
}

// weight operands of one conv for the forward pass / for dgrad
void set_fwd_weights(dr_handle* h, const Layer& L, int precision, ConvProblem& p) {
  p.w = h->params + L.w_off; p.w_ld = L.cout; p.flip_taps = 0;
  if (precision != DR_PREC_FP32 && h->wk) {
    const bool x3 = precision == DR_PREC_TF32X3;
    p.w_kmajor = (x3 ? h->wk_hi : h->wk) + L.wk_off;        // [tap][cout][cin_p]
    p.w_kmajor_lo = x3 ? h->wk_lo + L.wk_off : nullptr;
    p.wk_ld = (L.cin + 3) / 4 * 4;
  }
}
void set_dgrad_weights(dr_handle* h, const Layer& L, int precision, ConvProblem& p) {
  p.w = h->wk + L.wk_off; p.w_ld = (L.cin + 3) / 4 * 4; p.flip_taps = 1;   // rows (tap, cout), cin contiguous (padded)
  if (precision != DR_PREC_FP32) {
    const bool x3 = precision == DR_PREC_TF32X3;
    p.w_kmajor = (x3 ? h->wa_hi : h->wa) + L.wa_off;        // K-major for dgrad: [tap][cin][cout_p]
    p.w_kmajor_lo = x3 ? h->wa_lo + L.wa_off : nullptr;
    p.wk_ld = (L.cout + 3) / 4 * 4;
  }
}

int ensure_workspace(dr_handle* h, int B, bool train) {
  if (B <= 0 || B > h->cfg.max_batch) return fail(h, DR_ERR_ARG, "batch exceeds dr_config.max_batch");
  if (h->cap_B >= h->cfg.max_batch && (h->cap_train || !train)) return DR_OK;
  const int cap = h->cfg.max_batch;
  if (!h->act) {
    size_t bytes = h->act_per_crop * cap * sizeof(float);
    CUDA_TRY(h, cudaMalloc(&h->act, bytes)); h->ws_bytes += bytes;
    CUDA_TRY(h, cudaMemset(h->act, 0, bytes));
  }
  if (train && !h->cap_train) {
    size_t bytes = h->act_per_crop * cap * sizeof(float);
    CUDA_TRY(h, cudaMalloc(&h->gact, bytes)); h->ws_bytes += bytes;
    CUDA_TRY(h, cudaMemset(h->gact, 0, bytes));
    bytes = h->raw_per_crop * cap * sizeof(float);
    CUDA_TRY(h, cudaMalloc(&h->rawa, bytes)); h->ws_bytes += bytes;
    bytes = h->scratch_per_crop * cap * sizeof(float) * dr_handle::kScratchSlots;
    CUDA_TRY(h, cudaMalloc(&h->scratch, bytes)); h->ws_bytes += bytes;
    { const char* env = getenv("DENSEREG_SIDE_STREAM"); h->side_stream = !(env && env[0] == '0'); }
    if (h->side_stream) {
      CUDA_TRY(h, cudaStreamCreateWithFlags(&h->wgrad_stream, cudaStreamNonBlocking));
      CUDA_TRY(h, cudaStreamCreateWithFlags(&h->wgrad_stream2, cudaStreamNonBlocking));
      { const char* e = getenv("DENSEREG_WGRAD_STREAMS"); h->wgrad_streams = (e && e[0] == '1') ? 1 : 2; }
      for (int i = 0; i < dr_handle::kScratchSlots; ++i) {
        if (h->ev_ready[i]) continue;
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_ready[i], cudaEventDisableTiming));
        CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_wdone[i], cudaEventDisableTiming));
      }
      CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
      CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming));
    }
    h->cap_train = true;
  }
  h->cap_B = cap;
  return DR_OK;
}

// run-time side of the lane plan (OpPlan): which stream an op runs on, the event waits before it, the event record after it
struct LaneCtx {
  dr_handle* h; cudaStream_t st; bool on; const std::vector<OpPlan>* plan; std::vector<cudaEvent_t>* ev;
  bool started[kLanes];
  int begin() {                       // everything enqueued on `st` so far (inputs, memsets, previous passes) precedes every lane
    for (int l = 0; l < kLanes; ++l) started[l] = false;
    started[0] = true;
    if (on) CUDA_TRY(h, cudaEventRecord(h->ev_pass_start, st));
    return DR_OK;
  }
  cudaStream_t stream(int lane) const { return (on && lane > 0) ? h->lane_stream[lane] : st; }
  int enter(int oi, cudaStream_t* out) {
    const int lane = on ? h->ops[oi].lane : 0;
    cudaStream_t s = stream(lane);
    if (on) {
      if (!started[lane]) { CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_pass_start, 0)); started[lane] = true; }
      const OpPlan& pl = (*plan)[oi];
      for (int i = 0; i < pl.nwait; ++i) CUDA_TRY(h, cudaStreamWaitEvent(s, (*ev)[pl.wait_op[i]], 0));
    }
    *out = s;
    return DR_OK;
  }
  int leave(int oi) {
    if (on && (*plan)[oi].record) CUDA_TRY(h, cudaEventRecord((*ev)[oi], stream(h->ops[oi].lane)));
    return DR_OK;
  }
  int join() {                        // the caller's stream continues only after every lane has drained
    if (!on) return DR_OK;
    for (int l = 1; l < kLanes; ++l)
      if (started[l]) {
        CUDA_TRY(h, cudaEventRecord(h->ev_lane_done[l], h->lane_stream[l]));
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_lane_done[l], 0));
      }
    return DR_OK;
  }
};

int ensure_lanes(dr_handle* h) {
  if (!h->lanes_on || h->ev_pass_start) return DR_OK;
  for (int l = 1; l < kLanes; ++l) CUDA_TRY(h, cudaStreamCreateWithPriority(&h->lane_stream[l], cudaStreamNonBlocking, h->chain_prio));
  for (int l = 0; l < kLanes; ++l) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_lane_done[l], cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_pass_start, cudaEventDisableTiming));
  h->ev_fwd.assign(h->ops.size(), nullptr); h->ev_bwd.assign(h->ops.size(), nullptr);
  for (size_t i = 0; i < h->ops.size(); ++i) {
    if (h->plan_fwd[i].record) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_fwd[i], cudaEventDisableTiming));
    if (h->plan_bwd[i].record) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_bwd[i], cudaEventDisableTiming));
  }
  return DR_OK;
}

int forward_impl(dr_handle* h, int B, const float* dm_mm, const float* coms, int training, int update_state,
                 uint64_t dropout_seed, cudaStream_t st0) {
  if (!h->params || !h->state) return fail(h, DR_ERR_STATE, "dr_bind() not called");
  int rc = ensure_workspace(h, B, training != 0);
  if (rc) return rc;
  rc = ensure_lanes(h);
  if (rc) return rc;
  cudaStream_t st = st0;
  Exec X{h, B, st};
  const int IN = h->cfg.in_hw, OUT = h->cfg.out_hw;
  int nl = 0;
  // norm_dm + uvd / tiny_dm
  float* x0 = X.ptr(X.whole(h->buf_x0));
  float* tiny = X.ptr(X.whole(h->buf_tiny));
  nl += launch_norm_dm(B, IN, dm_mm, coms, x0, st);
  UvdDst ud; ud.n = (int)h->uvd_dst.size();
  if (ud.n > 8) return fail(h, DR_ERR_UNSUPPORTED, "num_stack > 4 not supported");
  for (int i = 0; i < ud.n; ++i) { ud.p[i] = X.ptr(h->uvd_dst[i]); ud.cs[i] = X.cs(h->uvd_dst[i]); }
  nl += launch_make_uvd(B, IN, OUT, x0, tiny, ud, st);
  if (training || h->precision != DR_PREC_FP32) { rc = ensure_prepped(h, h->precision, st); if (rc) return rc; }
  if (training) {
    CUDA_TRY(h, cudaMemsetAsync(h->sums, 0, h->n_sums * sizeof(double), st));
    CUDA_TRY(h, cudaMemsetAsync(h->counters, 0, h->layers.size() * sizeof(unsigned int), st));
  } else {
    dr_launch(fold_all_kernel, dim3((unsigned)h->layers.size()), dim3(128), 0, st, h->ltab, h->params, h->state, h->aff); ++nl;
  }
  LaneCtx lanes{h, st0, h->lanes_on, &h->plan_fwd, &h->ev_fwd, {}};
  rc = lanes.begin();
  if (rc) return rc;
  for (size_t oi = 0; oi < h->ops.size(); ++oi) {
    const Op& o = h->ops[oi];
    rc = lanes.enter((int)oi, &st);          // `st` = this op's lane stream from here on
    if (rc) return rc;
    switch (o.kind) {
      case OP_CONV: {
        const Layer& L = h->layers[o.layer];
        ConvProblem p; memset(&p, 0, sizeof(p));
        p.x = X.ptr(o.in); p.x_cs = X.cs(o.in);
        p.B = B; p.H = L.in_hw; p.W = L.in_hw; p.Cin = L.cin; p.Ho = L.out_hw; p.Wo = L.out_hw; p.Cout = L.cout;
        p.k = L.k; p.stride = L.stride; p.pad_t = p.pad_l = same_pad_before(L.in_hw, L.k, L.stride);
        set_fwd_weights(h, L, h->precision, p);
        p.chunk_kb = training ? 0 : h->chunk_eval; p.chunk_min_kb = h->chunk_min_kb;      // two-level accumulation: inference only
        const float* aff = h->aff + L.aff_off;
        const float* res = o.res.buf >= 0 ? X.ptr(o.res) : nullptr;
        const int res_cs = o.res.buf >= 0 ? X.cs(o.res) : 0;
        if (training && L.brn) {
          View rv = X.whole(o.raw);
          p.y = X.ptr(rv); p.y_cs = X.cs(rv);
          double* sums = h->sums + L.sum_off;
          const bool fuse = h->precision != DR_PREC_FP32 && conv_tc_eligible(p);
          if (fuse) {          // statistics + finalize fused into the tcgen05 conv epilogue
            p.stats = sums; p.stats_counter = h->counters + o.layer; p.bn_bg = h->params + L.p_off; p.bn_state = h->state + L.s_off;
            p.bn_aff = h->aff + L.aff_off; p.bn_bstat = h->bstat + L.bstat_off; p.bn_update_state = update_state;
          }
          RUN_TRY(nl, run_conv(h, p, h->precision, st));
          if (!fuse)                                     // SIMT path: separate statistics pass
            nl += launch_channel_stats_finalize(X.npix(rv), L.cout, p.y, p.y_cs, sums, h->counters + o.layer, h->params + L.p_off,
                                                h->state + L.s_off, h->aff + L.aff_off, h->bstat + L.bstat_off, update_state, st);
          nl += launch_brn_apply(X.npix(rv), L.cout, p.y, p.y_cs, aff, L.relu, res, res_cs, X.ptr(o.out), X.cs(o.out), st);
        } else {
          p.y = X.ptr(o.out); p.y_cs = X.cs(o.out);
          if (L.brn) { p.scale = aff; p.shift = aff + L.cout; }
          else { p.scale = nullptr; p.shift = h->params + L.p_off; }
          p.relu = L.relu; p.res = res; p.res_cs = res_cs; p.accumulate = o.accumulate;
          if (training && o.dropout_tag >= 0) { p.dropout = 1; p.drop_seed = dropout_seed; p.drop_tag = (uint32_t)o.dropout_tag; }
          RUN_TRY(nl, run_conv(h, p, h->precision, st));
        }
        break;
      }
      case OP_POOL: {
        const Buf& bi = h->bufs[o.in.buf];
        nl += launch_maxpool(B, bi.H, bi.W, o.in.C, o.k, X.ptr(o.in), X.cs(o.in), X.ptr(o.out), X.cs(o.out), st);
        break;
      }
      case OP_UPADD: {
        const Buf& bo = h->bufs[o.out.buf];
        nl += launch_upadd(B, bo.H, bo.W, o.out.C, X.ptr(o.in), X.cs(o.in), X.ptr(o.res), X.cs(o.res), X.ptr(o.out), X.cs(o.out), st);
        break;
      }
      case OP_MASKCOPY:
        nl += launch_copy_view(X.npix(o.in), o.in.C, X.ptr(o.in), X.cs(o.in), X.ptr(o.out), X.cs(o.out), 0, tiny, st);
        break;
    }
    rc = lanes.leave((int)oi);
    if (rc) return rc;
  }
  rc = lanes.join();
  if (rc) return rc;
  h->launches += nl;
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

struct LaneCtx;
int reduce_range(dr_handle* h, int64_t lo, int64_t hi, cudaStream_t st, LaneCtx* lanes);   // NCCL all-reduce of grads[lo,hi) (defined with the communicator below)

int apply_fills(dr_handle* h, Exec& X, const GradWrite& g, cudaStream_t st) {
  int nl = 0;
  for (int i = 0; i < g.nfill; ++i) nl += launch_fill_view(X.npix(g.fill[i]), g.fill[i].C, X.gptr(g.fill[i]), X.cs(g.fill[i]), 0.f, st);
  return nl;
}

int backward_impl(dr_handle* h, int B, const float* poses, const float* cfgs, const float* coms, float* loss_out, cudaStream_t st0,
                  cudaEvent_t ev_after_loss = nullptr) {
  if (!h->grads) return fail(h, DR_ERR_STATE, "grads buffer not bound");
  cudaStream_t st = st0;
  Exec X{h, B, st};
  int nl = 0;
  const int OUT = h->cfg.out_hw, J = h->cfg.num_jnt, S = h->cfg.num_stack;
  CUDA_TRY(h, cudaMemsetAsync(h->sums_bw, 0, h->n_sums * sizeof(double), st));
  CUDA_TRY(h, cudaMemsetAsync(h->loss_acc, 0, 4 * sizeof(double), st));
  // loss + dL/d(outputs)
  LossArgs la; memset(&la, 0, sizeof(la));
  la.B = B; la.hw = OUT; la.J = J; la.S = S;
  la.tiny = X.ptr(X.whole(h->buf_tiny)); la.poses = poses; la.cfgs = cfgs; la.coms = coms;
  for (int s = 0; s < S; ++s) {
    la.hm[s] = X.ptr(h->v_hm[s]); la.hm3[s] = X.ptr(h->v_hm3[s]); la.um[s] = X.ptr(h->v_um[s]); la.cs[s] = X.cs(h->v_hm[s]);
    la.ghm[s] = X.gptr(h->v_hm[s]); la.ghm3[s] = X.gptr(h->v_hm3[s]); la.gum[s] = X.gptr(h->v_um[s]); la.gcs[s] = X.cs(h->v_hm[s]);
  }
  la.loss_acc = h->loss_acc;
  nl += launch_loss(la, st);
  nl += launch_wd(h->n_params, h->params, h->wdmask, h->grads, h->loss_acc + 3, st);
  if (loss_out) nl += launch_finish_loss(h->loss_acc, loss_out, st);
  if (ev_after_loss) CUDA_TRY(h, cudaEventRecord(ev_after_loss, st));   // pipeline: the caller's inputs (crops, poses, cfgs, coms) are not read after this point

  int lane_convs[kLanes] = {};
  int wgrad_count = 0;
  LaneCtx lanes{h, st0, h->lanes_on, &h->plan_bwd, &h->ev_bwd, {}};
  { int rc = lanes.begin(); if (rc) return rc; }
  const bool overlap = h->overlap_armed && h->nccl_comm && h->comm_world > 1;
  size_t next_bucket = 0;
  h->overlap_armed = false;
  for (int oi = (int)h->ops.size() - 1; oi >= 0; --oi) {
    const Op& o = h->ops[oi];
    if (overlap) {           // every bucket whose last contributing op (in this reverse walk) lies behind us is final: reduce it now
      while (next_bucket < h->buckets.size() && h->buckets[next_bucket].first_op > oi) {
        int rc = reduce_range(h, h->buckets[next_bucket].lo, h->buckets[next_bucket].hi, st0, &lanes);
        if (rc) return rc;
        ++next_bucket;
      }
    }
    { int rc = lanes.enter(oi, &st); if (rc) return rc; }          // `st` = this op's lane stream from here on
    switch (o.kind) {
      case OP_CONV: {
        const Layer& L = h->layers[o.layer];
        const View& gv = o.gsrc.buf >= 0 ? o.gsrc : o.out;          // gradient aliasing: see Op
        const float* dy = X.gptr(gv); const int dy_cs = X.cs(gv);
        const size_t np = X.npix(o.out);
        if (o.res.buf >= 0 && !o.res_grad_fused) {
          nl += apply_fills(h, X, o.gw_res, st);
          nl += launch_copy_view(np, o.res.C, dy, dy_cs, X.gptr(o.res), X.cs(o.res), o.gw_res.acc, nullptr, st);
        }
        const int ln = lanes.on ? o.lane : 0;                         // d(raw) scratch slots are per lane (each lane cycles through three)
        const int slot = ln * dr_handle::kSlotsPerLane + lane_convs[ln] % dr_handle::kSlotsPerLane; ++lane_convs[ln];
        float* dz = h->scratch + (size_t)slot * h->scratch_per_crop * h->cap_B; const int dz_cs = (L.cout + 3) / 4 * 4;
        if (h->side_stream) CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_wdone[slot], 0));   // the wgrad that last read this slot is done
        if (L.brn) {
          View rv = X.whole(o.raw);
          double* sums = h->sums_bw + L.sum_off;
          // small layers: reduce + apply by one thread-block cluster in one launch; otherwise two grid-wide passes
          if (launch_brn_bwd_small(np, L.cout, dy, dy_cs, X.ptr(rv), X.cs(rv), h->aff + L.aff_off, h->bstat + L.bstat_off, h->params + L.p_off, L.relu,
                                   dz, dz_cs, h->grads + L.p_off, st)) {
            nl += 1;
          } else {
            nl += launch_brn_bwd_reduce(np, L.cout, dy, dy_cs, X.ptr(rv), X.cs(rv), h->aff + L.aff_off, h->bstat + L.bstat_off, L.relu, sums, st);
            nl += launch_brn_bwd_apply(np, L.cout, dy, dy_cs, X.ptr(rv), X.cs(rv), h->aff + L.aff_off, h->bstat + L.bstat_off,
                                       h->params + L.p_off, L.relu, sums, dz, dz_cs, h->grads + L.p_off, st);
          }
        } else {
          nl += launch_bias_bwd(np, L.cout, dy, dy_cs, X.ptr(o.out), X.cs(o.out), L.relu, o.dropout_tag >= 0, dz, dz_cs, h->grads + L.p_off, st);
        }
        WgradProblem wp; memset(&wp, 0, sizeof(wp));
        wp.x = X.ptr(o.in); wp.x_cs = X.cs(o.in); wp.dy = dz; wp.dy_cs = dz_cs;
        wp.B = B; wp.H = L.in_hw; wp.W = L.in_hw; wp.Cin = L.cin; wp.Ho = L.out_hw; wp.Wo = L.out_hw; wp.Cout = L.cout;
        wp.k = L.k; wp.stride = L.stride; wp.pad_t = wp.pad_l = same_pad_before(L.in_hw, L.k, L.stride);
        wp.dw = h->grads + L.w_off;
        if (h->side_stream) {
          cudaStream_t ws = (h->wgrad_streams > 1 && (wgrad_count++ & 1)) ? h->wgrad_stream2 : h->wgrad_stream;
          CUDA_TRY(h, cudaEventRecord(h->ev_ready[slot], st));
          CUDA_TRY(h, cudaStreamWaitEvent(ws, h->ev_ready[slot], 0));
          RUN_TRY(nl, run_wgrad(h, wp, h->precision, ws));
          CUDA_TRY(h, cudaEventRecord(h->ev_wdone[slot], ws));
        } else {
          RUN_TRY(nl, run_wgrad(h, wp, h->precision, st));
        }
        if (o.need_dgrad) {
          nl += apply_fills(h, X, o.gw_in, st);
          ConvProblem p; memset(&p, 0, sizeof(p));
          p.x = dz; p.x_cs = dz_cs; p.B = B; p.H = L.out_hw; p.W = L.out_hw; p.Cin = L.cout;
          p.Ho = L.in_hw; p.Wo = L.in_hw; p.Cout = L.cin; p.k = L.k; p.stride = 1;
          p.pad_t = p.pad_l = L.k - 1 - same_pad_before(L.in_hw, L.k, L.stride);
          set_dgrad_weights(h, L, h->precision, p);
          p.y = X.gptr(o.in); p.y_cs = X.cs(o.in); p.accumulate = o.gw_in.acc;
          if (o.dres.buf >= 0) { p.res = X.gptr(o.dres); p.res_cs = X.cs(o.dres); }      // identity skip: + d(block output)
          RUN_TRY(nl, run_conv(h, p, h->precision, st));
        }
        break;
      }
      case OP_POOL: {
        const Buf& bi = h->bufs[o.in.buf];
        nl += apply_fills(h, X, o.gw_in, st);
        nl += launch_maxpool_bwd(B, bi.H, bi.W, o.in.C, o.k, X.ptr(o.in), X.cs(o.in), X.gptr(o.out), X.cs(o.out),
                                 X.gptr(o.in), X.cs(o.in), o.gw_in.acc, st);
        break;
      }
      case OP_UPADD: {
        const Buf& bo = h->bufs[o.out.buf];
        if (!o.in_grad_fused) {
          nl += apply_fills(h, X, o.gw_in, st);
          nl += launch_copy_view(X.npix(o.out), o.out.C, X.gptr(o.out), X.cs(o.out), X.gptr(o.in), X.cs(o.in), o.gw_in.acc, nullptr, st);
        }
        nl += apply_fills(h, X, o.gw_res, st);
        nl += launch_upadd_bwd_lo(B, bo.H, bo.W, o.out.C, X.gptr(o.out), X.cs(o.out), X.gptr(o.res), X.cs(o.res), o.gw_res.acc, st);
        break;
      }
      case OP_MASKCOPY:
        nl += apply_fills(h, X, o.gw_in, st);
        nl += launch_copy_view(X.npix(o.in), o.in.C, X.gptr(o.out), X.cs(o.out), X.gptr(o.in), X.cs(o.in), o.gw_in.acc,
                               X.ptr(X.whole(h->buf_tiny)), st);
        break;
    }
    { int rc = lanes.leave(oi); if (rc) return rc; }
  }
  { int rc = lanes.join(); if (rc) return rc; }
  st = st0;
  if (overlap) {
    while (next_bucket < h->buckets.size()) {
      int rc = reduce_range(h, h->buckets[next_bucket].lo, h->buckets[next_bucket].hi, st0, nullptr);
      if (rc) return rc;
      ++next_bucket;
    }
    CUDA_TRY(h, cudaEventRecord(h->ev_comm_done, h->comm_stream));
    h->reduced_in_backward = true;
  }
  if (h->side_stream) {          // the optimiser step / next micro-batch on `st` must see every filter gradient
    CUDA_TRY(h, cudaEventRecord(h->ev_join, h->wgrad_stream));
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join, 0));
    CUDA_TRY(h, cudaEventRecord(h->ev_join2, h->wgrad_stream2));
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join2, 0));
  }
  h->launches += nl;
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

// ---------------------------------------------------------------------------------------------
// NCCL (resolved at run time: the library neither links libnccl nor needs it unless dr_comm_init is called; inside a PyTorch
// process the already loaded libnccl.so.2 is reused)
// ---------------------------------------------------------------------------------------------
struct DrNcclId { char internal[128]; };                   // == ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES 128)
struct NcclApi {
  typedef int (*GetUniqueId_t)(void*);
  typedef int (*CommInitRank_t)(void**, int, DrNcclId, int);
  typedef int (*AllReduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  typedef int (*CommDestroy_t)(void*);
  typedef const char* (*GetErrorString_t)(int);
  void* lib = nullptr;
  int (*get_unique_id)(void*) = nullptr;
  void* comm_init_rank = nullptr;
  AllReduce_t all_reduce = nullptr;
  CommDestroy_t comm_destroy = nullptr;
  GetErrorString_t error_string = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!a.lib) a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) {
      a.get_unique_id = reinterpret_cast<int (*)(void*)>(dlsym(a.lib, "ncclGetUniqueId"));
      a.comm_init_rank = dlsym(a.lib, "ncclCommInitRank");
      a.all_reduce = reinterpret_cast<NcclApi::AllReduce_t>(dlsym(a.lib, "ncclAllReduce"));
      a.comm_destroy = reinterpret_cast<NcclApi::CommDestroy_t>(dlsym(a.lib, "ncclCommDestroy"));
      a.error_string = reinterpret_cast<NcclApi::GetErrorString_t>(dlsym(a.lib, "ncclGetErrorString"));
      a.ok = a.get_unique_id && a.comm_init_rank && a.all_reduce && a.comm_destroy;
    }
  }
  return a;
}
int nccl_fail(dr_handle* h, const char* what, int code) {
  NcclApi& a = nccl_api();
  h->err = std::string(what) + ": " + (a.error_string ? a.error_string(code) : "NCCL error") + " (" + std::to_string(code) + ")";
  return DR_ERR_CUDA;
}

// gradient buckets for the overlapped all-reduce: ~equal parameter counts, cut at layer boundaries, last layers first (their gradients
// are final first in the reverse walk).  A bucket is final once the walk has finished the EARLIEST (forward order) conv op among its layers.
void plan_buckets(dr_handle* h, int nbuckets) {
  h->buckets.clear();
  const int nl = (int)h->layers.size();
  std::vector<int> op_of_layer(nl, 0);
  for (int oi = 0; oi < (int)h->ops.size(); ++oi)
    if (h->ops[oi].kind == OP_CONV) op_of_layer[h->ops[oi].layer] = oi;
  const int64_t per = ((int64_t)h->n_params + nbuckets - 1) / nbuckets;
  int hi_layer = nl;                                         // exclusive
  while (hi_layer > 0) {
    const int64_t hi_off = hi_layer == nl ? (int64_t)h->n_params : h->layers[hi_layer].w_off;
    int lo_layer = hi_layer - 1;
    while (lo_layer > 0 && hi_off - h->layers[lo_layer].w_off < per) --lo_layer;
    if ((int)h->buckets.size() == nbuckets - 1) lo_layer = 0;  // the last bucket takes everything that is left
    int first_op = 1 << 30;
    for (int l = lo_layer; l < hi_layer; ++l) if (op_of_layer[l] < first_op) first_op = op_of_layer[l];
    h->buckets.push_back(dr_handle::Bucket{h->layers[lo_layer].w_off, hi_off, first_op});
    hi_layer = lo_layer;
  }
}

// all-reduce(sum) of grads[lo,hi) on the communication stream, after everything enqueued so far on `st` and on the wgrad stream
int reduce_range(dr_handle* h, int64_t lo, int64_t hi, cudaStream_t st, LaneCtx* lanes) {
  NcclApi& a = nccl_api();
  CUDA_TRY(h, cudaEventRecord(h->ev_bucket_main, st));
  CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_bucket_main, 0));
  if (lanes && lanes->on)                                   // bias / BRN parameter gradients are written on the lane streams
    for (int l = 1; l < kLanes; ++l)
      if (lanes->started[l]) {
        CUDA_TRY(h, cudaEventRecord(h->ev_lane_done[l], h->lane_stream[l]));
        CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_lane_done[l], 0));
      }
  if (h->side_stream && h->wgrad_stream) {
    CUDA_TRY(h, cudaEventRecord(h->ev_bucket_side, h->wgrad_stream));
    CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_bucket_side, 0));
    CUDA_TRY(h, cudaEventRecord(h->ev_bucket_side2, h->wgrad_stream2));
    CUDA_TRY(h, cudaStreamWaitEvent(h->comm_stream, h->ev_bucket_side2, 0));
  }
  const int rc = a.all_reduce(h->grads + lo, h->grads + lo, (size_t)(hi - lo), /*ncclFloat32*/ 7, /*ncclSum*/ 0, h->nccl_comm, h->comm_stream);
  if (rc != 0) return nccl_fail(h, "ncclAllReduce", rc);
  ++h->allreduce_calls;
  return DR_OK;
}

// device-side tables are created lazily (first dr_bind) so that dr_create / dr_get_layer work on a
// machine without a GPU (CPU-only CI checks the layer table against the oracle).
int init_device(dr_handle* h) {
  if (h->ltab) return DR_OK;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  std::vector<LayerDev> tab(h->layers.size());
  std::vector<float> wdm(h->n_params, 0.f);
  for (size_t i = 0; i < h->layers.size(); ++i) {
    const Layer& L = h->layers[i];
    tab[i] = LayerDev{L.cout, L.brn, L.k * L.k, L.cin, L.w_off, L.p_off, L.s_off, L.aff_off, L.wk_off, L.wa_off};
    if (L.wd > 0) for (int64_t j = 0; j < (int64_t)L.k * L.k * L.cin * L.cout; ++j) wdm[L.w_off + j] = L.wd;
  }
  CUDA_TRY(h, cudaMalloc(&h->ltab, tab.size() * sizeof(LayerDev)));
  CUDA_TRY(h, cudaMalloc(&h->wdmask, wdm.size() * sizeof(float)));
  CUDA_TRY(h, cudaMalloc(&h->aff, h->n_aff * sizeof(float)));
  CUDA_TRY(h, cudaMalloc(&h->bstat, h->n_bstat * sizeof(float)));
  CUDA_TRY(h, cudaMalloc(&h->sums, h->n_sums * sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&h->sums_bw, h->n_sums * sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&h->loss_acc, 4 * sizeof(double)));
  CUDA_TRY(h, cudaMalloc(&h->counters, h->layers.size() * sizeof(unsigned int)));
  CUDA_TRY(h, cudaMalloc(&h->counters_bw, h->layers.size() * sizeof(unsigned int)));
  CUDA_TRY(h, cudaMalloc(&h->clamp_dev, sizeof(int32_t)));
  CUDA_TRY(h, cudaMemcpy(h->ltab, tab.data(), tab.size() * sizeof(LayerDev), cudaMemcpyHostToDevice));
  CUDA_TRY(h, cudaMemcpy(h->wdmask, wdm.data(), wdm.size() * sizeof(float), cudaMemcpyHostToDevice));
  h->ws_bytes += tab.size() * sizeof(LayerDev) + (wdm.size() + h->n_aff + h->n_bstat) * sizeof(float) + 2 * h->n_sums * sizeof(double);
  return DR_OK;
}


// ---------------------------------------------------------------------------------------------
// micro-batch pipeline (dr_handle::twin)
// ---------------------------------------------------------------------------------------------
int pipe_fail(dr_handle* h, dr_handle* e, int rc) { if (e != h) h->err = e->err; return rc; }

// The stream operations of the pipeline as DATA: pipe_loss_backward / pipe_join execute these lists with CUDA calls, dr_debug_pipeline_plan
// exports the same lists, and tests/test_pipeline_plan.py replays them on a model of CUDA's stream / event semantics (forward passes in
// order, one backward pass at a time, loss and inputs ordered on the caller's stream, forward(i+1) really free to overlap backward(i)).
enum { PE_IN = 0, PE_FWD0 = 1, PE_BWD0 = 3, PE_LOSS0 = 5, PE_COUNT = 7 };     // event ids: in, fwd[slot], bwd[slot], loss[slot]
enum { PS_CALLER = -1 };                                                     // stream ids: the caller's, or the slot's internal stream (0 / 1)
enum { PO_RECORD = 0, PO_WAIT = 1, PO_FORWARD = 2, PO_BACKWARD = 3 };        // PO_BACKWARD records `event` right after its loss kernels
struct PipeOp { int kind, stream, event; };
const int kPipeMaxOps = 12;

// slot of the next micro-batch.  Slot 0 (the handle itself) whenever the pass has to run there: the weight copies must be rebuilt, the pass
// all-reduces gradient buckets through the handle's communicator, or its launches are being timed one by one
int pipe_choose_slot(const dr_handle* h, bool* dirty) {
  *dirty = h->weights_dirty || h->prepped_precision < 0;
  if (*dirty || h->overlap_armed || h->trace_on || trace_env()) return 0;
  return h->pipe_next;
}

int pipe_plan_micro_batch(int k, bool dirty, PipeOp* o) {
  int n = 0;
  o[n++] = PipeOp{PO_RECORD, PS_CALLER, PE_IN};              // inputs / zeroed gradients / updated parameters of the caller's stream
  o[n++] = PipeOp{PO_WAIT, k, PE_IN};
  if (dirty) o[n++] = PipeOp{PO_WAIT, k, PE_BWD0 + 1};       // a pass still in flight in the second arena reads the weight copies about to be rebuilt
  o[n++] = PipeOp{PO_WAIT, k, PE_FWD0 + (k ^ 1)};            // BRN moving statistics: forward passes in micro-batch order
  o[n++] = PipeOp{PO_FORWARD, k, -1};
  o[n++] = PipeOp{PO_RECORD, k, PE_FWD0 + k};
  o[n++] = PipeOp{PO_WAIT, k, PE_BWD0 + (k ^ 1)};            // gradient buffer: one backward pass at a time
  o[n++] = PipeOp{PO_BACKWARD, k, PE_LOSS0 + k};
  o[n++] = PipeOp{PO_RECORD, k, PE_BWD0 + k};
  o[n++] = PipeOp{PO_WAIT, PS_CALLER, PE_LOSS0 + k};         // loss_out is valid and the inputs are free on the caller's stream
  return n;
}

// everything the pipeline still has in flight precedes whatever is enqueued on the caller's stream next (gradients complete, arenas idle)
int pipe_plan_join(PipeOp* o) {
  o[0] = PipeOp{PO_WAIT, PS_CALLER, PE_BWD0};
  o[1] = PipeOp{PO_WAIT, PS_CALLER, PE_BWD0 + 1};
  return 2;
}

int pipe_join(dr_handle* h, cudaStream_t st) {
  if (!h->twin || !h->pipe_pending) return DR_OK;
  PipeOp ops[kPipeMaxOps];
  const int n = pipe_plan_join(ops);
  for (int i = 0; i < n; ++i) CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_pipe[ops[i].event], 0));
  h->pipe_pending = false; h->pipe_next = 0;
  return DR_OK;
}

int pipe_init(dr_handle* h) {
  if (!h->twin || h->pipe_stream[0]) return DR_OK;
  for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaStreamCreateWithPriority(&h->pipe_stream[k], cudaStreamNonBlocking, h->chain_prio));
  for (int i = 0; i < PE_COUNT; ++i) CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_pipe[i], cudaEventDisableTiming));
  return DR_OK;
}

// the twin reads this handle's weight copies (built by this handle's forward pass whenever the parameters changed)
void pipe_share_weights(dr_handle* h) {
  dr_handle* t = h->twin;
  t->wk = h->wk; t->wa = h->wa; t->wk_hi = h->wk_hi; t->wk_lo = h->wk_lo; t->wa_hi = h->wa_hi; t->wa_lo = h->wa_lo;
  t->weights_dirty = false; t->prepped_precision = h->prepped_precision;
}

int pipe_loss_backward(dr_handle* h, int B, const float* dm_mm, const float* poses_mm, const float* cfgs, const float* coms, float* loss_out,
                       uint64_t dropout_seed, int update_state, cudaStream_t caller) {
  int rc = pipe_init(h);
  if (rc) return rc;
  bool dirty = false;
  const int k = pipe_choose_slot(h, &dirty);
  dr_handle* e = k ? h->twin : h;
  PipeOp ops[kPipeMaxOps];
  const int n = pipe_plan_micro_batch(k, dirty, ops);
  for (int i = 0; i < n; ++i) {
    const PipeOp& op = ops[i];
    cudaStream_t s = op.stream == PS_CALLER ? caller : h->pipe_stream[op.stream];
    switch (op.kind) {
      case PO_RECORD: CUDA_TRY(h, cudaEventRecord(h->ev_pipe[op.event], s)); break;
      case PO_WAIT: CUDA_TRY(h, cudaStreamWaitEvent(s, h->ev_pipe[op.event], 0)); break;
      case PO_FORWARD:
        if (k) { pipe_share_weights(h); ++h->pipe_twin_runs; }
        rc = forward_impl(e, B, dm_mm, coms, 1, update_state, dropout_seed, s);
        if (rc) return pipe_fail(h, e, rc);
        break;
      case PO_BACKWARD:
        rc = backward_impl(e, B, poses_mm, cfgs, coms, loss_out, s, h->ev_pipe[op.event]);
        if (rc) return pipe_fail(h, e, rc);
        break;
    }
  }
  h->pipe_next = k ^ 1; h->pipe_pending = true;
  return DR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int dr_version(void) { return DR_VERSION; }

int dr_create(dr_handle** out, const dr_config* cfg) {
  if (!out || !cfg) return DR_ERR_ARG;
  *out = nullptr;
  if (cfg->kernel_size != 3 || cfg->in_hw != 128 || cfg->out_hw != 32 || cfg->num_stack < 1 || cfg->num_stack > 4 ||
      cfg->num_fea < 8 || cfg->num_fea % 4 != 0 || cfg->num_jnt < 1 || cfg->num_jnt > 64 || cfg->max_batch < 1)
    return DR_ERR_ARG;
  dr_handle* h = new dr_handle();
  h->cfg = *cfg;
  h->precision = cfg->precision;
  // CTA-pair (cta_group::2) 3xTF32 kernel for the big layers: on by default; dr_config.reserved[1] < 0 or DENSEREG_TC_PAIR=0 turns it off
  { const char* env = getenv("DENSEREG_TC_PAIR"); h->tc_pair = cfg->reserved[1] >= 0 && !(env && env[0] == '0'); }
  { const char* env = getenv("DENSEREG_PREP_ONCE"); h->prep_once = !(env && env[0] == '0'); }
  { const char* env = getenv("DENSEREG_LANES"); h->lanes_on = !(env && env[0] == '0'); }
  { const char* e = getenv("DENSEREG_TC_CHUNK_EVAL"); if (e) h->chunk_eval = atoi(e) > 0 ? atoi(e) : 0; }
  { const char* e = getenv("DENSEREG_TC_CHUNK_MINKB"); if (e) h->chunk_min_kb = atoi(e) > 0 ? atoi(e) : 0; }
  { const char* e = getenv("DENSEREG_CHAIN_PRIO");
    if (e && e[0] == '1') { int lo = 0, hi = 0; if (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess) h->chain_prio = hi; else cudaGetLastError(); } }
  Builder b{h, 0, 0, 0};
  b.build();
  if (b.plan_overflow) { delete h; return DR_ERR_UNSUPPORTED; }
  if (getenv("DENSEREG_DUMP_PLAN")) {                   // debug: the lane plan (op, lane, waits, record) of both passes
    for (size_t i = 0; i < h->ops.size(); ++i) {
      const Op& o = h->ops[i];
      const char* kind = o.kind == OP_CONV ? "conv" : o.kind == OP_POOL ? "pool" : o.kind == OP_UPADD ? "upadd" : "maskcopy";
      fprintf(stderr, "PLAN %3zu %-8s %-24s lane %d | fwd wait", i, kind, o.kind == OP_CONV ? h->layers[o.layer].name : "", o.lane);
      for (int k = 0; k < h->plan_fwd[i].nwait; ++k) fprintf(stderr, " %d", h->plan_fwd[i].wait_op[k]);
      fprintf(stderr, " rec %d | bwd wait", h->plan_fwd[i].record);
      for (int k = 0; k < h->plan_bwd[i].nwait; ++k) fprintf(stderr, " %d", h->plan_bwd[i].wait_op[k]);
      fprintf(stderr, " rec %d\n", h->plan_bwd[i].record);
    }
  }   // a gradient view with > 4 unwritten channel gaps (never on um_v1)
  // micro-batch pipeline: a second handle on the same buffers (see dr_handle::twin).  dr_config.reserved[2] == 2 asks for it;
  // DENSEREG_PIPELINE=2 / =1 forces it on / off for A/B measurements
  int depth = cfg->reserved[2];
  { const char* e = getenv("DENSEREG_PIPELINE"); if (e && (e[0] == '1' || e[0] == '2')) depth = e[0] - '0'; }
  if (depth == 2 && cfg->reserved[2] != kTwinMark) {
    dr_config c2 = *cfg;
    c2.reserved[0] = 0; c2.reserved[2] = kTwinMark;
    if (dr_create(&h->twin, &c2) != DR_OK) { delete h; return DR_ERR_UNSUPPORTED; }
    h->twin->is_twin = true;
  }
  *out = h;
  return DR_OK;
}

int dr_destroy(dr_handle* h) {
  if (!h) return DR_ERR_ARG;
  if (h->twin) {
    for (int k = 0; k < 2; ++k) if (h->pipe_stream[k]) cudaStreamSynchronize(h->pipe_stream[k]);
    dr_destroy(h->twin); h->twin = nullptr;
    for (int k = 0; k < 2; ++k) if (h->pipe_stream[k]) cudaStreamDestroy(h->pipe_stream[k]);
    for (cudaEvent_t ev : h->ev_pipe) if (ev) cudaEventDestroy(ev);
  }
  if (h->is_twin) h->wk = h->wa = h->wk_hi = h->wk_lo = h->wa_hi = h->wa_lo = nullptr;     // owned by the first handle
  if (h->nccl_comm) { nccl_api().comm_destroy(h->nccl_comm); h->nccl_comm = nullptr; }
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  if (h->ev_bucket_main) cudaEventDestroy(h->ev_bucket_main);
  if (h->ev_bucket_side) cudaEventDestroy(h->ev_bucket_side);
  if (h->ev_bucket_side2) cudaEventDestroy(h->ev_bucket_side2);
  if (h->ev_comm_done) cudaEventDestroy(h->ev_comm_done);
  if (h->infer_graph.exec) cudaGraphExecDestroy(h->infer_graph.exec);
  if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
  if (h->wgrad_stream) cudaStreamDestroy(h->wgrad_stream);
  if (h->wgrad_stream2) cudaStreamDestroy(h->wgrad_stream2);
  if (h->ev_join2) cudaEventDestroy(h->ev_join2);
  for (int i = 0; i < dr_handle::kScratchSlots; ++i) { if (h->ev_ready[i]) cudaEventDestroy(h->ev_ready[i]); if (h->ev_wdone[i]) cudaEventDestroy(h->ev_wdone[i]); }
  for (int l = 1; l < kLanes; ++l) if (h->lane_stream[l]) cudaStreamDestroy(h->lane_stream[l]);
  for (int l = 0; l < kLanes; ++l) if (h->ev_lane_done[l]) cudaEventDestroy(h->ev_lane_done[l]);
  if (h->ev_pass_start) cudaEventDestroy(h->ev_pass_start);
  for (cudaEvent_t e : h->ev_fwd) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_bwd) if (e) cudaEventDestroy(e);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaFree(h->act); cudaFree(h->gact); cudaFree(h->rawa); cudaFree(h->scratch); cudaFree(h->aff); cudaFree(h->bstat);
  cudaFree(h->wk); cudaFree(h->wa); cudaFree(h->wk_hi); cudaFree(h->wk_lo); cudaFree(h->wa_hi); cudaFree(h->wa_lo); cudaFree(h->wdmask); cudaFree(h->sums); cudaFree(h->sums_bw); cudaFree(h->loss_acc);
  cudaFree(h->clamp_dev); cudaFree(h->crop_scratch); cudaFree(h->ltab); cudaFree(h->counters); cudaFree(h->counters_bw);
  delete h;
  return DR_OK;
}

const char* dr_last_error(const dr_handle* h) { return h ? h->err.c_str() : "null handle"; }
size_t dr_param_count(const dr_handle* h) { return h ? h->n_params : 0; }
size_t dr_state_count(const dr_handle* h) { return h ? h->n_state : 0; }
int dr_num_layers(const dr_handle* h) { return h ? (int)h->layers.size() : 0; }
int dr_debug_get_output(dr_handle* h, int layer, int B, float* dst, int grad, void* stream) {
  if (!h || !dst || layer < 0 || layer >= (int)h->layers.size() || B < 1 || B > h->cap_B) return DR_ERR_ARG;
  if (grad && !h->gact) return fail(h, DR_ERR_STATE, "no gradient workspace");
  { int rc = pipe_join(h, (cudaStream_t)stream); if (rc) return rc; }     // (pipeline: this handle's arena holds the last micro-batch that ran in slot 0)
  Exec X{h, B, (cudaStream_t)stream};
  for (const Op& o : h->ops) {
    if (o.kind == OP_CONV && o.layer == layer) {
      const View& gv = (grad && o.gsrc.buf >= 0) ? o.gsrc : o.out;
      h->launches += launch_gather_outputs(X.npix(o.out), o.out.C, grad ? X.gptr(gv) : X.ptr(o.out), X.cs(gv), dst, X.st);
      CUDA_TRY(h, cudaGetLastError());
      return DR_OK;
    }
  }
  return DR_ERR_ARG;
}

int dr_trace(dr_handle* h, int on) {
  if (!h) return DR_ERR_ARG;
  h->trace_on = on != 0;
  if (on) h->trace.clear();
  return DR_OK;
}
int dr_trace_count(const dr_handle* h) { return h ? (int)h->trace.size() : 0; }
int dr_trace_get(const dr_handle* h, int idx, dr_trace_rec* out) {
  if (!h || !out || idx < 0 || idx >= (int)h->trace.size()) return DR_ERR_ARG;
  *out = h->trace[idx];
  return DR_OK;
}

int dr_num_ops(const dr_handle* h) { return h ? (int)h->ops.size() : 0; }
int dr_debug_op(const dr_handle* h, int idx, dr_op_info* out) {
  if (!h || !out || idx < 0 || idx >= (int)h->ops.size()) return DR_ERR_ARG;
  const Op& o = h->ops[idx];
  memset(out, 0, sizeof(*out));
  out->kind = (int32_t)o.kind; out->lane = o.lane; out->layer = o.layer; out->need_dgrad = o.need_dgrad; out->raw_buf = o.raw;
  out->in_buf = o.in.buf; out->in_c0 = o.in.coff; out->in_c = o.in.C;
  out->out_buf = o.out.buf; out->out_c0 = o.out.coff; out->out_c = o.out.C;
  out->res_buf = o.res.buf; out->res_c0 = o.res.coff; out->res_c = o.res.C;
  out->gsrc_buf = o.gsrc.buf; out->gsrc_c0 = o.gsrc.coff; out->gsrc_c = o.gsrc.C;
  out->dres_buf = o.dres.buf; out->dres_c0 = o.dres.coff; out->dres_c = o.dres.C;
  out->res_grad_fused = o.res_grad_fused; out->in_grad_fused = o.in_grad_fused;
  const OpPlan* pl[2] = {&h->plan_fwd[idx], &h->plan_bwd[idx]};
  for (int k = 0; k < 2; ++k) {
    out->nwait[k] = pl[k]->nwait; out->record[k] = pl[k]->record;
    for (int i = 0; i < pl[k]->nwait && i < 3; ++i) out->wait_op[k][i] = pl[k]->wait_op[i];
  }
  return DR_OK;
}

int64_t dr_launch_count(const dr_handle* h) { return h ? h->launches + (h->twin ? h->twin->launches : 0) : 0; }
int64_t dr_tc_launch_count(const dr_handle* h) { return h ? h->tc_launches + (h->twin ? h->twin->tc_launches : 0) : 0; }
size_t dr_workspace_bytes(const dr_handle* h) { return h ? h->ws_bytes + (h->twin ? h->twin->ws_bytes : 0) : 0; }

int dr_get_layer(const dr_handle* h, int idx, dr_layer_info* out) {
  if (!h || !out || idx < 0 || idx >= (int)h->layers.size()) return DR_ERR_ARG;
  const Layer& L = h->layers[idx];
  memset(out, 0, sizeof(*out));
  memcpy(out->name, L.name, sizeof(out->name));
  out->k = L.k; out->stride = L.stride; out->cin = L.cin; out->cout = L.cout; out->brn = L.brn; out->relu = L.relu;
  out->wd = L.wd; out->w_off = L.w_off; out->p_off = L.p_off; out->s_off = L.brn ? L.s_off : -1;
  out->in_hw = L.in_hw; out->out_hw = L.out_hw;
  return DR_OK;
}

int dr_bind(dr_handle* h, float* params, float* state, float* grads, float* adam_m, float* adam_v) {
  if (!h || !params || !state) return DR_ERR_ARG;
  h->params = params; h->state = state; h->grads = grads; h->adam_m = adam_m; h->adam_v = adam_v;
  h->weights_dirty = true;
  int rc = init_device(h);
  if (rc || !h->twin) return rc;
  rc = dr_bind(h->twin, params, state, grads, adam_m, adam_v);      // same buffers: the two arenas of the micro-batch pipeline
  if (rc) h->err = h->twin->err;
  return rc;
}

int dr_params_changed(dr_handle* h) {
  if (!h) return DR_ERR_ARG;
  h->weights_dirty = true;
  return DR_OK;
}

int dr_init_params(dr_handle* h, uint64_t seed, float stddev, void* stream) {
  if (!h) return DR_ERR_ARG;
  if (!h->params || !h->state) return fail(h, DR_ERR_STATE, "dr_bind() not called");
  cudaStream_t st = (cudaStream_t)stream;
  { int rc = pipe_join(h, st); if (rc) return rc; }
  h->launches += launch_init_trunc_normal(h->n_params, h->params, stddev, seed, st);
  init_state_kernel<<<(unsigned)h->layers.size(), 128, 0, st>>>(h->ltab, h->params, h->state); ++h->launches;
  h->weights_dirty = true;
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

int dr_norm_dm(dr_handle* h, int B, int hw, const float* dm_mm, const float* coms, float* out, void* stream) {
  if (!h || !dm_mm || !coms || !out || B < 1 || hw < 1) return DR_ERR_ARG;
  h->launches += launch_norm_dm(B, hw, dm_mm, coms, out, (cudaStream_t)stream);
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

static int copy_outputs(dr_handle* h, int B, float* const* hm, float* const* hm3, float* const* um, cudaStream_t st) {
  Exec X{h, B, st};
  for (int s = 0; s < h->cfg.num_stack; ++s) {
    if (hm && hm[s]) h->launches += launch_gather_outputs(X.npix(h->v_hm[s]), h->v_hm[s].C, X.ptr(h->v_hm[s]), X.cs(h->v_hm[s]), hm[s], st);
    if (hm3 && hm3[s]) h->launches += launch_gather_outputs(X.npix(h->v_hm3[s]), h->v_hm3[s].C, X.ptr(h->v_hm3[s]), X.cs(h->v_hm3[s]), hm3[s], st);
    if (um && um[s]) h->launches += launch_gather_outputs(X.npix(h->v_um[s]), h->v_um[s].C, X.ptr(h->v_um[s]), X.cs(h->v_um[s]), um[s], st);
  }
  return DR_OK;
}

int dr_forward(dr_handle* h, int B, const float* dm_mm, const float* coms,
               float* const* hm, float* const* hm3, float* const* um,
               int is_training, int update_state, uint64_t dropout_seed, void* stream) {
  if (!h || !dm_mm || !coms) return DR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = pipe_join(h, st);
  if (rc) return rc;
  rc = forward_impl(h, B, dm_mm, coms, is_training, update_state, dropout_seed, st);
  if (rc) return rc;
  copy_outputs(h, B, hm, hm3, um, st);
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

int dr_vote(dr_handle* h, int B, int H, int W, int J,
            const float* hm, const float* hm3, const float* um, const float* dm_norm,
            const float* cfgs, const float* coms,
            float* xyz_mm, int32_t* top5_idx, int32_t* clamp_count, void* stream) {
  if (!h || !hm || !hm3 || !um || !dm_norm || !cfgs || !coms || !xyz_mm) return DR_ERR_ARG;
  if (B < 1 || H < 1 || W < 1 || J < 1 || J > 64 || H * W < 5) return DR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (clamp_count) CUDA_TRY(h, cudaMemsetAsync(clamp_count, 0, sizeof(int32_t), st));
  h->launches += launch_vote(B, H, W, J, hm, J, hm3, J, um, 3 * J, dm_norm, cfgs, coms, xyz_mm, top5_idx, clamp_count, st);
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

static int infer_enqueue(dr_handle* h, int B, const float* dm_mm, const float* cfgs, const float* coms, float* xyz_mm, int32_t* top5_idx,
                         cudaStream_t st) {
  int rc = forward_impl(h, B, dm_mm, coms, 0, 0, 0, st);
  if (rc) return rc;
  Exec X{h, B, st};
  const int s = h->cfg.num_stack - 1, J = h->cfg.num_jnt, OUT = h->cfg.out_hw;
  h->launches += launch_vote(B, OUT, OUT, J, X.ptr(h->v_hm[s]), X.cs(h->v_hm[s]), X.ptr(h->v_hm3[s]), X.cs(h->v_hm3[s]),
                             X.ptr(h->v_um[s]), X.cs(h->v_um[s]), X.ptr(X.whole(h->buf_tiny)), cfgs, coms, xyz_mm, top5_idx,
                             h->clamp_dev, st);
  return DR_OK;
}

int dr_infer(dr_handle* h, int B, const float* dm_mm, const float* cfgs, const float* coms,
             float* xyz_mm, int32_t* top5_idx, void* stream) {
  if (!h || !dm_mm || !cfgs || !coms || !xyz_mm) return DR_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  { int rc = pipe_join(h, st); if (rc) return rc; }
  if (h->cfg.reserved[0] == 0) {                          // eager
    int rc = infer_enqueue(h, B, dm_mm, cfgs, coms, xyz_mm, top5_idx, st);
    if (rc) return rc;
    CUDA_TRY(h, cudaGetLastError());
    return DR_OK;
  }
  if (h->precision != DR_PREC_FP32) { int rc = ensure_prepped(h, h->precision, st); if (rc) return rc; }   // never inside the captured graph
  dr_handle::InferGraph& g = h->infer_graph;
  const bool same = g.B == B && g.dm == dm_mm && g.cfg == cfgs && g.com == coms && g.xyz == xyz_mm && g.top5 == top5_idx;
  if (!same) {                                            // new key: drop the old graph, run eagerly once (allocations, attributes)
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    g.B = B; g.dm = dm_mm; g.cfg = cfgs; g.com = coms; g.xyz = xyz_mm; g.top5 = top5_idx; g.warm = 0;
  }
  if (!g.exec && g.warm >= 1) {                           // second call with the same key: capture
    if (!h->capture_stream) CUDA_TRY(h, cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamSynchronize(st));               // the eager warm-up on `st` is done before the internal stream touches the arena
    CUDA_TRY(h, cudaStreamBeginCapture(h->capture_stream, cudaStreamCaptureModeThreadLocal));
    static int pdl_graph = -1;                            // programmatic edges are kept inside the captured graph (B=1: 1.68 -> 1.56 ms); DENSEREG_PDL_GRAPH=0: plain nodes
    if (pdl_graph < 0) { const char* e = getenv("DENSEREG_PDL_GRAPH"); pdl_graph = (e && e[0] == '0') ? 0 : 1; }
    if (!pdl_graph) dr_pdl_suspend(1);                    // plain kernel nodes inside the graph
    int rc = infer_enqueue(h, B, dm_mm, cfgs, coms, xyz_mm, top5_idx, h->capture_stream);
    dr_pdl_suspend(0);
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(h->capture_stream, &graph);
    if (rc || ce != cudaSuccess || !graph) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      g.warm = -1000000;                                  // capture failed: stay eager for this key
    } else {
      ce = cudaGraphInstantiate(&g.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; g.warm = -1000000; }
    }
  }
  if (g.exec) {
    CUDA_TRY(h, cudaGraphLaunch(g.exec, st));
    return DR_OK;
  }
  ++g.warm;
  int rc = infer_enqueue(h, B, dm_mm, cfgs, coms, xyz_mm, top5_idx, st);
  if (rc) return rc;
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

int dr_loss_backward(dr_handle* h, int B, const float* dm_mm, const float* poses_mm,
                     const float* cfgs, const float* coms, float* loss_out,
                     uint64_t dropout_seed, int update_state, void* stream) {
  if (!h || !dm_mm || !poses_mm || !cfgs || !coms) return DR_ERR_ARG;
  if (!h->grads) return fail(h, DR_ERR_STATE, "grads buffer not bound");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->twin) return pipe_loss_backward(h, B, dm_mm, poses_mm, cfgs, coms, loss_out, dropout_seed, update_state, st);
  int rc = forward_impl(h, B, dm_mm, coms, 1, update_state, dropout_seed, st);
  if (rc) return rc;
  return backward_impl(h, B, poses_mm, cfgs, coms, loss_out, st);
}

int dr_pipeline_join(dr_handle* h, void* stream) {
  if (!h) return DR_ERR_ARG;
  return pipe_join(h, (cudaStream_t)stream);
}

int dr_pipeline_depth(const dr_handle* h) { return h ? (h->twin ? 2 : 1) : 0; }

int dr_debug_pipeline_plan(dr_handle* h, int what, dr_pipe_op* out, int cap) {
  if (!h || !out || cap < kPipeMaxOps || what < 0 || what > 3) return DR_ERR_ARG;
  if (!h->twin) return fail(h, DR_ERR_STATE, "dr_debug_pipeline_plan: the handle has no micro-batch pipeline (dr_config.reserved[2] != 2)");
  if (h->params) return fail(h, DR_ERR_STATE, "dr_debug_pipeline_plan is a dry run: only on a handle that was never bound");
  PipeOp ops[kPipeMaxOps];
  int n = 0;
  switch (what) {
    case 0: {                                    // the next dr_loss_backward, with the bookkeeping the real call (and the passes it runs) would do
      bool dirty = false;
      const int k = pipe_choose_slot(h, &dirty);
      n = pipe_plan_micro_batch(k, dirty, ops);
      h->weights_dirty = false; h->prepped_precision = h->precision;       // ensure_prepped() in the forward pass
      h->overlap_armed = false;                                            // consumed by the backward pass
      h->pipe_next = k ^ 1; h->pipe_pending = true;
      break;
    }
    case 1: case 2:                              // dr_pipeline_join / dr_zero_grads (1), dr_optimizer_step (2): join, (2) parameters changed
      if (h->pipe_pending) n = pipe_plan_join(ops);
      h->pipe_pending = false; h->pipe_next = 0;
      if (what == 2) h->weights_dirty = true;
      break;
    case 3: h->overlap_armed = true; break;      // dr_comm_overlap_next_backward with a communicator
  }
  for (int i = 0; i < n; ++i) { out[i].kind = ops[i].kind; out[i].stream = ops[i].stream; out[i].event = ops[i].event; }
  return n;
}

int dr_comm_unique_id(void* out128) {
  if (!out128) return DR_ERR_ARG;
  NcclApi& a = nccl_api();
  if (!a.ok) return DR_ERR_UNSUPPORTED;
  return a.get_unique_id(out128) == 0 ? DR_OK : DR_ERR_CUDA;
}

int dr_comm_init(dr_handle* h, int rank, int world, const void* nccl_unique_id128) {
  if (!h || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_unique_id128)) return DR_ERR_ARG;
  if (h->nccl_comm) return fail(h, DR_ERR_STATE, "dr_comm_init called twice");
  h->comm_rank = rank; h->comm_world = world;
  if (world == 1) return DR_OK;
  NcclApi& a = nccl_api();
  if (!a.ok) return fail(h, DR_ERR_UNSUPPORTED, "libnccl.so.2 not found (dlopen)");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  DrNcclId id; memcpy(&id, nccl_unique_id128, sizeof(id));
  void* comm = nullptr;
  const int rc = reinterpret_cast<NcclApi::CommInitRank_t>(a.comm_init_rank)(&comm, world, id, rank);
  if (rc != 0) return nccl_fail(h, "ncclCommInitRank", rc);
  h->nccl_comm = comm;
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_bucket_main, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_bucket_side, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_bucket_side2, cudaEventDisableTiming));
  CUDA_TRY(h, cudaEventCreateWithFlags(&h->ev_comm_done, cudaEventDisableTiming));
  int nb = 4; { const char* e = getenv("DENSEREG_COMM_BUCKETS"); if (e && atoi(e) > 0) nb = atoi(e); }
  plan_buckets(h, nb);
  return DR_OK;
}

int dr_comm_overlap_next_backward(dr_handle* h) {
  if (!h) return DR_ERR_ARG;
  h->overlap_armed = h->nccl_comm != nullptr && h->comm_world > 1;
  return DR_OK;
}

int64_t dr_comm_allreduce_count(const dr_handle* h) { return h ? h->allreduce_calls : 0; }

int dr_zero_grads(dr_handle* h, void* stream) {
  if (!h) return DR_ERR_ARG;
  if (!h->grads) return fail(h, DR_ERR_STATE, "grads buffer not bound");
  { int rc = pipe_join(h, (cudaStream_t)stream); if (rc) return rc; }
  CUDA_TRY(h, cudaMemsetAsync(h->grads, 0, h->n_params * sizeof(float), (cudaStream_t)stream));
  return DR_OK;
}

int dr_optimizer_step(dr_handle* h, int accum_steps, int world, float lr, int64_t step, void* stream) {
  if (!h || accum_steps < 1 || world < 1 || step < 1) return DR_ERR_ARG;
  if (!h->grads || !h->adam_m || !h->adam_v || !h->params) return fail(h, DR_ERR_STATE, "grads / adam buffers not bound");
  cudaStream_t ost = (cudaStream_t)stream;
  { int rc = pipe_join(h, ost); if (rc) return rc; }
  if (h->nccl_comm && h->comm_world > 1) {
    if (world != h->comm_world) return fail(h, DR_ERR_ARG, "dr_optimizer_step: world differs from the communicator's size");
    if (!h->reduced_in_backward) {                          // not overlapped with the last backward: one all-reduce of the whole buffer now
      int rc = reduce_range(h, 0, (int64_t)h->n_params, ost, nullptr);
      if (rc) return rc;
      CUDA_TRY(h, cudaEventRecord(h->ev_comm_done, h->comm_stream));
    }
    CUDA_TRY(h, cudaStreamWaitEvent(ost, h->ev_comm_done, 0));
    h->reduced_in_backward = false;
  }
  // hourglass_um_crop_tiny.py:77,439; TF ApplyAdam arithmetic in fp32: beta powers by repeated fp32 multiplication (TF keeps them as
  // fp32 variables), alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)
  const float b1 = 0.5f, b2 = 0.999f;
  if (step == h->pow_step + 1) { h->b1p *= b1; h->b2p *= b2; }                 // the usual case: one multiplication per step, like TF's update of its power variables
  else { h->b1p = 1.0f; h->b2p = 1.0f; for (int64_t t = 0; t < step; ++t) { h->b1p *= b1; h->b2p *= b2; if (h->b1p == 0.0f && h->b2p == 0.0f) break; } }
  h->pow_step = step;
  const float b1p = h->b1p, b2p = h->b2p;
  const float alpha = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
  h->launches += launch_adam(h->n_params, h->params, h->grads, h->adam_m, h->adam_v, (float)(accum_steps * world), 0.2f,
                             alpha, 1.0f - b1, 1.0f - b2, 1e-8f, ost);
  h->weights_dirty = true;
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

static int crop_common(dr_handle* h, int B, int in_h, int in_w, const float* frames, const float* poses, int J, const float* bbx,
                       const float* cfg_host6, int out_hw, float pad, int icvl, float* dm_out, float* cfg_out, float* com_out, void* stream) {
  if (!h || !frames || !cfg_host6 || !dm_out || !cfg_out || !com_out || B < 1 || in_h < 2 || in_w < 2 || out_hw < 2 || out_hw > 1024)
    return DR_ERR_ARG;
  const size_t need = crop_scratch_bytes(B);
  if (need > h->crop_scratch_cap) {
    cudaFree(h->crop_scratch); h->crop_scratch = nullptr; h->crop_scratch_cap = 0;
    CUDA_TRY(h, cudaMalloc(&h->crop_scratch, need)); h->crop_scratch_cap = need;
  }
  h->launches += launch_crop(B, in_h, in_w, frames, poses, J, bbx, cfg_host6, out_hw, pad, icvl, h->crop_scratch, dm_out, cfg_out, com_out,
                             (cudaStream_t)stream);
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

int dr_crop_from_xyz_pose(dr_handle* h, int B, int in_h, int in_w, const float* frames, const float* poses, int J,
                          const float* cfg_host6, int out_hw, float pad, int icvl, float* dm_out, float* cfg_out, float* com_out, void* stream) {
  if (!poses || J < 1) return DR_ERR_ARG;
  return crop_common(h, B, in_h, in_w, frames, poses, J, nullptr, cfg_host6, out_hw, pad, icvl, dm_out, cfg_out, com_out, stream);
}

int dr_crop_from_bbx(dr_handle* h, int B, int in_h, int in_w, const float* frames, const float* bbx, const float* cfg_host6, int out_hw,
                     float* dm_out, float* cfg_out, float* com_out, void* stream) {
  if (!bbx) return DR_ERR_ARG;
  return crop_common(h, B, in_h, in_w, frames, nullptr, 0, bbx, cfg_host6, out_hw, 0.f, 0, dm_out, cfg_out, com_out, stream);
}

int dr_data_aug(dr_handle* h, int B, int hw, int J, const float* dms, const float* poses, const float* cfgs, const float* coms,
                const float* cossin, const float* edge_ratio, float* dms_out, float* poses_out, void* stream) {
  if (!h || !dms || !poses || !cfgs || !coms || !cossin || !edge_ratio || !dms_out || !poses_out || B < 1 || hw < 2 || J < 1) return DR_ERR_ARG;
  h->launches += launch_data_aug(B, hw, J, dms, poses, cfgs, coms, cossin, edge_ratio, dms_out, poses_out, (cudaStream_t)stream);
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

int dr_debug_conv(dr_handle* h, int layer, int B, const float* x, float* y, int precision, void* stream) {
  const bool reuse = (precision & 0x100) != 0 && h && h->wk;
  const int force_pair = (precision & 0x200) != 0;          // debug: CTA-pair kernel for this call regardless of the handle's setting
  const int precision_flags = precision;
  precision &= 0xff;
  if (!h || layer < 0 || layer >= (int)h->layers.size() || !x || !y || B < 1) return DR_ERR_ARG;
  if (!h->params) return fail(h, DR_ERR_STATE, "dr_bind() not called");
  { int rc = pipe_join(h, (cudaStream_t)stream); if (rc) return rc; }
  const Layer& L = h->layers[layer];
  ConvProblem p; memset(&p, 0, sizeof(p));
  p.x = x; p.x_cs = L.cin; p.B = B; p.H = L.in_hw; p.W = L.in_hw; p.Cin = L.cin; p.Ho = L.out_hw; p.Wo = L.out_hw; p.Cout = L.cout;
  p.k = L.k; p.stride = L.stride; p.pad_t = p.pad_l = same_pad_before(L.in_hw, L.k, L.stride);
  if (precision != DR_PREC_FP32 && !reuse) { int rc = ensure_prepped(h, precision, (cudaStream_t)stream); if (rc) return rc; }
  set_fwd_weights(h, L, precision, p); p.y = y; p.y_cs = L.cout; p.pair = force_pair ? 2 : 0;
  p.chunk_kb = (precision_flags & 0x400) ? h->chunk_eval : 0; p.chunk_min_kb = h->chunk_min_kb;          // debug flag 0x400: two-level accumulation as in inference
  { int64_t nl = 0; RUN_TRY(nl, run_conv(h, p, precision, (cudaStream_t)stream)); h->launches += nl; }
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

int dr_debug_conv_bwd(dr_handle* h, int layer, int B, const float* x, const float* dy, float* dx, float* dw, int precision, void* stream) {
  if (!h || layer < 0 || layer >= (int)h->layers.size() || !x || !dy || B < 1) return DR_ERR_ARG;
  if (!h->params) return fail(h, DR_ERR_STATE, "dr_bind() not called");
  cudaStream_t st = (cudaStream_t)stream;
  { int rc = pipe_join(h, st); if (rc) return rc; }
  const Layer& L = h->layers[layer];
  const size_t nw = (size_t)L.k * L.k * L.cin * L.cout;
  if (dw) {
    CUDA_TRY(h, cudaMemsetAsync(dw, 0, nw * sizeof(float), st));
    WgradProblem wp; memset(&wp, 0, sizeof(wp));
    wp.x = x; wp.x_cs = L.cin; wp.dy = dy; wp.dy_cs = L.cout; wp.B = B; wp.H = L.in_hw; wp.W = L.in_hw; wp.Cin = L.cin;
    wp.Ho = L.out_hw; wp.Wo = L.out_hw; wp.Cout = L.cout; wp.k = L.k; wp.stride = L.stride;
    wp.pad_t = wp.pad_l = same_pad_before(L.in_hw, L.k, L.stride); wp.dw = dw;
    { int64_t nl = 0; RUN_TRY(nl, run_wgrad(h, wp, precision, st)); h->launches += nl; }
  }
  if (dx) {
    if (L.stride != 1) return fail(h, DR_ERR_UNSUPPORTED, "dgrad only for stride-1 convs (the stem conv has no input gradient)");
    { int rc = ensure_prepped(h, precision, st); if (rc) return rc; }
    ConvProblem p; memset(&p, 0, sizeof(p));
    p.x = dy; p.x_cs = L.cout; p.B = B; p.H = L.out_hw; p.W = L.out_hw; p.Cin = L.cout; p.Ho = L.in_hw; p.Wo = L.in_hw; p.Cout = L.cin;
    p.k = L.k; p.stride = 1; p.pad_t = p.pad_l = L.k - 1 - same_pad_before(L.in_hw, L.k, 1);
    set_dgrad_weights(h, L, precision, p); p.y = dx; p.y_cs = L.cin;
    { int64_t nl = 0; RUN_TRY(nl, run_conv(h, p, precision, st)); h->launches += nl; }
  }
  CUDA_TRY(h, cudaGetLastError());
  return DR_OK;
}

}  // extern "C"
