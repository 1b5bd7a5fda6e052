"""DenseRegEngine: thin Python owner of the flat parameter/state/gradient tensors and the dr_handle.

Mirrors the reference's plugin seams:
  network.um_v1.detect_net(dm_inputs, cfgs, coms, num_jnt, is_training)   -> DenseRegEngine.forward
  JointDetectionModel._xyz_estimation(hms, oms, hm3s, dms, cfgs, coms)    -> DenseRegEngine.vote
  JointDetectionModel.test / loss / opt                                   -> infer / loss_backward / optimizer_step
All tensors are torch CUDA tensors used as raw buffers; every call goes through the C-ABI.
"""
import ctypes as C
import torch
from . import _ffi


class DenseRegError(RuntimeError):
    pass


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.int32), (t.dtype, t.is_cuda)
    return C.c_void_p(t.data_ptr())


class DenseRegEngine:
    def __init__(self, num_stack=2, num_fea=128, num_jnt=16, max_batch=40, precision="tf32x3", device=0,
                 kernel_size=3, training=True, infer_graph=False, tc_pair=True, pipeline=1):
        if not torch.cuda.is_available():
            raise DenseRegError("densereg_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _ffi.load()
        self.device = torch.device("cuda", device)
        self.S, self.F, self.J = num_stack, num_fea, num_jnt
        self.max_batch = max_batch
        cfg = _ffi.DrConfig(num_stack=num_stack, num_fea=num_fea, kernel_size=kernel_size, num_jnt=num_jnt,
                            in_hw=128, out_hw=32, max_batch=max_batch,
                            precision=_ffi.PRECISIONS[precision] if isinstance(precision, str) else precision,
                            device=device)
        cfg.reserved[0] = 1 if infer_graph else 0          # dr_infer via a captured CUDA graph (same buffers every call)
        cfg.reserved[1] = 0 if tc_pair else -1             # CTA-pair (cta_group::2) 3xTF32 conv kernel for the big layers (default on)
        cfg.reserved[2] = 2 if (pipeline == 2 and training) else 0   # micro-batch pipeline: forward of micro-batch i+1 next to backward of i (join())
        self._h = C.c_void_p()
        torch.cuda.set_device(self.device)
        rc = self.lib.dr_create(C.byref(self._h), C.byref(cfg))
        if rc != 0:
            raise DenseRegError("dr_create failed with %d" % rc)
        self.n_params = self.lib.dr_param_count(self._h)
        self.n_state = self.lib.dr_state_count(self._h)
        kw = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(self.n_params, **kw)
        self.state = torch.zeros(self.n_state, **kw)
        self._grads = torch.zeros(self.n_params, **kw) if training else None
        self.adam_m = torch.zeros(self.n_params, **kw) if training else None
        self.adam_v = torch.zeros(self.n_params, **kw) if training else None
        self._check(self.lib.dr_bind(self._h, _ptr(self.params), _ptr(self.state), _ptr(self._grads),
                                     _ptr(self.adam_m), _ptr(self.adam_v)))
        self.loss_buf = torch.zeros(5, **kw)

    # ------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise DenseRegError("libdensereg_sm100 error %d: %s" % (rc, self.lib.dr_last_error(self._h).decode()))

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.dr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def layers(self):
        out = []
        for i in range(self.lib.dr_num_layers(self._h)):
            li = _ffi.DrLayerInfo()
            self._check(self.lib.dr_get_layer(self._h, i, C.byref(li)))
            out.append(dict(name=li.name.decode(), k=li.k, stride=li.stride, cin=li.cin, cout=li.cout, brn=li.brn,
                            relu=li.relu, wd=li.wd, w_off=li.w_off, p_off=li.p_off, s_off=li.s_off,
                            in_hw=li.in_hw, out_hw=li.out_hw))
        return out

    def _layers_cached(self):
        if not hasattr(self, "_layers"):
            self._layers = self.layers()
        return self._layers

    def init_params(self, seed=0, stddev=0.01):
        self._check(self.lib.dr_init_params(self._h, seed, stddev, self._stream()))

    def load_flat(self, params, state=None):
        """Copy flat fp32 parameter (and BRN state) vectors (e.g. from a checkpoint) into the bound buffers."""
        self.params.copy_(params.to(self.device, torch.float32))
        if state is not None:
            self.state.copy_(state.to(self.device, torch.float32))
        self.params_changed()

    def params_changed(self):
        """Tell the library that `self.params` was written from outside (it rebuilds its tensor-core weight copies)."""
        self._check(self.lib.dr_params_changed(self._h))

    # ------------------------------------------------------------------------------------------
    def norm_dm(self, dm_mm, coms):
        B, hw = dm_mm.shape[0], dm_mm.shape[1]
        out = torch.empty_like(dm_mm)
        self._check(self.lib.dr_norm_dm(self._h, B, hw, _ptr(dm_mm), _ptr(coms), _ptr(out), self._stream()))
        return out

    def forward(self, dm_mm, coms, is_training=False, update_state=False, dropout_seed=0):
        """detect_net on raw depth crops (B,128,128,1) mm.  Returns dict of per-stack NHWC tensors like the
        reference's end_points {'hm_outs','hm3_outs','um_outs'} (network/um_v1.py:72-75,170-172)."""
        B = dm_mm.shape[0]
        kw = dict(dtype=torch.float32, device=self.device)
        hms = [torch.empty(B, 32, 32, self.J, **kw) for _ in range(self.S)]
        hm3s = [torch.empty(B, 32, 32, self.J, **kw) for _ in range(self.S)]
        ums = [torch.empty(B, 32, 32, 3 * self.J, **kw) for _ in range(self.S)]
        arr = lambda ts: (C.c_void_p * self.S)(*[t.data_ptr() for t in ts])
        self._check(self.lib.dr_forward(self._h, B, _ptr(dm_mm), _ptr(coms), arr(hms), arr(hm3s), arr(ums),
                                        int(is_training), int(update_state), dropout_seed, self._stream()))
        return {"hm_outs": hms, "hm3_outs": hm3s, "um_outs": ums}

    def vote(self, hm, hm3, um, dm_norm, cfgs, coms, return_top5=False):
        """_resume_om + _xyz_estimation + unnorm_xyz_pose on dense maps -> xyz mm (B,3J)."""
        B, H, W, J = hm.shape
        xyz = torch.empty(B, 3 * J, dtype=torch.float32, device=self.device)
        top5 = torch.empty(B, J, 5, dtype=torch.int32, device=self.device)
        clamp = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.dr_vote(self._h, B, H, W, J, _ptr(hm), _ptr(hm3), _ptr(um), _ptr(dm_norm), _ptr(cfgs),
                                     _ptr(coms), _ptr(xyz), _ptr(top5), _ptr(clamp), self._stream()))
        return (xyz, top5, clamp) if return_top5 else xyz

    def infer(self, dm_mm, cfgs, coms, out=None, top5=None):
        """JointDetectionModel.test: raw crops -> xyz mm (B,3J)."""
        B = dm_mm.shape[0]
        if out is None:
            out = torch.empty(B, 3 * self.J, dtype=torch.float32, device=self.device)
        self._check(self.lib.dr_infer(self._h, B, _ptr(dm_mm), _ptr(cfgs), _ptr(coms), _ptr(out), _ptr(top5), self._stream()))
        return out

    def crop_from_xyz_pose(self, frames, poses, cfg, out_hw=128, pad=20.0, icvl=False):
        """data/preprocess.py crop_from_xyz_pose + center_of_mass on full depth frames (B,in_h,in_w) -> (dms, cfgs, coms)."""
        B, in_h, in_w = frames.shape
        kw = dict(dtype=torch.float32, device=self.device)
        dms = torch.empty(B, out_hw, out_hw, 1, **kw); cfgs = torch.empty(B, 6, **kw); coms = torch.empty(B, 3, **kw)
        c6 = (C.c_float * 6)(*[float(x) for x in cfg])
        self._check(self.lib.dr_crop_from_xyz_pose(self._h, B, in_h, in_w, _ptr(frames), _ptr(poses), poses.shape[1] // 3, c6, out_hw,
                                                   float(pad), int(icvl), _ptr(dms), _ptr(cfgs), _ptr(coms), self._stream()))
        return dms, cfgs, coms

    def crop_from_bbx(self, frames, bbx, cfg, out_hw=128):
        """data/preprocess.py crop_from_bbx + center_of_mass (NYU test boxes [top,left,bottom,right,d_th])."""
        B, in_h, in_w = frames.shape
        kw = dict(dtype=torch.float32, device=self.device)
        dms = torch.empty(B, out_hw, out_hw, 1, **kw); cfgs = torch.empty(B, 6, **kw); coms = torch.empty(B, 3, **kw)
        c6 = (C.c_float * 6)(*[float(x) for x in cfg])
        self._check(self.lib.dr_crop_from_bbx(self._h, B, in_h, in_w, _ptr(frames), _ptr(bbx), c6, out_hw, _ptr(dms), _ptr(cfgs),
                                              _ptr(coms), self._stream()))
        return dms, cfgs, coms

    def data_aug(self, dms, poses, cfgs, coms, cossin, edge_ratio):
        """data/preprocess.py data_aug with caller-drawn randomness: cossin (B,2), edge_ratio (B,2) [h, w]."""
        B, hw = dms.shape[0], dms.shape[1]
        dms_out = torch.empty_like(dms); poses_out = torch.empty_like(poses)
        self._check(self.lib.dr_data_aug(self._h, B, hw, poses.shape[1] // 3, _ptr(dms), _ptr(poses), _ptr(cfgs), _ptr(coms), _ptr(cossin),
                                         _ptr(edge_ratio), _ptr(dms_out), _ptr(poses_out), self._stream()))
        return dms_out, poses_out

    # ---- data-parallel communicator (replaces model/train_multi_gpu.py:16-39) ---------------------------------
    def comm_init(self, rank, world, unique_id=None):
        """Create the in-library NCCL communicator.  `unique_id`: 128 bytes from rank 0's comm_unique_id(); when None and
        torch.distributed is initialised, rank 0's id is broadcast through it (plumbing only -- the gradient all-reduce itself runs
        inside libdensereg_sm100.so)."""
        self.rank, self.world = rank, world
        if world == 1:
            return
        if unique_id is None:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise DenseRegError("comm_init needs a unique id or an initialised torch.distributed group to broadcast it")
            dev = self.device if dist.get_backend() == "nccl" else torch.device("cpu")
            buf = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = torch.frombuffer(bytearray(self.comm_unique_id()), dtype=torch.uint8).clone()
            buf = buf.to(dev)
            dist.broadcast(buf, src=0)
            unique_id = bytes(buf.cpu().numpy().tobytes())
        idbuf = (C.c_char * 128).from_buffer_copy(unique_id)
        self._check(self.lib.dr_comm_init(self._h, rank, world, C.cast(idbuf, C.c_void_p)))

    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        rc = self.lib.dr_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise DenseRegError("dr_comm_unique_id failed with %d (libnccl.so.2 not found?)" % rc)
        return bytes(buf.raw)

    def comm_overlap_next_backward(self):
        """The next loss_backward() is the last micro-batch of the optimiser step: all-reduce gradient buckets as they become final."""
        self._check(self.lib.dr_comm_overlap_next_backward(self._h))

    @property
    def allreduce_count(self):
        return int(self.lib.dr_comm_allreduce_count(self._h))

    def join(self):
        """Micro-batch pipeline (pipeline=2): order the current stream after every backward pass still in flight.  zero_grads /
        optimizer_step / forward / infer join by themselves, and so does the `grads` property; only a caller that kept its own
        reference to the gradient tensor needs this."""
        self._check(self.lib.dr_pipeline_join(self._h, self._stream()))

    @property
    def grads(self):
        """The bound flat gradient buffer (accum_op's accumulators, train_single_gpu.py:69-84), complete on the current stream: with the
        micro-batch pipeline on, every backward pass still in flight is joined first."""
        if self._grads is not None and getattr(self, "_h", None) is not None and self._h.value and self.lib.dr_pipeline_depth(self._h) == 2:
            self.join()
        return self._grads

    @property
    def pipeline_depth(self):
        return int(self.lib.dr_pipeline_depth(self._h))

    def zero_grads(self):
        self._check(self.lib.dr_zero_grads(self._h, self._stream()))

    def loss_backward(self, dm_mm, poses_mm, cfgs, coms, dropout_seed=0, update_state=True):
        """One micro-batch of JointDetectionModel.loss + backward; grads accumulate.  Returns the device
        tensor {total, hm, hm3, um, reg} (no sync)."""
        B = dm_mm.shape[0]
        self._check(self.lib.dr_loss_backward(self._h, B, _ptr(dm_mm), _ptr(poses_mm), _ptr(cfgs), _ptr(coms),
                                              _ptr(self.loss_buf), dropout_seed, int(update_state), self._stream()))
        return self.loss_buf

    def optimizer_step(self, step, lr, accum_steps=1, world=1):
        self._check(self.lib.dr_optimizer_step(self._h, accum_steps, world, float(lr), int(step), self._stream()))

    def debug_conv(self, layer, x, precision="fp32", reuse_weights=False, out=None, pair=False):
        L = self._layers_cached()[layer]
        B = x.shape[0]
        y = out if out is not None else torch.empty(B, L["out_hw"], L["out_hw"], L["cout"], dtype=torch.float32, device=self.device)
        self._check(self.lib.dr_debug_conv(self._h, layer, B, _ptr(x), _ptr(y), _ffi.PRECISIONS[precision] | (0x100 if reuse_weights else 0) | (0x200 if pair else 0),
                                           self._stream()))
        return y

    def debug_conv_bwd(self, layer, x, dy, precision="fp32", want_dx=True):
        L = self._layers_cached()[layer]
        dx = torch.empty_like(x) if want_dx else None
        dw = torch.empty(L["k"] * L["k"] * L["cin"] * L["cout"], dtype=torch.float32, device=self.device)
        self._check(self.lib.dr_debug_conv_bwd(self._h, layer, x.shape[0], _ptr(x), _ptr(dy), _ptr(dx), _ptr(dw),
                                               _ffi.PRECISIONS[precision], self._stream()))
        return dx, dw

    def debug_get_output(self, layer, B, grad=False):
        L = self._layers_cached()[layer]
        out = torch.empty(B, L["out_hw"], L["out_hw"], L["cout"], dtype=torch.float32, device=self.device)
        self._check(self.lib.dr_debug_get_output(self._h, layer, B, _ptr(out), int(grad), self._stream()))
        return out

    def trace(self, on=True):
        """Per-launch timing of the conv-type kernels (serialises them; measurement passes only)."""
        self._check(self.lib.dr_trace(self._h, int(on)))

    def trace_records(self):
        out = []
        rec = _ffi.DrTraceRec()
        for i in range(self.lib.dr_trace_count(self._h)):
            self._check(self.lib.dr_trace_get(self._h, i, C.byref(rec)))
            out.append(dict(kind=("conv", "dgrad", "wgrad")[rec.kind], B=rec.B, hw=rec.hw, cin=rec.cin, cout=rec.cout, k=rec.k,
                            kernel=("ffma", "tcgen05", "tcgen05_pair")[rec.kernel], ms=rec.ms))
        return out

    @property
    def launch_count(self):
        return int(self.lib.dr_launch_count(self._h))

    @property
    def tc_launch_count(self):
        return int(self.lib.dr_tc_launch_count(self._h))

    @property
    def workspace_bytes(self):
        return int(self.lib.dr_workspace_bytes(self._h))
