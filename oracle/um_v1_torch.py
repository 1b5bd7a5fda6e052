"""Oracle (test infrastructure): PyTorch-CPU fp32 restatement of the um_v1 network, the
training loss with its ground-truth synthesis, and the optimiser step.

PARITY UNPINNED (see oracle/__init__.py): follows the reference source + TF 1.x op semantics
(SURVEY.md appendix B); TF 1.3 itself cannot run here.

Reference functions restated (all in /root/reference):
  network/um_v1.py:18-48     _residual        -> Net._residual
  network/um_v1.py:51-69     _hourglass       -> Net._hourglass
  network/um_v1.py:71-185    detect_net       -> Net.forward
  network/slim/ops.py:43-185 batch_norm (Batch ReNorm) -> Net._brn
  network/slim/ops.py:219-299 conv2d          -> Net._conv
  network/slim/ops.py:640-677 max_pool, upsampling_nearest
  network/slim/ops.py:710-728 dropout         -> Net._dropout (mask from the shared counter hash)
  network/slim/losses.py:56-72 l2_regularizer -> loss_and_grads (reg term)
  model/hourglass_um_crop_tiny.py:195-274  _hm_3d, _hm_2d, _um -> gt_maps
  model/hourglass_um_crop_tiny.py:323-371  loss
  model/hourglass_um_crop_tiny.py:436-439, model/train_single_gpu.py:45-49,69-88 -> adam_step, lr_at

Parameter order == TF variable creation order of um_v1 (needed for the flat buffers shared
with the CUDA engine): per conv `weights` (HWIO), then `beta`,`gamma` (BRN) or `biases`.
State per BRN conv: moving_mean[C], moving_variance[C], biased_mean[C], biased_variance[C],
then r_max, d_max, curr_t, local_step (zero-debias slots of assign_moving_average, TF 1.3).
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

BRN_DECAY = 0.99      # um_v1.py:9
BRN_EPS = 0.001       # um_v1.py:10
WD = 0.0005           # um_v1.py:35,87,131,...
MAX_DIST_2D = 4.0     # hourglass_um_crop_tiny.py:193
MAX_DIST_3D = 0.8     # hourglass_um_crop_tiny.py:194


# --------------------------------------------------------------------------------------
# layer table (creation order)
# --------------------------------------------------------------------------------------
class ConvSpec:
    __slots__ = ("name", "k", "stride", "cin", "cout", "brn", "relu", "wd", "w_off", "p_off", "s_off", "idx")

    def __init__(self, name, k, stride, cin, cout, brn, relu, wd):
        self.name, self.k, self.stride, self.cin, self.cout = name, k, stride, cin, cout
        self.brn, self.relu, self.wd = brn, relu, wd


def build_specs(num_stack, num_fea, num_jnt):
    """Enumerate convs in the order um_v1.detect_net creates their variables."""
    specs = []

    def conv(name, k, cin, cout, stride=1, brn=True, relu=True, wd=WD):
        specs.append(ConvSpec(name, k, stride, cin, cout, brn, relu, wd))

    def res(name, cin, cout=None):
        cout = cin if cout is None else cout
        h = cin // 2
        conv(name + "/c1", 1, cin, h)
        conv(name + "/c2", 3, h, h)
        conv(name + "/c3", 1, h, cout)
        if cout != cin:
            conv(name + "/skip", 1, cin, cout)

    def hourglass(name, n, Fc):
        res(f"{name}/n{n}/upper1", Fc)
        res(f"{name}/n{n}/lower1", Fc)
        if n > 1:
            hourglass(name, n - 1, Fc)
        res(f"{name}/n{n}/lower3", Fc)

    Fc, J = num_fea, num_jnt
    conv("stem/conv_1", 7, 1, 32, stride=2)
    res("stem/conv_2", 32, 64)
    res("stem/conv_3", 64)
    res("stem/conv_4", 64, Fc)
    for s in range(num_stack):
        p = f"s{s}"
        hourglass(p + "/hg", 4, Fc)
        res(p + "/ll_res", Fc)
        conv(p + "/ll", 1, Fc, Fc)
        conv(p + "/hm_out", 1, Fc, J, brn=False, relu=False)
        res(p + "/hm3_res", Fc + 3, 128)
        conv(p + "/hm3_out", 1, 128, J, brn=False, relu=False)
        res(p + "/um_res1", Fc + 2 * J, 256)
        res(p + "/um_res2", 256)
        res(p + "/um_mask_res1", Fc + 2 * J, 256)
        res(p + "/um_mask_res2", 256)
        res(p + "/um_comb", 512)
        conv(p + "/um_full1", 1, 515, 512, brn=False, relu=True)
        conv(p + "/um_full2", 1, 512, 512, brn=False, relu=True)
        conv(p + "/um_out", 1, 512, 3 * J, brn=False, relu=False)
        if s < num_stack - 1:
            conv(p + "/inter_out", 1, 5 * J, Fc, brn=False, relu=False, wd=0.0)
            conv(p + "/inter_ll", 1, Fc, Fc, brn=False, relu=False, wd=0.0)
    p_off = 0
    s_off = 0
    for i, c in enumerate(specs):
        c.idx = i
        c.w_off = p_off
        p_off += c.k * c.k * c.cin * c.cout
        c.p_off = p_off                      # beta|gamma or biases
        p_off += 2 * c.cout if c.brn else c.cout
        c.s_off = s_off
        if c.brn:
            s_off += 4 * c.cout + 4
    return specs, p_off, s_off


def same_pad(n, k, s):
    """TF SAME padding."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def dropout_mask(seed, layer_tag, n):
    """Shared counter-based hash (identical in densereg_b200/csrc/common.cuh: dr_hash_keep).
    keep iff top bit of mix(seed, tag, i) is set; keep prob 0.5 (ops.py:711)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over='ignore'):
        i = np.arange(n, dtype=np.uint64)
        x = (i + np.uint64(seed & 0xFFFFFFFFFFFFFFFF) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(layer_tag) * np.uint64(0xD1B54A32D192ED03)) & M
        x ^= x >> np.uint64(30); x = (x * np.uint64(0xBF58476D1CE4E5B9)) & M
        x ^= x >> np.uint64(27); x = (x * np.uint64(0x94D049BB133111EB)) & M
        x ^= x >> np.uint64(31)
    return ((x >> np.uint64(63)) & np.uint64(1)).astype(np.float32)


# --------------------------------------------------------------------------------------
# network
# --------------------------------------------------------------------------------------
class Net:
    """Functional um_v1 over flat fp32 buffers (params, state)."""

    def __init__(self, num_stack=2, num_fea=128, num_jnt=16):
        self.S, self.F, self.J = num_stack, num_fea, num_jnt
        self.specs, self.n_params, self.n_state = build_specs(num_stack, num_fea, num_jnt)
        self.by_name = {c.name: c for c in self.specs}

    # ---- parameter / state initialisation -------------------------------------------
    def init_params(self, seed=0, stddev=0.01):
        """ops.py:272 truncated_normal(stddev), biases 0 (ops.py:291), beta 0 / gamma 1 (ops.py:86-96)."""
        g = torch.Generator().manual_seed(seed)
        p = torch.zeros(self.n_params, dtype=torch.float32)
        for c in self.specs:
            n = c.k * c.k * c.cin * c.cout
            w = torch.empty(n)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=stddev, a=-2 * stddev, b=2 * stddev, generator=g)
            p[c.w_off:c.w_off + n] = w
            if c.brn:
                p[c.p_off + c.cout:c.p_off + 2 * c.cout] = 1.0     # gamma
        return p

    def init_state(self):
        """moving_mean 0, moving_variance 1, r_max 1, d_max 0, curr_t 0 (ops.py:100-128)."""
        s = torch.zeros(self.n_state, dtype=torch.float32)
        for c in self.specs:
            if c.brn:
                C = c.cout
                s[c.s_off + C:c.s_off + 2 * C] = 1.0
                s[c.s_off + 4 * C] = 1.0
        return s

    # ---- ops ---------------------------------------------------------------------------
    def _conv(self, x, c, training, residual=None):
        """ops.conv2d (ops.py:219-299): conv SAME -> BRN | bias -> optional ReLU."""
        P, St = self.params, self.state
        n = c.k * c.k * c.cin * c.cout
        w = P[c.w_off:c.w_off + n].view(c.k, c.k, c.cin, c.cout).permute(3, 2, 0, 1)   # HWIO -> OIHW
        H, W = x.shape[2], x.shape[3]
        pt, pb = same_pad(H, c.k, c.stride)
        pl, pr = same_pad(W, c.k, c.stride)
        if pt or pb or pl or pr:
            x = F.pad(x, (pl, pr, pt, pb))
        y = F.conv2d(x, w, None, stride=c.stride)
        self.trace[c.name + ":raw"] = y
        if c.brn:
            y = self._brn(y, c, training)
        else:
            y = y + P[c.p_off:c.p_off + c.cout].view(1, -1, 1, 1)
        if c.relu:
            y = F.relu(y)
        self.trace[c.name] = y
        return y

    def _brn(self, x, c, training):
        """ops.batch_norm == Batch ReNorm (ops.py:43-185)."""
        P, St = self.params, self.state
        C = c.cout
        beta = P[c.p_off:c.p_off + C].view(1, -1, 1, 1)
        gamma = P[c.p_off + C:c.p_off + 2 * C].view(1, -1, 1, 1)
        o = c.s_off
        mov_mean, mov_var = St[o:o + C], St[o + C:o + 2 * C]
        if not training:                                      # ops.py:173-180
            inv = torch.rsqrt(mov_var + BRN_EPS).view(1, -1, 1, 1) * gamma
            return x * inv + (beta - mov_mean.view(1, -1, 1, 1) * inv)
        mean = x.mean(dim=(0, 2, 3))                          # tf.nn.moments (ops.py:132)
        var = ((x - mean.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))
        with torch.no_grad():                                 # stop_gradient (ops.py:159,162)
            r_max, d_max = St[o + 4 * C].item(), St[o + 4 * C + 1].item()   # OLD values (unordered in TF)
            std = torch.sqrt(var + BRN_EPS)
            mov_std = torch.sqrt(mov_var + BRN_EPS)
            r = torch.clamp(std / mov_std, 1.0 / r_max, r_max)
            d = torch.clamp((mean - mov_mean) / mov_std, -d_max, d_max)
            self.pending_updates.append((c, mean.detach().clone(), var.detach().clone()))
        inv = torch.rsqrt(var + BRN_EPS).view(1, -1, 1, 1)
        y = x * inv + (-mean.view(1, -1, 1, 1) * inv)         # tf.nn.batch_normalization, no gamma/beta
        y = y * r.view(1, -1, 1, 1) + d.view(1, -1, 1, 1)     # ops.py:166
        return y * gamma + beta                               # ops.py:168-171

    def apply_state_updates(self):
        """UPDATE_OPS of one micro-step (ops.py:134-153; assign_moving_average with zero_debias,
        SURVEY.md appendix B.5/B.6).  Called after the forward pass (forward reads old values)."""
        St = self.state
        with torch.no_grad():
            for c, mean, var in self.pending_updates:
                C, o = c.cout, c.s_off
                t = St[o + 4 * C + 2].item()
                step = St[o + 4 * C + 3].item() + 1.0
                bm, bv = St[o + 2 * C:o + 3 * C], St[o + 3 * C:o + 4 * C]
                bm -= (bm - mean) * (1.0 - BRN_DECAY)
                bv -= (bv - var) * (1.0 - BRN_DECAY)
                corr = 1.0 - BRN_DECAY ** step
                St[o:o + C] = bm / corr
                St[o + C:o + 2 * C] = bv / corr
                St[o + 4 * C] = 3.0 / (1.0 + 2.0 * math.exp(-t))              # r_max  ops.py:141-144
                St[o + 4 * C + 1] = 5.0 / (5000.0 * math.exp(-2.0 * t))       # d_max  ops.py:146-149
                St[o + 4 * C + 2] = float(np.float32(t) + np.float32(1e-5))   # curr_t ops.py:151-153
                St[o + 4 * C + 3] = step
        self.pending_updates = []

    def _residual(self, x, name, training):
        """um_v1.py:18-48."""
        bn = self.by_name
        o = self._conv(x, bn[name + "/c1"], training)
        o = self._conv(o, bn[name + "/c2"], training)
        o = self._conv(o, bn[name + "/c3"], training)
        sk = self._conv(x, bn[name + "/skip"], training) if (name + "/skip") in bn else x
        return o + sk

    @staticmethod
    def _maxpool(x, k):
        """ops.max_pool (ops.py:640-669), SAME, stride 2; padding never wins (-inf)."""
        pt, pb = same_pad(x.shape[2], k, 2)
        pl, pr = same_pad(x.shape[3], k, 2)
        if pt or pb or pl or pr:
            x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
        return F.max_pool2d(x, k, 2)

    def _hourglass(self, x, name, n, training):
        """um_v1.py:51-69."""
        up1 = self._residual(x, f"{name}/n{n}/upper1", training)
        low = self._maxpool(x, 3)
        low = self._residual(low, f"{name}/n{n}/lower1", training)
        if n > 1:
            low = self._hourglass(low, name, n - 1, training)
        low = self._residual(low, f"{name}/n{n}/lower3", training)
        up2 = low.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)   # ops.py:671-677 NN x2
        return up1 + up2

    def _dropout(self, x, tag):
        """ops.dropout keep_prob 0.5 (ops.py:710-728): x * mask / 0.5; mask indexed in NHWC order."""
        B, C, H, W = x.shape
        m = dropout_mask(self.dropout_seed, tag, B * H * W * C).reshape(B, H, W, C)
        m = torch.from_numpy(m).permute(0, 3, 1, 2).to(x.dtype)
        return x * m * 2.0

    def forward(self, params, state, x0_nhwc, training=False, dropout_seed=0):
        """detect_net (um_v1.py:71-185).  x0_nhwc: (B,128,128,1) NORMALISED depth (norm_dm output).
        Returns lists hm_outs, hm3_outs, um_outs of NHWC tensors (B,32,32,{J,J,3J})."""
        self.params, self.state = params, state
        self.trace = {}
        self.pending_updates = []
        self.dropout_seed = dropout_seed
        bn, Fc, J = self.by_name, self.F, self.J
        x = x0_nhwc.permute(0, 3, 1, 2)
        B = x.shape[0]
        c1 = self._conv(x, bn["stem/conv_1"], training)
        c2 = self._residual(c1, "stem/conv_2", training)
        p1 = self._maxpool(c2, 2)
        c3 = self._residual(p1, "stem/conv_3", training)
        hg_ins = self._residual(c3, "stem/conv_4", training)
        tiny = x[:, :, ::4, ::4]                                           # um_v1.py:111 (appendix B.3)
        oh, ow = tiny.shape[2], tiny.shape[3]
        uu = (torch.arange(ow, dtype=torch.float32) / float(ow / 2) - 1.0).to(x.dtype).view(1, 1, 1, ow).expand(B, 1, oh, ow)
        vv = (torch.arange(oh, dtype=torch.float32) / float(oh / 2) - 1.0).to(x.dtype).view(1, 1, oh, 1).expand(B, 1, oh, ow)
        uvd = torch.cat([uu, vv, tiny], dim=1)                             # um_v1.py:121
        hms, hm3s, ums = [], [], []
        for s in range(self.S):
            p = f"s{s}"
            hg_outs = self._hourglass(hg_ins, p + "/hg", 4, training)
            ll = self._residual(hg_outs, p + "/ll_res", training)
            ll = self._conv(ll, bn[p + "/ll"], training)
            hm = self._conv(ll, bn[p + "/hm_out"], training)
            hm3_in = self._residual(torch.cat([ll, uvd], 1), p + "/hm3_res", training)
            hm3 = self._conv(hm3_in, bn[p + "/hm3_out"], training)
            cat = torch.cat([hg_outs, hm, hm3], 1)
            um_in = self._residual(self._residual(cat, p + "/um_res1", training), p + "/um_res2", training)
            mask = (tiny < -0.9).expand(-1, cat.shape[1], -1, -1)          # um_v1.py:147
            cat_m = torch.where(mask, torch.zeros_like(cat), cat)
            um_m = self._residual(self._residual(cat_m, p + "/um_mask_res1", training), p + "/um_mask_res2", training)
            comb = self._residual(torch.cat([um_in, um_m], 1), p + "/um_comb", training)
            comb = torch.cat([comb, uvd], 1)
            f1 = self._conv(comb, bn[p + "/um_full1"], training)
            if training:
                f1 = self._dropout(f1, 2 * s)
            f2 = self._conv(f1, bn[p + "/um_full2"], training)
            if training:
                f2 = self._dropout(f2, 2 * s + 1)
            um = self._conv(f2, bn[p + "/um_out"], training)
            hms.append(hm.permute(0, 2, 3, 1)); hm3s.append(hm3.permute(0, 2, 3, 1)); ums.append(um.permute(0, 2, 3, 1))
            if s < self.S - 1:
                t = self._conv(torch.cat([hm, hm3, um], 1), bn[p + "/inter_out"], training)
                it = self._conv(ll, bn[p + "/inter_ll"], training)
                hg_ins = hg_ins + t + it                                   # um_v1.py:183
        return hms, hm3s, ums

    def reg_loss(self, params):
        """losses.l2_regularizer (losses.py:56-72): wd * l2_loss(w) per regularised conv."""
        tot = 0.0
        for c in self.specs:
            if c.wd > 0:
                n = c.k * c.k * c.cin * c.cout
                tot = tot + c.wd * 0.5 * (params[c.w_off:c.w_off + n] ** 2).sum()
        return tot


# --------------------------------------------------------------------------------------
# ground-truth synthesis + loss  (hourglass_um_crop_tiny.py:195-274, 323-371)
# --------------------------------------------------------------------------------------
def gt_maps(dm_mm, poses_mm, cfgs, coms, out_hw=32):
    """Returns x0 (B,128,128,1) normalised depth and gt_hm, gt_hm3 (B,h,w,J), gt_um (B,h,w,3J) as torch
    tensors.  Restates _hm_2d (:213-247), norm_xyz_pose, norm_dm, generate_xyzs (preprocess.py),
    _hm_3d (:195-211), _um (:249-274)."""
    from . import vote_numpy as V
    f32 = np.float32
    dm_mm = np.asarray(dm_mm, f32); poses_mm = np.asarray(poses_mm, f32)
    cfgs = np.asarray(cfgs, f32); coms = np.asarray(coms, f32)
    B = dm_mm.shape[0]; J = poses_mm.shape[1] // 3
    h = w = out_hw
    x0 = V.norm_dm(dm_mm.reshape(B, dm_mm.shape[1], dm_mm.shape[2]), coms)
    d32 = V.tiny_dm(x0, out_hw)
    P = V.generate_xyzs(d32, cfgs, coms)                                   # (B,h,w,3)
    pose_n = ((poses_mm.reshape(B, J, 3) - coms[:, None, :]) / f32(100.0)).astype(f32)
    gt_om = (pose_n.reshape(B, 1, 1, J, 3) - P[:, :, :, None, :]).astype(f32)       # :343
    dist = np.sqrt((gt_om[..., 0] ** 2 + gt_om[..., 1] ** 2).astype(f32) + gt_om[..., 2] ** 2).astype(f32)
    gt_hm3 = np.maximum((f32(0.8) - dist) / f32(0.8), f32(0.0)).astype(f32)         # :206-208
    d = (f32(0.8) - gt_hm3 * f32(0.8)).astype(f32)                                   # :260
    m = d < f32(0.8 - 1e-2)                                                         # :269
    with np.errstate(divide='ignore', invalid='ignore'):
        gt_um = np.where(m[..., None], gt_om / d[..., None], f32(0.0)).astype(f32)
    fx, fy, cx, cy = V.scaled_cfg(cfgs, w, h)
    p3 = poses_mm.reshape(B, J, 3)
    uu = (p3[..., 0] * fx[:, None]) / p3[..., 2] + cx[:, None]                       # util.py:20
    vv = (p3[..., 1] * fy[:, None]) / p3[..., 2] + cy[:, None]
    xx = np.arange(w, dtype=f32).reshape(1, 1, w, 1)
    yy = np.arange(h, dtype=f32).reshape(1, h, 1, 1)
    dd = np.sqrt(((xx - uu[:, None, None, :]) ** 2 + (yy - vv[:, None, None, :]) ** 2).astype(f32))
    gt_hm = (np.maximum(f32(4.0) - dd, f32(0.0)) / f32(4.0)).astype(f32)            # :243-244
    t = torch.from_numpy
    return (t(x0[..., None].copy()), t(gt_hm), t(gt_hm3), t(gt_um.reshape(B, h, w, 3 * J)))


def loss_and_grads(net, params, state, dm_mm, poses_mm, cfgs, coms, dropout_seed=0, update_state=True, dtype=torch.float32):
    """One micro-batch of hourglass_um_crop_tiny.py:323-371 (no data_aug) + autograd.
    Returns dict(total, hm, hm3, um, reg), grad (flat, same layout as params), outputs.
    dtype=torch.float64 runs the same graph in double (used by the tests to measure the fp32 noise floor of
    the gradient: ReLU / BRN boundary flips make two fp32 implementations differ by ~3e-3 in the deep layers)."""
    x0, gt_hm, gt_hm3, gt_um = [t.to(dtype) for t in gt_maps(dm_mm, poses_mm, cfgs, coms)]
    p = params.detach().clone().to(dtype).requires_grad_(True)
    st = state if dtype == torch.float32 else state.to(dtype)
    hms, hm3s, ums = net.forward(p, st, x0, training=True, dropout_seed=dropout_seed)
    hm_l = sum(0.5 * ((e - gt_hm) ** 2).sum() for e in hms)               # tf.nn.l2_loss :353
    hm3_l = sum(0.5 * ((e - gt_hm3) ** 2).sum() for e in hm3s)            # :357
    um_l = sum(0.5 * ((e - gt_um) ** 2).sum() for e in ums)               # :363
    reg = net.reg_loss(p)                                                  # :366
    total = reg + hm_l + um_l + hm3_l                                      # :371
    total.backward()
    if update_state:
        net.apply_state_updates()
    return (dict(total=total.item(), hm=hm_l.item(), hm3=hm3_l.item(), um=um_l.item(), reg=reg.item()),
            p.grad.detach(), (hms, hm3s, ums))


def lr_at(step, decay_steps, init_lr=1e-3, factor=0.1):
    """tf.train.exponential_decay(staircase=True) (train_single_gpu.py:45-49)."""
    return init_lr * factor ** (step // decay_steps)


def adam_step(params, grad_sum, m, v, step, lr, accum_steps=5, world=1, beta1=0.5, beta2=0.999, eps=1e-8, clip=0.2):
    """train_single_gpu.py:84-88 + tf.train.AdamOptimizer.
    grad_sum = sum of micro-batch grads (over accum_steps and ranks). step = 1-based Adam step.
    Updates params, m, v in place (fp32).

    Arithmetic form: TF's ApplyAdam kernel (tensorflow/core/kernels/training_ops.cc, TF 1.3, un-vendored) evaluates the recurrences of
    B.12 entirely in fp32 as
        alpha = lr * sqrt(1 - beta2_power) / (1 - beta1_power)      beta*_power = fp32 variables multiplied by beta* once per step
        m += (g - m) * (1 - beta1);   v += (g*g - v) * (1 - beta2);   var -= (m * alpha) / (sqrt(v) + eps)
    -- in particular (1 - beta2) is the fp32 difference 1.0f - 0.999f = 0.000999987..., not 0.001 (1.3e-5 apart)."""
    f32 = np.float32
    with torch.no_grad():
        g = torch.clamp(grad_sum / float(accum_steps * world), -clip, clip)
        b1p, b2p = f32(1.0), f32(1.0)
        for _ in range(int(step)):
            b1p = f32(b1p * f32(beta1)); b2p = f32(b2p * f32(beta2))
        alpha = f32(f32(lr) * np.sqrt(f32(1.0) - b2p, dtype=f32) / (f32(1.0) - b1p))
        m.add_((g - m) * float(f32(1.0) - f32(beta1)))
        v.add_((g * g - v) * float(f32(1.0) - f32(beta2)))
        params.sub_((m * float(alpha)) / (torch.sqrt(v) + eps))
    return params
