"""GPU parity: single convolutions (dr_debug_conv / dr_debug_conv_bwd through the C-ABI) vs the oracle's
F.conv2d with explicit TF SAME padding, forward, dgrad and wgrad.  fp32 bar: 2e-5 of the output scale."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
from gpu_util import cu, dump, relerr

pytestmark = pytest.mark.gpu

LAYERS = ["stem/conv_1", "stem/conv_2/c2", "stem/conv_2/skip", "stem/conv_4/c3", "s0/hg/n4/upper1/c2", "s0/hg/n1/lower1/c2",
          "s0/hg/n2/lower3/c1", "s0/hm_out", "s0/hm3_res/c1", "s0/hm3_res/c2", "s0/hm3_res/skip", "s0/um_res1/c2",
          "s0/um_comb/c2", "s0/um_full1", "s0/um_out", "s0/inter_out"]


@pytest.fixture(scope="module")
def setup(built_lib):
    from densereg_b200.engine import DenseRegEngine
    from oracle import um_v1_torch as U
    eng = DenseRegEngine(num_stack=2, num_fea=128, num_jnt=14, max_batch=2, training=False)
    net = U.Net(2, 128, 14)
    p = net.init_params(1, stddev=0.1)
    eng.load_flat(p)
    return eng, net, p


def oracle_conv(net, p, c, x_nhwc):
    from oracle.um_v1_torch import same_pad
    n = c.k * c.k * c.cin * c.cout
    w = p[c.w_off:c.w_off + n].view(c.k, c.k, c.cin, c.cout).permute(3, 2, 0, 1)
    x = x_nhwc.permute(0, 3, 1, 2)
    pt, pb = same_pad(x.shape[2], c.k, c.stride); pl, pr = same_pad(x.shape[3], c.k, c.stride)
    return F.conv2d(F.pad(x, (pl, pr, pt, pb)), w, None, stride=c.stride).permute(0, 2, 3, 1)


@pytest.mark.parametrize("precision", ["fp32"])
def test_conv_forward_and_backward(setup, precision):
    eng, net, p = setup
    names = [l["name"] for l in eng.layers()]
    rep = {}
    for name in LAYERS:
        c = net.by_name[name]; li = names.index(name)
        hw = eng.layers()[li]["in_hw"]
        g = torch.Generator().manual_seed(li)
        x = torch.randn(2, hw, hw, c.cin, generator=g)
        xr = x.clone().requires_grad_(True)
        pr = p.clone().requires_grad_(True)
        y_ref = oracle_conv(net, pr, c, xr)
        dy = torch.randn(y_ref.shape, generator=g)
        y_ref.backward(dy)
        y = eng.debug_conv(li, cu(x), precision)
        torch.cuda.synchronize()
        e = dict(fwd=relerr(y.cpu().numpy(), y_ref.detach().numpy()))
        dx, dw = eng.debug_conv_bwd(li, cu(x), cu(dy), precision, want_dx=(c.stride == 1))
        n = c.k * c.k * c.cin * c.cout
        e["wgrad"] = relerr(dw.cpu().numpy(), pr.grad[c.w_off:c.w_off + n].numpy())
        if c.stride == 1:
            e["dgrad"] = relerr(dx.cpu().numpy(), xr.grad.numpy())
        rep[name] = e
    dump("conv_err_%s.json" % precision, rep)
    tol = 2e-5 if precision == "fp32" else 2e-3
    bad = {k: v for k, v in rep.items() if max(v.values()) > tol}
    assert not bad, bad


TC_LAYERS = ["stem/conv_2/c2", "stem/conv_2/skip", "stem/conv_4/c3", "s0/hg/n4/upper1/c2", "s0/hg/n3/upper1/c1", "s0/hg/n2/lower3/c1",
             "s0/hg/n1/lower1/c2", "s0/hg/n1/upper1/c3", "s0/hm_out", "s0/um_res1/c1", "s0/um_res1/c2", "s0/um_comb/c2", "s0/um_full2", "s0/um_out",
             "s0/um_res2/c3"]


@pytest.mark.parametrize("precision,tol", [("tf32", 4e-3), ("tf32x3", 6e-5)])
@pytest.mark.parametrize("B", [2, 5])
def test_conv_tensor_core_path(setup, precision, tol, B):
    """tcgen05 implicit GEMM (TMA + TMEM) forward and dgrad vs the oracle conv.  tf32: one pass, inputs truncated to
    10 mantissa bits (bar 4e-3 of the output scale); tf32x3: split hi/lo, fp32-class (bar 6e-5: the tensor core's fp32 accumulation truncates, ~2e-5 at K=2304)."""
    eng, net, p = setup
    names = [l["name"] for l in eng.layers()]
    rep = {}
    for name in TC_LAYERS:
        c = net.by_name[name]; li = names.index(name)
        hw = eng.layers()[li]["in_hw"]
        g = torch.Generator().manual_seed(li + 7)
        x = torch.randn(B, hw, hw, c.cin, generator=g)
        xr = x.clone().requires_grad_(True)
        y_ref = oracle_conv(net, p, c, xr)
        dy = torch.randn(y_ref.shape, generator=g)
        y_ref.backward(dy)
        t0 = eng.tc_launch_count
        y = eng.debug_conv(li, cu(x), precision)
        pr = p.clone().requires_grad_(True)
        oracle_conv(net, pr, c, x).backward(dy)
        dx, dw = eng.debug_conv_bwd(li, cu(x), cu(dy), precision)
        torch.cuda.synchronize()
        nw = c.k * c.k * c.cin * c.cout
        rep[name] = dict(fwd=relerr(y.cpu().numpy(), y_ref.detach().numpy()), dgrad=relerr(dx.cpu().numpy(), xr.grad.numpy()),
                         wgrad=relerr(dw.cpu().numpy(), pr.grad[c.w_off:c.w_off + nw].numpy()),
                         tc_launches=eng.tc_launch_count - t0)
    dump("conv_tc_err_%s_B%d.json" % (precision, B), rep)
    bad = {k: v for k, v in rep.items() if max(v["fwd"], v["dgrad"], v["wgrad"]) > tol}
    assert not bad, bad
    assert sum(v["tc_launches"] for v in rep.values()) >= len(TC_LAYERS), "tensor-core path was not taken"


PAIR_CASES = [("s0/um_comb/c2", 40), ("s0/um_full2", 40), ("s0/um_res2/c2", 40), ("s0/um_res1/skip", 40), ("s0/um_comb/c3", 3), ("s0/um_out", 7),
              ("s0/hg/n2/lower3/c1", 5), ("s0/hg/n3/upper1/c2", 3)]


def test_pair_kernel_bit_identical_to_one_cta_kernel(built_lib):
    """conv_tc_pair_kernel (tcgen05 cta_group::2, 256-pixel tiles) must reproduce conv_tc_kernel bit for bit: same k order, same three
    products per k-step, same epilogue.  Covers big layers at the bench batch, odd tile counts (phantom peer tile), tiles that span
    several images (4x4 maps), ragged Cin (160) and a narrow Cout (48).  The one-CTA kernel itself is pinned to the oracle above."""
    from densereg_b200.engine import DenseRegEngine
    eng = DenseRegEngine(2, 128, 16, max_batch=40, precision="tf32x3", training=False, tc_pair=False)
    eng.init_params(0, 0.05)
    L = eng.layers(); names = [l["name"] for l in L]
    for name, B in PAIR_CASES:
        li = names.index(name); l = L[li]
        g = torch.Generator(device="cuda").manual_seed(li)
        x = torch.randn(B, l["in_hw"], l["in_hw"], l["cin"], device="cuda", generator=g)
        y0 = eng.debug_conv(li, x, "tf32x3")                                   # handle has the pair path off -> one-CTA kernel
        y1 = eng.debug_conv(li, x, "tf32x3", reuse_weights=True, pair=True)    # forced CTA-pair kernel
        torch.cuda.synchronize()
        assert bool(torch.isfinite(y1).all()), name
        assert torch.equal(y0, y1), "%s B=%d: max |diff| %.3e" % (name, B, float((y0 - y1).abs().max()))


def test_pair_default_training_step_matches_one_cta(built_lib):
    """Whole micro-batch (B=40 so that the big layers take the pair kernel) with the pair path on (default) vs off: same loss; the conv kernels
    are bit-identical (test above) but the fused BRN statistics are summed per CTA in fp32 before the double atomics, and the two kernels tile
    the pixels differently, so the statistics differ in the last bit and a few ReLU / BRN decisions at |z| ~ 1e-7 flip: gradients agree at the
    fp32 noise floor of the graph (tests/test_gpu_net.py), measured 6e-4 of the largest gradient."""
    from densereg_b200.engine import DenseRegEngine
    from densereg_b200 import synth
    B, J = 40, 16
    d, po, cf, co = [torch.from_numpy(a).cuda() for a in synth.make_batch(B, J, seed=0)]
    out = {}
    for pair in (False, True):
        eng = DenseRegEngine(2, 128, J, max_batch=B, precision="tf32x3", training=True, tc_pair=pair)
        eng.init_params(0)
        eng.zero_grads()
        loss = eng.loss_backward(d, po, cf, co, dropout_seed=1).clone()
        torch.cuda.synchronize()
        out[pair] = (loss.cpu(), eng.grads.clone().cpu())
        eng.close(); del eng
    la, ga = out[False]; lb, gb = out[True]
    assert float(((la - lb).abs() / la.abs().clamp_min(1e-30)).max()) < 1e-6
    assert float((ga - gb).abs().max()) <= 5e-3 * float(ga.abs().max())
    assert float((ga - gb).norm() / ga.norm()) < 1e-2
