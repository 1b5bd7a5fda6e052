"""CPU tests of the network oracle (oracle/um_v1_torch.py): shape table == SURVEY.md A.1/A.2, SAME padding
cases, BRN train/eval consistency, layer table of the CUDA library == oracle table, golden statistics."""
import ctypes as C
import os
import numpy as np
import pytest
import torch
from oracle import um_v1_torch as U
from oracle import vote_numpy as V
from densereg_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("S,F,J,convs,brn,params", [
    (1, 64, 16, 77, 72, 2204512), (2, 128, 16, 146, 134, 5856352),
    (2, 128, 14, 146, 134, 5827804), (2, 128, 21, 146, 134, 5929262)])
def test_layer_table_counts(S, F, J, convs, brn, params):
    net = U.Net(S, F, J)
    assert len(net.specs) == convs and sum(c.brn for c in net.specs) == brn and net.n_params == params
    macs = 0
    for c in net.specs:
        hw = 64 if (c.name.startswith("stem/conv_1") or c.name.startswith("stem/conv_2")) else None
        if hw is None:
            hw = 32
            if "/hg/n" in c.name:
                n = int(c.name.split("/hg/n")[1][0])
                hw = 32 >> (4 - n) if "upper1" in c.name else 32 >> (5 - n)
        macs += hw * hw * c.k * c.k * c.cin * c.cout
    expect = {(1, 64, 16): 2.1310e9, (2, 128, 16): 4.8950e9, (2, 128, 14): 4.8658e9, (2, 128, 21): 4.9695e9}[(S, F, J)]
    assert abs(macs - expect) / expect < 2e-4


def test_same_padding_cases():
    assert U.same_pad(128, 7, 2) == (2, 3)
    assert U.same_pad(32, 3, 1) == (1, 1)
    for n in (32, 16, 8, 4):
        assert U.same_pad(n, 3, 2) == (0, 1)
    assert U.same_pad(64, 2, 2) == (0, 0) and U.same_pad(32, 1, 1) == (0, 0)


def test_forward_shapes_and_brn_modes():
    net = U.Net(1, 64, 16)
    p, s = net.init_params(0), net.init_state()
    dms, poses, cfgs, coms = synth.make_batch(2, 16, seed=0)
    x0 = torch.from_numpy(V.norm_dm(dms[..., 0], coms)[..., None])
    hms, hm3s, ums = net.forward(p, s, x0, training=False)
    assert hms[0].shape == (2, 32, 32, 16) and hm3s[0].shape == (2, 32, 32, 16) and ums[0].shape == (2, 32, 32, 48)
    s2 = s.clone()
    net.forward(p, s2, x0, training=True, dropout_seed=1)
    net.apply_state_updates()
    c = net.specs[0]
    # zero-debiased EMA: after ONE update the moving stats equal the batch stats
    raw = net.trace[c.name + ":raw"]
    np.testing.assert_allclose(s2[c.s_off:c.s_off + c.cout].numpy(), raw.mean(dim=(0, 2, 3)).detach().numpy(), rtol=1e-4, atol=1e-7)
    assert abs(s2[c.s_off + 4 * c.cout].item() - 1.0) < 1e-6 and abs(s2[c.s_off + 4 * c.cout + 1].item() - 1e-3) < 1e-9


def test_dropout_hash_is_half_and_deterministic():
    m = U.dropout_mask(7, 3, 1 << 16)
    assert 0.49 < m.mean() < 0.51
    assert np.array_equal(m, U.dropout_mask(7, 3, 1 << 16))
    assert not np.array_equal(m, U.dropout_mask(8, 3, 1 << 16))


def test_cuda_library_layer_table_matches_oracle(built_lib):
    from densereg_b200 import _ffi
    lib = built_lib
    for (S, F, J) in [(2, 128, 16), (1, 64, 16), (2, 128, 14), (2, 128, 21)]:
        cfg = _ffi.DrConfig(num_stack=S, num_fea=F, kernel_size=3, num_jnt=J, in_hw=128, out_hw=32, max_batch=4,
                            precision=0, device=0)
        h = C.c_void_p()
        assert lib.dr_create(C.byref(h), C.byref(cfg)) == 0
        net = U.Net(S, F, J)
        assert lib.dr_param_count(h) == net.n_params and lib.dr_state_count(h) == net.n_state
        assert lib.dr_num_layers(h) == len(net.specs)
        for i, c in enumerate(net.specs):
            li = _ffi.DrLayerInfo()
            assert lib.dr_get_layer(h, i, C.byref(li)) == 0
            assert (li.name.decode(), li.k, li.stride, li.cin, li.cout, li.brn, li.relu, li.w_off, li.p_off) == \
                   (c.name, c.k, c.stride, c.cin, c.cout, int(c.brn), int(c.relu), c.w_off, c.p_off)
            assert abs(li.wd - c.wd) < 1e-9
        lib.dr_destroy(h)


def test_library_exports_every_declared_symbol(built_lib):
    import re
    from densereg_b200 import _ffi
    hdr = open(os.path.join(os.path.dirname(GOLD), "..", "include", "densereg.h")).read()
    declared = set(re.findall(r"DR_API [a-z_0-9 \*]+?(dr_[a-z_0-9]+)\(", hdr))
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    for name in declared:
        assert hasattr(built_lib, name)
    assert built_lib.dr_version() == 100


def test_net_golden_statistics():
    g = np.load(os.path.join(GOLD, "net_S1F64J16.npz"))
    net = U.Net(1, 64, 16)
    p, s = net.init_params(int(g["seed"]), stddev=float(g["stddev"])), net.init_state()
    dms, poses, cfgs, coms = synth.make_batch(1, 16, seed=int(g["data_seed"]))
    x0 = torch.from_numpy(V.norm_dm(dms[..., 0], coms)[..., None])
    hms, hm3s, ums = net.forward(p, s, x0, training=False)
    np.testing.assert_allclose(hms[0].numpy()[0, ::4, ::4], g["hm_sub"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(ums[0].numpy()[0, ::4, ::4], g["um_sub"], rtol=2e-4, atol=1e-6)


def test_c_abi_error_behaviour_without_a_gpu(built_lib):
    """Argument and call-order checks of the C-ABI happen before any CUDA call, so they are testable here: bad arguments -> DR_ERR_ARG (-1),
    wrong call order -> DR_ERR_STATE (-3) with a message from dr_last_error; nothing throws across the boundary."""
    from densereg_b200 import _ffi
    lib = built_lib
    h = C.c_void_p()
    ok = dict(num_stack=2, num_fea=128, kernel_size=3, num_jnt=16, in_hw=128, out_hw=32, max_batch=4, precision=0, device=0)
    for bad in (dict(kernel_size=5), dict(in_hw=256), dict(out_hw=64), dict(num_stack=0), dict(num_stack=5), dict(num_fea=6), dict(num_fea=130),
                dict(num_jnt=0), dict(num_jnt=65), dict(max_batch=0)):
        cfg = _ffi.DrConfig(**dict(ok, **bad))
        assert lib.dr_create(C.byref(h), C.byref(cfg)) == -1, bad
        assert not h.value
    assert lib.dr_create(None, None) == -1
    cfg = _ffi.DrConfig(**ok)
    assert lib.dr_create(C.byref(h), C.byref(cfg)) == 0 and h.value
    li = _ffi.DrLayerInfo()
    assert lib.dr_get_layer(h, -1, C.byref(li)) == -1 and lib.dr_get_layer(h, lib.dr_num_layers(h), C.byref(li)) == -1
    assert lib.dr_get_layer(h, 0, None) == -1
    assert lib.dr_bind(h, None, None, None, None, None) == -1                      # params and state are mandatory
    fake = C.c_void_p(0x1000)                                                      # never dereferenced: every call below fails its checks first
    # call order: nothing bound yet
    assert lib.dr_init_params(h, 0, C.c_float(0.01), None) == -3 and b"dr_bind" in lib.dr_last_error(h)
    assert lib.dr_debug_conv(h, 0, 1, fake, fake, 0, None) == -3
    # no gradient / Adam buffers bound: the training entry points refuse
    assert lib.dr_loss_backward(h, 1, fake, fake, fake, fake, fake, 0, 1, None) == -3 and b"grads" in lib.dr_last_error(h)
    assert lib.dr_zero_grads(h, None) == -3
    assert lib.dr_optimizer_step(h, 5, 1, C.c_float(1e-3), 1, None) == -3
    # argument checks
    assert lib.dr_optimizer_step(h, 0, 1, C.c_float(1e-3), 1, None) == -1 and lib.dr_optimizer_step(h, 5, 0, C.c_float(1e-3), 1, None) == -1
    assert lib.dr_infer(h, 1, None, fake, fake, fake, None, None) == -1
    assert lib.dr_loss_backward(h, 1, None, fake, fake, fake, fake, 0, 1, None) == -1
    assert lib.dr_vote(h, 1, 32, 32, 65, fake, fake, fake, fake, fake, fake, fake, None, None, None) == -1             # J > 64
    assert lib.dr_vote(h, 1, 2, 2, 16, fake, fake, fake, fake, fake, fake, fake, None, None, None) == -1               # fewer than 5 pixels: no top-5
    assert lib.dr_norm_dm(h, 0, 128, fake, fake, fake, None) == -1
    assert lib.dr_debug_conv(h, 10 ** 6, 1, fake, fake, 0, None) == -1
    assert lib.dr_data_aug(h, 1, 128, 16, fake, fake, fake, fake, fake, fake, fake, None, None) == -1                   # null output
    # binding needs a device: on a GPU-less machine it reports a CUDA error code instead of crashing
    import torch
    rc = lib.dr_bind(h, fake, fake, None, None, None)
    assert rc == 0 if torch.cuda.is_available() else rc in (0, -2)
    assert lib.dr_destroy(h) == 0


def test_pipeline_c_abi_without_a_gpu(built_lib):
    """Micro-batch pipeline (include/densereg.h, dr_pipeline_join): the second arena is a host-side object until dr_bind, so creation, the
    depth query, the layer table and destruction are checkable here; the training entry points still refuse without bound gradients."""
    from densereg_b200 import _ffi
    lib = built_lib
    ok = dict(num_stack=1, num_fea=64, kernel_size=3, num_jnt=16, in_hw=128, out_hw=32, max_batch=4, precision=2, device=0)
    assert lib.dr_pipeline_depth(None) == 0 and lib.dr_pipeline_join(None, None) == -1
    for want, depth in ((0, 1), (2, 2), (1, 1)):
        cfg = _ffi.DrConfig(**ok)
        cfg.reserved[2] = want
        h = C.c_void_p()
        assert lib.dr_create(C.byref(h), C.byref(cfg)) == 0 and h.value
        assert lib.dr_pipeline_depth(h) == depth
        assert lib.dr_pipeline_join(h, None) == 0                      # nothing in flight: a no-op, no CUDA call
        assert lib.dr_launch_count(h) == 0 and lib.dr_tc_launch_count(h) == 0
        n = lib.dr_num_layers(h)
        assert n > 0 and lib.dr_param_count(h) > 0                     # one parameter set whatever the depth
        fake = C.c_void_p(0x1000)
        assert lib.dr_loss_backward(h, 1, fake, fake, fake, fake, fake, 0, 1, None) == -3 and b"grads" in lib.dr_last_error(h)
        assert lib.dr_zero_grads(h, None) == -3
        assert lib.dr_destroy(h) == 0


def test_import_sets_hardware_queue_default():
    """densereg_b200/__init__.py: CUDA_DEVICE_MAX_CONNECTIONS defaults to 32 at import (before the CUDA context exists) and a user's value wins."""
    import subprocess, sys
    code = "import os; import densereg_b200; print(os.environ['CUDA_DEVICE_MAX_CONNECTIONS'])"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k != "CUDA_DEVICE_MAX_CONNECTIONS"}
    assert subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True).stdout.strip() == "32"
    env["CUDA_DEVICE_MAX_CONNECTIONS"] = "4"
    assert subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True).stdout.strip() == "4"
