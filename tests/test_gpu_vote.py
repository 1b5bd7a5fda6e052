"""GPU parity: offset-vote kernel (through the C-ABI dr_vote) vs the CPU oracle.
Bar: top-5 index lists bit-exact, xyz within 1e-3 mm (BASELINE.json north_star)."""
import os
import numpy as np
import pytest
import torch
from gpu_util import cu, dump

pytestmark = pytest.mark.gpu
TOL_MM = 1e-3
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def eng(built_lib):
    from densereg_b200.engine import DenseRegEngine
    return DenseRegEngine(num_stack=1, num_fea=64, num_jnt=16, max_batch=2, training=False)


def run_vote(eng, hm, hm3, um, dmn, cfgs, coms):
    xyz, top5, clamp = eng.vote(cu(hm), cu(hm3), cu(um), cu(dmn), cu(cfgs), cu(coms), return_top5=True)
    torch.cuda.synchronize()
    return xyz.cpu().numpy(), top5.cpu().numpy(), int(clamp.item())


@pytest.mark.parametrize("name", ["vote_J16", "vote_J14", "vote_J21"])
def test_vote_matches_committed_golden(eng, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    xyz, top5, _ = run_vote(eng, g["hm"], g["hm3"], g["um"], g["dmn"], g["cfgs"], g["coms"])
    assert np.array_equal(top5, g["top5"])
    assert np.abs(xyz - g["xyz"]).max() <= TOL_MM


@pytest.mark.parametrize("J,hw,B,seed", [(16, 32, 4, 0), (14, 32, 5, 1), (21, 32, 3, 2), (21, 64, 2, 3), (21, 128, 1, 4),
                                         (1, 32, 2, 5), (16, 8, 2, 6), (33, 16, 2, 7)])
def test_vote_matches_oracle(eng, J, hw, B, seed):
    from oracle import vote_numpy as V
    from densereg_b200 import synth
    hm, hm3, um, dmn, cfgs, coms = synth.make_vote_maps(B, J, hw=hw, seed=seed)
    ref_xyz, ref_top5, aux = V.xyz_estimation(hm, hm3, um, dmn, cfgs, coms, return_aux=True)
    xyz, top5, clamp = run_vote(eng, hm, hm3, um, dmn, cfgs, coms)
    assert np.array_equal(top5, ref_top5), "top-5 index lists differ"
    assert clamp == aux["clamped"]
    ok = np.isfinite(ref_xyz)
    err = np.abs(xyz - ref_xyz)[ok].max()
    dump("vote_err_J%d_hw%d.json" % (J, hw), dict(max_err_mm=float(err), clamped=clamp))
    assert err <= TOL_MM
    assert np.array_equal(np.isnan(xyz), np.isnan(ref_xyz))


def test_vote_ties_and_masked_background(eng):
    """Quantised scores force many exact ties (tf.nn.top_k: lower index wins) and an almost fully masked depth map
    forces zero-score candidates to be selected."""
    from oracle import vote_numpy as V
    from densereg_b200 import synth
    hm, hm3, um, dmn, cfgs, coms = synth.make_vote_maps(3, 16, hw=32, seed=21)
    hm = np.round(hm * 4) / 4
    hm3 = np.round(hm3 * 4) / 4
    dmn[1] = -1.0
    dmn[1, 5, 7] = 0.3
    ref_xyz, ref_top5 = V.xyz_estimation(hm, hm3, um, dmn, cfgs, coms)
    xyz, top5, _ = run_vote(eng, hm, hm3, um, dmn, cfgs, coms)
    assert np.array_equal(top5, ref_top5)
    ok = np.isfinite(ref_xyz)
    assert np.abs(xyz - ref_xyz)[ok].max() <= TOL_MM
    assert np.array_equal(np.isfinite(xyz), ok)


def test_vote_roundtrip_property_full_size(eng):
    """Size-independent property at the microbench shape: perfect maps synthesised from a known pose with the
    reference's own GT synthesis must vote back to that pose (hourglass_um_crop_tiny.py:336-346 -> :457-462)."""
    from oracle import um_v1_torch as U, vote_numpy as V
    from densereg_b200 import synth
    B, J = 64, 21
    dms, poses, cfgs, coms = synth.make_batch(B, J, seed=33)
    x0, gt_hm, gt_hm3, gt_um = U.gt_maps(dms[..., 0], poses, cfgs, coms)
    d32 = V.tiny_dm(x0[..., 0].numpy())
    xyz, top5, clamp = run_vote(eng, gt_hm.numpy(), gt_hm3.numpy(), gt_um.numpy(), d32, cfgs, coms)
    assert np.abs(xyz - poses).max() <= TOL_MM
