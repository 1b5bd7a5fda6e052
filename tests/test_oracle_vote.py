"""CPU tests of the vote oracle (oracle/vote_numpy.py): round-trip property through the reference's own
GT synthesis, tf.nn.top_k tie rule, mean-shift seed rule, norm_dm edge cases, committed golden vectors."""
import os
import numpy as np
import pytest
from oracle import vote_numpy as V
from oracle import um_v1_torch as U
from densereg_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("J,seed", [(16, 0), (14, 1), (21, 2), (16, 3), (16, 4)])
def test_gt_maps_vote_back_to_pose(J, seed):
    # hourglass_um_crop_tiny.py:336-346 (GT synthesis) -> :457-462 (vote) must invert each other
    dms, poses, cfgs, coms = synth.make_batch(2, J, seed=seed)
    x0, gt_hm, gt_hm3, gt_um = U.gt_maps(dms[..., 0], poses, cfgs, coms)
    d32 = V.tiny_dm(x0[..., 0].numpy())
    xyz, top5, aux = V.xyz_estimation(gt_hm.numpy(), gt_hm3.numpy(), gt_um.numpy(), d32, cfgs, coms, return_aux=True)
    assert not np.isnan(xyz).any()
    assert np.abs(xyz - poses).max() < 1e-3          # mm
    xyz64 = V.xyz_estimation_f64(gt_hm.numpy(), gt_hm3.numpy(), gt_um.numpy(), d32, cfgs, coms, top5)
    assert np.abs(xyz - xyz64).max() < 1e-3          # fp32 oracle vs float64 restatement


def test_top_k_tie_rule_lower_index_first():
    v = np.array([1.0, 3.0, 3.0, 0.5, 3.0, 2.0, 2.0], np.float32)
    assert V.top_k_sorted(v, 5).tolist() == [1, 2, 4, 5, 6]
    z = np.array([0.0, -0.0, 0.0, -0.0, 0.0, 0.0], np.float32)     # -0 == +0 -> index order
    assert V.top_k_sorted(z, 5).tolist() == [0, 1, 2, 3, 4]


def test_norm_dm_edges():
    com = np.array([[0, 0, 400.0]], np.float32)
    dm = np.array([[[0.0, 99.9, 100.0, 100.1, 250.0, 400.0, 549.9, 550.0, 551.0]]], np.float32)
    out = V.norm_dm(dm, com)[0, 0]
    # valid iff 100 < d < 550 (preprocess.py:181)
    assert out[0] == -1 and out[1] == -1 and out[2] == -1 and out[7] == -1 and out[8] == -1
    np.testing.assert_allclose(out[3:7], (np.array([100.1, 250, 400, 549.9], np.float32) - 250) / 300, rtol=1e-6)
    assert out[3] < 0      # values between z_c-300 and z_c-150 normalise to [-0.5, 0)


def test_seed_rule_last_max_cell_and_degenerate_weights():
    # all weights zero -> histogram max 0 -> LAST row-major cell (3,3,3) is the seed (tf.where(...)[-1])
    B, H, W, J = 1, 8, 8, 1
    hm = np.zeros((B, H, W, J), np.float32)
    hm3 = np.random.RandomState(0).uniform(0.1, 1, (B, H, W, J)).astype(np.float32)
    um = np.zeros((B, H, W, 3 * J), np.float32)
    dmn = np.zeros((B, H, W), np.float32)
    cfg = np.array([[240, 240, 64, 64, 128, 128]], np.float32)
    com = np.array([[0, 0, 400.0]], np.float32)
    with np.errstate(all="ignore"):
        xyz, top5 = V.xyz_estimation(hm, hm3, um, dmn, cfg, com)
    assert np.isnan(xyz).all()          # 0/0 like the reference graph (degenerate case)


@pytest.mark.parametrize("name", ["vote_J16", "vote_J14", "vote_J21"])
def test_oracle_matches_committed_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    xyz, top5 = V.xyz_estimation(g["hm"], g["hm3"], g["um"], g["dmn"], g["cfgs"], g["coms"])
    assert np.array_equal(top5, g["top5"])
    np.testing.assert_allclose(xyz, g["xyz"], atol=1e-4)
