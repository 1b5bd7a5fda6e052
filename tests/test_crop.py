"""Crop + centre-of-mass front-end: CPU tests of the oracle (oracle/crop_numpy.py) and GPU parity of
dr_crop_from_xyz_pose / dr_crop_from_bbx through the C-ABI (bit-exact crops: the kernel is compiled --fmad=false)."""
import numpy as np
import pytest
import torch
from oracle import crop_numpy as C
from densereg_b200 import synth


@pytest.mark.parametrize("dataset,J", [("icvl", 16), ("nyu", 14), ("msra", 21)])
def test_oracle_crop_properties(dataset, J):
    frames, poses, cfg = synth.make_frames(3, J, dataset, seed=2)
    for b in range(3):
        crop, ncfg = C.crop_from_xyz_pose(frames[b], poses[b], cfg, icvl=(dataset == "icvl"))
        assert crop.shape == (128, 128) and crop.dtype == np.float32
        p = poses[b].reshape(-1, 3)
        u = p[:, 0] * ncfg[0] / p[:, 2] + ncfg[2]; v = p[:, 1] * ncfg[1] / p[:, 2] + ncfg[3]
        assert u.min() > 0 and u.max() < 128 and v.min() > 0 and v.max() < 128     # joints stay inside the crop (pad 20 px)
        assert ncfg[4] == 128 and ncfg[5] == 128
        com = C.center_of_mass(crop, ncfg)
        assert com[2] >= 200.0 and abs(com[2] - crop[crop > 0].mean()) < 1e-2
        # the hand is kept, a far wall is removed by the depth threshold
        assert 0.05 < (crop > 0).mean() < 0.95 and crop.max() < (500.0 if dataset == "icvl" else 700.0)


def test_oracle_bilinear_identity_and_halving():
    img = np.arange(64 * 64, dtype=np.float32).reshape(64, 64)
    assert np.array_equal(C._resize_bilinear(img, 64, 64), img)
    half = C._resize_bilinear(img, 32, 32)                       # legacy coordinates: src = 2*dst exactly -> pure subsample
    assert np.array_equal(half, img[::2, ::2])


def test_oracle_bbx_crop_matches_pose_crop_box():
    frames, poses, cfg = synth.make_frames(1, 14, "nyu", seed=5)
    crop_p, cfg_p = C.crop_from_xyz_pose(frames[0], poses[0], cfg)
    # rebuild the same box by hand and feed it to crop_from_bbx (data/nyu.py:109-111 supplies such boxes at test time)
    p = poses[0].reshape(-1, 3); u = p[:, 0] * cfg[0] / p[:, 2] + cfg[2]; v = p[:, 1] * cfg[1] / p[:, 2] + cfg[3]
    top = int(min(max(v.min() - 20, 0), cfg[5] - 40)); left = int(min(max(u.min() - 20, 0), cfg[4] - 40))
    bottom = int(max(min(v.max() + 20, cfg[5]), top + 39)); right = int(max(min(u.max() + 20, cfg[4]), left + 39))
    uu = np.clip(u.astype(int), 0, 639); vv = np.clip(v.astype(int), 0, 479); dd = frames[0][vv, uu]; d_th = dd[dd > 100].min() + 250
    crop_b, cfg_b = C.crop_from_bbx(frames[0], [top, left, bottom, right, d_th], cfg)
    assert np.array_equal(crop_p, crop_b) and np.allclose(cfg_p, cfg_b)


@pytest.mark.gpu
@pytest.mark.parametrize("dataset,J,B", [("icvl", 16, 5), ("nyu", 14, 3), ("msra", 21, 4)])
def test_gpu_crop_matches_oracle(built_lib, dataset, J, B):
    from densereg_b200.engine import DenseRegEngine
    eng = DenseRegEngine(1, 64, J, max_batch=1, training=False)
    frames, poses, cfg = synth.make_frames(B, J, dataset, seed=11)
    dms, cfgs, coms = eng.crop_from_xyz_pose(torch.from_numpy(frames).cuda(), torch.from_numpy(poses).cuda(), cfg, icvl=(dataset == "icvl"))
    torch.cuda.synchronize()
    dms, cfgs, coms = dms.cpu().numpy(), cfgs.cpu().numpy(), coms.cpu().numpy()
    for b in range(B):
        crop, ncfg = C.crop_from_xyz_pose(frames[b], poses[b], cfg, icvl=(dataset == "icvl"))
        assert np.array_equal(dms[b, :, :, 0], crop), "crop not bit-exact"
        np.testing.assert_allclose(cfgs[b], ncfg, rtol=1e-6)
        np.testing.assert_allclose(coms[b], C.center_of_mass(crop, ncfg), rtol=2e-6, atol=1e-4)
    # NYU-style boxes
    bbx = np.array([[30, 40, 200, 230, 650.0]] * B, np.float32)
    bbx[:, 0] += np.arange(B) * 3
    d2, c2, m2 = eng.crop_from_bbx(torch.from_numpy(frames).cuda(), torch.from_numpy(bbx).cuda(), cfg)
    for b in range(B):
        crop, ncfg = C.crop_from_bbx(frames[b], bbx[b], cfg)
        assert np.array_equal(d2[b, :, :, 0].cpu().numpy(), crop)
        np.testing.assert_allclose(c2[b].cpu().numpy(), ncfg, rtol=1e-6)


@pytest.mark.gpu
def test_gpu_frames_to_xyz_pipeline(built_lib):
    """frames -> crop/CoM -> network -> vote: the (dms, cfgs, coms) produced on the GPU feed dr_infer directly."""
    from densereg_b200.engine import DenseRegEngine
    eng = DenseRegEngine(1, 64, 16, max_batch=4, training=False)
    eng.init_params(0, 0.05)
    frames, poses, cfg = synth.make_frames(4, 16, "icvl", seed=3)
    dms, cfgs, coms = eng.crop_from_xyz_pose(torch.from_numpy(frames).cuda(), torch.from_numpy(poses).cuda(), cfg, icvl=True)
    xyz = eng.infer(dms, cfgs, coms)
    torch.cuda.synchronize()
    assert xyz.shape == (4, 48) and torch.isfinite(xyz).float().mean() > 0.9


def test_oracle_data_aug_properties():
    dms, poses, cfgs, coms = synth.make_batch(2, 16, seed=0)
    o, p = C.data_aug(dms[0, :, :, 0], poses[0], cfgs[0], coms[0], 1.0, 0.0, [1.0, 1.0])
    assert np.array_equal(o, dms[0, :, :, 0]) and np.abs(p - poses[0]).max() < 1e-4              # identity
    a = 0.7
    o, p = C.data_aug(dms[0, :, :, 0], poses[0], cfgs[0], coms[0], np.float32(np.cos(a)), np.float32(np.sin(a)), [1.08, 0.93])
    pp = p.reshape(-1, 3); u = pp[:, 0] * cfgs[0, 0] / pp[:, 2] + cfgs[0, 2]; v = pp[:, 1] * cfgs[0, 1] / pp[:, 2] + cfgs[0, 3]
    ui = np.clip(np.round(u).astype(int), 0, 127); vi = np.clip(np.round(v).astype(int), 0, 127)
    assert (o[vi, ui] > 0).mean() > 0.9          # image warp and pose warp agree: the joints still sit on the hand
    assert np.array_equal(C._crop_or_pad(np.ones((130, 120), np.float32), 128, 128).sum(), 128 * 120)


@pytest.mark.gpu
def test_gpu_data_aug_matches_oracle(built_lib):
    from densereg_b200.engine import DenseRegEngine
    B, J = 6, 16
    eng = DenseRegEngine(1, 64, J, max_batch=1, training=False)
    dms, poses, cfgs, coms = synth.make_batch(B, J, seed=4)
    rng = np.random.RandomState(0)
    ang = rng.uniform(-np.pi, np.pi, B).astype(np.float32); ang[0] = 0.0
    cs = np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32)
    er = np.clip(rng.normal(1.0, 0.2, (B, 2)), 0.9, 1.1).astype(np.float32); er[0] = 1.0; er[1] = [1.1, 0.9]; er[2] = [0.9, 1.1]
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    do, po = eng.data_aug(cu(dms), cu(poses), cu(cfgs), cu(coms), cu(cs), cu(er))
    torch.cuda.synchronize()
    do, po = do.cpu().numpy(), po.cpu().numpy()
    for b in range(B):
        o, p = C.data_aug(dms[b, :, :, 0], poses[b], cfgs[b], coms[b], cs[b, 0], cs[b, 1], er[b])
        assert np.array_equal(do[b, :, :, 0], o), "augmented crop %d not bit-exact" % b
        np.testing.assert_allclose(po[b], p, rtol=1e-5, atol=1e-3)
