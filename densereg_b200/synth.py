"""Seeded synthetic depth crops / poses / camera configs.

There are no datasets on the box (ICVL/NYU/MSRA need network access), so every config in
BASELINE.json runs on synthetic 128x128x1 crops of the reference's shape: a smooth hand-like
blob at depth com_z +- 60 mm covering 25-45 % of the crop, background 0 (-> -1 after norm_dm,
data/preprocess.py:176-187), crop intrinsics cfg=[fx,fy,cx,cy,128,128] in the range the
reference's crop_from_xyz_pose produces (data/preprocess.py:70-78), and J joints on the blob
surface +-30 mm.  Pure NumPy, deterministic per seed.
"""
import numpy as np
from scipy.ndimage import gaussian_filter

DATASET_JOINTS = {"icvl": 16, "nyu": 14, "msra": 21}   # data/icvl.py:17, nyu.py:40-45, msra.py:17


def make_batch(batch, num_jnt, seed=0, hw=128):
    rng = np.random.RandomState(seed)
    dms = np.zeros((batch, hw, hw, 1), np.float32)
    poses = np.zeros((batch, 3 * num_jnt), np.float32)
    cfgs = np.zeros((batch, 6), np.float32)
    coms = np.zeros((batch, 3), np.float32)
    yy, xx = np.mgrid[0:hw, 0:hw].astype(np.float32)
    for b in range(batch):
        com_z = rng.uniform(250.0, 800.0)
        f = rng.uniform(150.0, 450.0)
        cx, cy = rng.uniform(48.0, 80.0, size=2)
        # the crop centre (hw/2) back-projects to the centre of mass
        com = np.array([(hw / 2 - cx) * com_z / f, (hw / 2 - cy) * com_z / f, com_z], np.float32)
        mask = np.zeros((hw, hw), bool)
        target = rng.uniform(0.25, 0.45)
        while mask.mean() < target:
            ex, ey = rng.uniform(30, 98, size=2)
            ra, rb = rng.uniform(8, 30, size=2)
            th = rng.uniform(0, np.pi)
            dx, dy = xx - ex, yy - ey
            u = dx * np.cos(th) + dy * np.sin(th)
            v = -dx * np.sin(th) + dy * np.cos(th)
            mask |= (u / ra) ** 2 + (v / rb) ** 2 <= 1.0
        relief = gaussian_filter(rng.randn(hw, hw).astype(np.float32), 6.0)
        relief = relief / (np.abs(relief).max() + 1e-6)
        dm = np.where(mask, com_z + 60.0 * relief, 0.0).astype(np.float32)
        ys, xs = np.nonzero(mask)
        pick = rng.randint(0, len(ys), size=num_jnt)
        z = dm[ys[pick], xs[pick]] + rng.uniform(-30, 30, size=num_jnt)
        px = (xs[pick] - cx) * z / f
        py = (ys[pick] - cy) * z / f
        dms[b, :, :, 0] = dm
        poses[b] = np.stack([px, py, z], axis=1).reshape(-1).astype(np.float32)
        cfgs[b] = [f, f, cx, cy, hw, hw]
        coms[b] = com
    return dms, poses, cfgs, coms


def make_vote_maps(batch, num_jnt, hw=32, seed=0):
    """Vote-microbench inputs: hm ~ cone peaks + N(0,0.05), hm3 ~ clip(U,0,1),
    um ~ unit vectors * N(1,0.1), dm_norm ~ {-1 w.p. 0.6, U(-0.4,0.9)}; cfg/com as make_batch."""
    rng = np.random.RandomState(seed)
    J = num_jnt
    yy, xx = np.mgrid[0:hw, 0:hw].astype(np.float32)
    pu = rng.uniform(4, hw - 4, size=(batch, 1, 1, J)).astype(np.float32)
    pv = rng.uniform(4, hw - 4, size=(batch, 1, 1, J)).astype(np.float32)
    dist = np.sqrt((xx[None, :, :, None] - pu) ** 2 + (yy[None, :, :, None] - pv) ** 2)
    r = 4.0 * hw / 32.0
    hm = (np.maximum(r - dist, 0) / r + 0.05 * rng.randn(batch, hw, hw, J)).astype(np.float32)
    hm3 = np.clip(rng.uniform(-0.2, 1.0, size=(batch, hw, hw, J)), 0, 1).astype(np.float32)
    um = rng.randn(batch, hw, hw, J, 3).astype(np.float32)
    um /= np.linalg.norm(um, axis=-1, keepdims=True) + 1e-6
    um *= (1.0 + 0.1 * rng.randn(batch, hw, hw, J, 1)).astype(np.float32)
    um = um.reshape(batch, hw, hw, 3 * J).astype(np.float32)
    bg = rng.uniform(size=(batch, hw, hw)) < 0.6
    dmn = np.where(bg, -1.0, rng.uniform(-0.4, 0.9, size=(batch, hw, hw))).astype(np.float32)
    cfgs = np.zeros((batch, 6), np.float32); coms = np.zeros((batch, 3), np.float32)
    for b in range(batch):
        com_z = rng.uniform(250.0, 800.0); f = rng.uniform(150.0, 450.0)
        cx, cy = rng.uniform(48.0, 80.0, size=2)
        cfgs[b] = [f, f, cx, cy, 128, 128]
        coms[b] = [(64 - cx) * com_z / f, (64 - cy) * com_z / f, com_z]
    return hm, hm3, um, dmn, cfgs, coms


# full-frame camera intrinsics of the three datasets [fx, fy, cx, cy, w, h] (data/icvl.py:12, nyu.py:13, msra.py:13)
DATASET_CFG = {"icvl": (241.42, 241.42, 160.0, 120.0, 320.0, 240.0),
               "nyu": (588.235, 587.084, 320.0, 240.0, 640.0, 480.0),
               "msra": (241.42, 241.42, 160.0, 120.0, 320.0, 240.0)}


def make_frames(batch, num_jnt, dataset="icvl", seed=0):
    """Synthetic FULL depth frames (B,h,w) mm with a hand-like blob, a background wall, joints on the blob (xyz mm in camera
    coordinates) -- input of the crop + centre-of-mass front-end (data/preprocess.py:10-142)."""
    rng = np.random.RandomState(seed)
    fx, fy, cx, cy, w, h = DATASET_CFG[dataset]
    w, h = int(w), int(h)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    frames = np.zeros((batch, h, w), np.float32)
    poses = np.zeros((batch, 3 * num_jnt), np.float32)
    for b in range(batch):
        z0 = rng.uniform(300.0, 450.0)
        ex, ey = rng.uniform(0.3 * w, 0.7 * w), rng.uniform(0.3 * h, 0.7 * h)
        ra, rb = rng.uniform(0.08, 0.16) * w, rng.uniform(0.12, 0.22) * h
        mask = ((xx - ex) / ra) ** 2 + ((yy - ey) / rb) ** 2 <= 1.0
        relief = gaussian_filter(rng.randn(h, w).astype(np.float32), 8.0)
        relief /= (np.abs(relief).max() + 1e-6)
        dm = np.where(mask, z0 + 40.0 * relief, rng.uniform(700.0, 900.0) if rng.rand() < 0.5 else 0.0).astype(np.float32)
        ys, xs = np.nonzero(mask)
        pick = rng.randint(0, len(ys), size=num_jnt)
        z = dm[ys[pick], xs[pick]] + rng.uniform(-15, 15, size=num_jnt)
        poses[b] = np.stack([(xs[pick] - cx) * z / fx, (ys[pick] - cy) * z / fy, z], axis=1).reshape(-1)
        frames[b] = dm
    return frames, poses, np.array([fx, fy, cx, cy, w, h], np.float32)
