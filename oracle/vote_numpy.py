"""Oracle (test infrastructure): NumPy fp32 restatement of the offset-vote post-process.

PARITY UNPINNED (see oracle/__init__.py).  Every op is done as a separate fp32 NumPy op in
the order the reference's TF graph applies them, so the fp32 roundings (and therefore the
top-5 index lists and the truncated re-projection indices) are reproducible by the CUDA
kernel, which is compiled with --fmad=false for the same reason.

Reference functions restated (all in /root/reference):
  data/preprocess.py:176-187   norm_dm
  data/preprocess.py:189-232   generate_xyzs_from_multi_cfgs
  data/preprocess.py:144-170   norm_xyz_pose / unnorm_xyz_pose
  data/util.py:20              _pro (perspective projection)
  model/hourglass_um_crop_tiny.py:276-299  _resume_om
  model/hourglass_um_crop_tiny.py:598-627  _generate_candidates (tf.nn.top_k, sorted, ties -> lower index)
  model/hourglass_um_crop_tiny.py:629-682  _get_candidate_weights
  model/hourglass_um_crop_tiny.py:684-741  _weighted_mean_shift
  model/hourglass_um_crop_tiny.py:743-785  _xyz_estimation
"""
import numpy as np

f32 = np.float32
D_RANGE = f32(300.0)          # data/preprocess.py:172
POSE_NORM_RATIO = f32(100.0)  # data/preprocess.py:173
MAX_DIST_3D = f32(0.8)        # hourglass_um_crop_tiny.py:194
NUM_PT = 5                    # hourglass_um_crop_tiny.py:770
NUM_IT = 10                   # hourglass_um_crop_tiny.py:775
BAND_WIDTH = 0.4              # hourglass_um_crop_tiny.py:775


def norm_dm(dms, coms):
    """data/preprocess.py:176-187.  dms (B,H,W) mm, coms (B,3) -> normalised (B,H,W) fp32."""
    dms = np.asarray(dms, f32)
    coms = np.asarray(coms, f32)
    cz = coms[:, 2].reshape(-1, 1, 1)
    max_depth = cz + D_RANGE * f32(0.5)
    min_depth = cz - D_RANGE * f32(0.5)
    mask = (dms < max_depth) & (dms > (min_depth - D_RANGE * f32(0.5)))
    normed = (dms - min_depth) / D_RANGE
    return np.where(mask, normed, f32(-1.0)).astype(f32)


def tiny_dm(normed_dms, out_hw=32):
    """tf.image.resize_images(.., method=2) at integer scale == strided subsample
    (um_v1.py:111, hourglass_um_crop_tiny.py:340,453)."""
    s = normed_dms.shape[1] // out_hw
    return np.ascontiguousarray(normed_dms[:, ::s, ::s])


def scaled_cfg(cfg, w, h):
    """CameraConfig rescale used at preprocess.py:212-216 and hourglass...:646-650 (fp32)."""
    cfg = np.asarray(cfg, f32)
    w_ratio = cfg[..., 4] / f32(w)
    h_ratio = cfg[..., 5] / f32(h)
    return (cfg[..., 0] / w_ratio, cfg[..., 1] / h_ratio,
            cfg[..., 2] / w_ratio, cfg[..., 3] / h_ratio)


def generate_xyzs(dms, cfgs, coms):
    """data/preprocess.py:189-232.  dms (B,H,W) normalised -> (B,H,W,3) normalised points."""
    dms = np.asarray(dms, f32)
    cfgs = np.asarray(cfgs, f32)
    coms = np.asarray(coms, f32)
    B, H, W = dms.shape
    cz = coms[:, 2].reshape(-1, 1, 1)
    min_depth = cz - D_RANGE * f32(0.5)
    max_depth = cz + D_RANGE * f32(0.5)
    zz = np.where(dms < f32(-0.99), np.ones_like(dms) * max_depth, dms * D_RANGE + min_depth).astype(f32)
    # tf.meshgrid(range(h), range(w)) 'xy': xx[i,j]=j (column), yy[i,j]=i (row)
    xx = np.broadcast_to(np.arange(W, dtype=f32).reshape(1, 1, W), (B, H, W))
    yy = np.broadcast_to(np.arange(H, dtype=f32).reshape(1, H, 1), (B, H, W))
    fx, fy, cx, cy = [v.reshape(-1, 1, 1) for v in scaled_cfg(cfgs, W, H)]
    xx = ((xx - cx) * (zz / fx)).astype(f32)
    yy = ((yy - cy) * (zz / fy)).astype(f32)
    xx = (xx - coms[:, 0].reshape(-1, 1, 1)) / POSE_NORM_RATIO
    yy = (yy - coms[:, 1].reshape(-1, 1, 1)) / POSE_NORM_RATIO
    zz = (zz - cz) / POSE_NORM_RATIO
    return np.stack([xx, yy, zz], axis=-1).astype(f32)


def resume_om(hm3, um):
    """hourglass_um_crop_tiny.py:276-299.  om = um * (0.8 - hm3*0.8) per joint."""
    hm3 = np.asarray(hm3, f32)
    um = np.asarray(um, f32)
    d = MAX_DIST_3D - hm3 * MAX_DIST_3D                    # (B,H,W,J)
    return (um * np.repeat(d, 3, axis=-1)).astype(f32)      # (B,H,W,3J)


def top_k_sorted(v, k):
    """tf.nn.top_k(sorted=True): descending values, equal values -> lower index first."""
    order = np.lexsort((np.arange(v.shape[0]), -v.astype(np.float64)))
    return order[:k].astype(np.int32)


def xyz_estimation(hm, hm3, um, dms, cfgs, coms, return_aux=False):
    """hourglass_um_crop_tiny.py:743-785 followed by unnorm_xyz_pose (preprocess.py:157-170).

    hm, hm3 (B,H,W,J), um (B,H,W,3J): raw last-stack outputs; dms (B,H,W) normalised depth at
    H x W; cfgs (B,6) crop intrinsics; coms (B,3) mm.
    Returns xyz_mm (B,3J) fp32 and top5 (B,J,5) int32 [and aux dict].
    Out-of-range re-projection indices (tf.gather_nd raises on TF-CPU) are CLAMPED to the map
    and counted -- a documented deviation shared with the CUDA kernel.
    """
    hm = np.asarray(hm, f32); hm3 = np.asarray(hm3, f32); um = np.asarray(um, f32)
    dms = np.asarray(dms, f32); cfgs = np.asarray(cfgs, f32); coms = np.asarray(coms, f32)
    B, H, W, J = hm.shape
    om = resume_om(hm3, um)
    P = generate_xyzs(dms, cfgs, coms)                                  # (B,H,W,3)
    votes = (np.tile(P, (1, 1, 1, J)) + om).astype(f32)                 # :756-760
    refined = ((hm + f32(1.0)) * hm3).astype(f32)                       # :764
    mask = np.where(dms < f32(-0.99), f32(0.0), f32(1.0)).astype(f32)   # :767
    refined = (refined * mask[..., None]).astype(f32)                   # :768
    fx, fy, cx, cy = scaled_cfg(cfgs, W, H)
    inv_sigma = f32(-1.0 / (2 * BAND_WIDTH * BAND_WIDTH))               # :732
    out = np.zeros((B, J, 3), f32)
    top5 = np.zeros((B, J, NUM_PT), np.int32)
    clamped = 0
    wts = np.zeros((B, J, NUM_PT), f32)
    for b in range(B):
        R = refined[b].reshape(H * W, J)
        V = votes[b].reshape(H * W, 3 * J)
        for j in range(J):
            idx = top_k_sorted(R[:, j], NUM_PT)                         # :617
            top5[b, j] = idx
            can = V[idx, 3 * j:3 * j + 3].astype(f32)                   # (5,3) :618-621
            # _get_candidate_weights :640-664
            q = (can * POSE_NORM_RATIO + coms[b]).astype(f32)
            u = (q[:, 0] * fx[b]) / q[:, 2] + cx[b]
            v = (q[:, 1] * fy[b]) / q[:, 2] + cy[b]
            with np.errstate(invalid='ignore'):
                uu = np.nan_to_num((u + f32(0.5)).astype(f32), nan=0.0, posinf=1e9, neginf=-1e9)
                vv = np.nan_to_num((v + f32(0.5)).astype(f32), nan=0.0, posinf=1e9, neginf=-1e9)
            uu = np.clip(np.trunc(uu), -2**31, 2**31 - 1).astype(np.int64)   # tf.to_int32 truncates
            vv = np.clip(np.trunc(vv), -2**31, 2**31 - 1).astype(np.int64)
            oob = (uu < 0) | (uu >= W) | (vv < 0) | (vv >= H)
            clamped += int(oob.sum())
            uu = np.clip(uu, 0, W - 1); vv = np.clip(vv, 0, H - 1)
            w = hm[b, vv, uu, j].astype(f32)                             # raw hm, may be < 0
            wts[b, j] = w
            # _weighted_mean_shift :694-724
            quan = np.clip((can + f32(1.0)) * f32(2.0), f32(0.0), f32(2 * 2.0 - 0.1)).astype(f32)
            quan = np.trunc(quan).astype(np.int64)
            hist = np.zeros((4, 4, 4), f32)
            for k in range(NUM_PT):                                      # scatter_nd sums duplicates in order
                hist[quan[k, 0], quan[k, 1], quan[k, 2]] += w[k]
            cells = np.argwhere(hist == hist.max())                      # row-major; [-1] = last
            cur = cells[-1].astype(f32) / f32(2.0) - f32(1.0)
            cur = (cur + f32(0.5 / 2.0)).astype(f32)
            with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
                for _ in range(NUM_IT):
                    diff = (can - cur).astype(f32)
                    s = ((diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]).astype(f32)
                         + diff[:, 2] * diff[:, 2]).astype(f32)         # reduce_sum over 3, left to right
                    s = np.exp((inv_sigma * s).astype(f32)).astype(f32)
                    s = (s * w).astype(f32)
                    prod = (can * s[:, None]).astype(f32)
                    num = prod[0]
                    for k in range(1, NUM_PT):
                        num = (num + prod[k]).astype(f32)
                    den = s[0]
                    for k in range(1, NUM_PT):
                        den = f32(den + s[k])
                    cur = (num / den).astype(f32)
            out[b, j] = cur
    xyz = (out * POSE_NORM_RATIO + coms[:, None, :]).astype(f32)         # unnorm_xyz_pose
    xyz = xyz.reshape(B, 3 * J)
    if return_aux:
        return xyz, top5, dict(clamped=clamped, weights=wts, votes=votes, refined=refined)
    return xyz, top5


def xyz_estimation_f64(hm, hm3, um, dms, cfgs, coms, top5):
    """float64 cross-check of the floating-point part given the fp32-exact top-5 lists.
    Used only to bound the fp32 oracle's own rounding error in tests."""
    hm = np.asarray(hm, np.float64); hm3 = np.asarray(hm3, np.float64); um = np.asarray(um, np.float64)
    dms = np.asarray(dms, np.float64); cfgs = np.asarray(cfgs, np.float64); coms = np.asarray(coms, np.float64)
    B, H, W, J = hm.shape
    out = np.zeros((B, J, 3))
    for b in range(B):
        r = cfgs[b, 4] / W
        fx, fy, cx, cy = cfgs[b, 0] / r, cfgs[b, 1] / r, cfgs[b, 2] / r, cfgs[b, 3] / r
        for j in range(J):
            can = np.zeros((5, 3)); w = np.zeros(5)
            for k, p in enumerate(top5[b, j]):
                i, jj = divmod(int(p), W)
                d = dms[b, i, jj]
                z = coms[b, 2] + 150.0 if d < -0.99 else d * 300.0 + coms[b, 2] - 150.0
                x = (jj - cx) * z / fx; y = (i - cy) * z / fy
                Pn = np.array([(x - coms[b, 0]) / 100, (y - coms[b, 1]) / 100, (z - coms[b, 2]) / 100])
                dd = 0.8 - 0.8 * hm3[b, i, jj, j]
                can[k] = Pn + um[b, i, jj, 3 * j:3 * j + 3] * dd
                q = can[k] * 100 + coms[b]
                uu = int(np.clip(np.trunc(q[0] * fx / q[2] + cx + 0.5), 0, W - 1))
                vv = int(np.clip(np.trunc(q[1] * fy / q[2] + cy + 0.5), 0, H - 1))
                w[k] = hm[b, vv, uu, j]
            quan = np.trunc(np.clip((can + 1) * 2, 0, 3.9)).astype(int)
            hist = np.zeros((4, 4, 4))
            for k in range(5):
                hist[tuple(quan[k])] += w[k]
            cur = np.argwhere(hist == hist.max())[-1] / 2.0 - 1.0 + 0.25
            for _ in range(NUM_IT):
                s = w * np.exp(-((can - cur) ** 2).sum(-1) / (2 * 0.16))
                cur = (can * s[:, None]).sum(0) / s.sum()
            out[b, j] = cur * 100 + coms[b]
    return out.reshape(B, 3 * J)
