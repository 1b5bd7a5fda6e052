"""Oracle (test infrastructure): NumPy fp32 restatement of the crop + centre-of-mass front-end.

PARITY UNPINNED (see oracle/__init__.py).  Reference functions restated (all in /root/reference):
  data/preprocess.py:10-79    crop_from_xyz_pose  (bbox from the projected pose +-pad, crop, pad to square, bilinear resize,
                                                   depth threshold, rescaled camera cfg)
  data/preprocess.py:81-129   crop_from_bbx       (NYU test boxes [top,left,bottom,right,d_th], data/nyu.py:109-111)
  data/preprocess.py:131-142  center_of_mass
  data/util.py:20,41-49       _pro / xyz2uvd_op
TF 1.3 semantics used: tf.cast / tf.to_int32 truncate toward zero; tf.image.resize_images default = BILINEAR, align_corners=False,
legacy coordinates src = dst * (in/out) (no half-pixel offset), top/bottom = floor/ceil clipped to in-1, lerp as
top + (bottom - top) * frac in fp32; crop_to_bounding_box / pad_to_bounding_box are zero-filled copies.
"""
import numpy as np

f32 = np.float32


def _resize_bilinear(img, out_h, out_w):
    """tensorflow/core/kernels/resize_bilinear_op.cc (TF 1.x, align_corners=False): img (h,w) fp32 -> (out_h,out_w)."""
    in_h, in_w = img.shape
    hs = f32(in_h) / f32(out_h); ws = f32(in_w) / f32(out_w)
    ys = (np.arange(out_h, dtype=f32) * hs).astype(f32); xs = (np.arange(out_w, dtype=f32) * ws).astype(f32)
    y0 = np.floor(ys).astype(np.int64); y1 = np.minimum(np.ceil(ys).astype(np.int64), in_h - 1)
    x0 = np.floor(xs).astype(np.int64); x1 = np.minimum(np.ceil(xs).astype(np.int64), in_w - 1)
    yl = (ys - y0.astype(f32)).astype(f32)[:, None]; xl = (xs - x0.astype(f32)).astype(f32)[None, :]
    tl = img[y0][:, x0]; tr = img[y0][:, x1]; bl = img[y1][:, x0]; br = img[y1][:, x1]
    top = (tl + (tr - tl) * xl).astype(f32)
    bot = (bl + (br - bl) * xl).astype(f32)
    return (top + (bot - top) * yl).astype(f32)


def _square_resize(dm, top, left, bottom, right, out_hw):
    crop = dm[top:bottom, left:right]
    L = max(bottom - top, right - left)
    off_h = int((L - bottom + top) / 2); off_w = int((L - right + left) / 2)      # tf.to_int32(tf.divide(..)) truncates
    sq = np.zeros((L, L), f32)
    sq[off_h:off_h + crop.shape[0], off_w:off_w + crop.shape[1]] = crop
    return _resize_bilinear(sq, out_hw, out_hw), L, off_h, off_w


def _new_cfg(cfg, top, left, L, off_h, off_w, out_hw):
    ratio = f32(L / out_hw)                                                        # python float division, then cast (:70-71)
    return np.array([cfg[0] / ratio, cfg[1] / ratio, (cfg[2] - f32(left) + f32(off_w)) / ratio,
                     (cfg[3] - f32(top) + f32(off_h)) / ratio, out_hw, out_hw], f32)


def crop_from_xyz_pose(dm, pose, cfg, out_hw=128, pad=20.0, icvl=False):
    """dm (in_h,in_w) mm, pose (3J,) xyz mm, cfg [fx,fy,cx,cy,w,h] -> crop (out,out) fp32, new cfg (6,)."""
    dm = np.asarray(dm, f32); pose = np.asarray(pose, f32).reshape(-1, 3); cfg = np.asarray(cfg, f32)
    in_h, in_w = dm.shape
    u = (pose[:, 0] * cfg[0]) / pose[:, 2] + cfg[2]                                # util.py:20
    v = (pose[:, 1] * cfg[1]) / pose[:, 2] + cfg[3]
    pad = f32(pad)
    top = np.minimum(np.maximum(v.min() - pad, f32(0.0)), cfg[5] - 2 * pad)        # :29-32
    left = np.minimum(np.maximum(u.min() - pad, f32(0.0)), cfg[4] - 2 * pad)
    bottom = np.maximum(np.minimum(v.max() + pad, cfg[5]), f32(top) + 2 * pad - 1)
    right = np.maximum(np.minimum(u.max() + pad, cfg[4]), f32(left) + 2 * pad - 1)
    top, left, bottom, right = int(top), int(left), int(bottom), int(right)        # tf.cast(int32) truncates
    crop, L, off_h, off_w = _square_resize(dm, top, left, bottom, right, out_hw)
    uu = np.clip(u.astype(np.int32), 0, in_w - 1); vv = np.clip(v.astype(np.int32), 0, in_h - 1)   # :56-57
    dd = dm[vv, uu]; dd = dd[dd > 100]
    d_th = (dd.min() + f32(250.0)) if dd.size else f32(np.inf)                     # reduce_min of empty = +inf in TF
    thr = f32(500.0) if icvl else d_th                                             # :62-65
    crop = np.where(crop < thr, crop, f32(0.0)).astype(f32)
    return crop, _new_cfg(cfg, top, left, L, off_h, off_w, out_hw)


def crop_from_bbx(dm, bbx, cfg, out_hw=128):
    """data/preprocess.py:81-129; bbx = [top,left,bottom,right,d_th]."""
    dm = np.asarray(dm, f32); cfg = np.asarray(cfg, f32)
    top, left, bottom, right = [int(x) for x in np.asarray(bbx[:4], f32)]
    crop, L, off_h, off_w = _square_resize(dm, top, left, bottom, right, out_hw)
    crop = np.where(crop < f32(bbx[4]), crop, f32(0.0)).astype(f32)
    return crop, _new_cfg(cfg, top, left, L, off_h, off_w, out_hw)


def center_of_mass(crop, cfg):
    """data/preprocess.py:131-142: mean depth of the positive pixels (>= 200), back-projected crop centre."""
    crop = np.asarray(crop, f32); cfg = np.asarray(cfg, f32)
    c_h, c_w = crop.shape
    pos = crop[crop > 0]
    ave_d = f32(pos.astype(np.float64).mean()) if pos.size else f32(np.nan)
    ave_d = np.maximum(ave_d, f32(200.0)) if pos.size else f32(200.0)   # tf.maximum(nan, 200) is nan in TF; empty crops are degenerate
    ave_u, ave_v = f32(c_w / 2), f32(c_h / 2)
    return np.array([(ave_u - cfg[2]) * ave_d / cfg[0], (ave_v - cfg[3]) * ave_d / cfg[1], ave_d], f32)


# --------------------------------------------------------------------------------------------------------------------
# data augmentation: data/preprocess.py:234-267 data_aug
# --------------------------------------------------------------------------------------------------------------------
def _rotate_nearest(img, cost, sint):
    """tf.contrib.image.rotate(dm, angle) with the default NEAREST interpolation (TF 1.x contrib/image):
    angles_to_projective_transforms -> [cos, -sin, x_off, sin, cos, y_off, 0, 0] maps OUTPUT (x,y) to INPUT (x',y'),
    sampled at (round(y'), round(x')) (std::round, half away from zero) with zero fill."""
    h, w = img.shape
    cost, sint = f32(cost), f32(sint)
    x_off = ((f32(w - 1) - (cost * f32(w - 1) - sint * f32(h - 1))) / f32(2.0)).astype(f32)
    y_off = ((f32(h - 1) - (sint * f32(w - 1) + cost * f32(h - 1))) / f32(2.0)).astype(f32)
    xo = np.arange(w, dtype=f32)[None, :]; yo = np.arange(h, dtype=f32)[:, None]
    xi = ((cost * xo + (-sint) * yo).astype(f32) + x_off).astype(f32)
    yi = ((sint * xo + cost * yo).astype(f32) + y_off).astype(f32)
    rx = np.where(xi >= 0, np.floor(xi + f32(0.5)), np.ceil(xi - f32(0.5))).astype(np.int64)
    ry = np.where(yi >= 0, np.floor(yi + f32(0.5)), np.ceil(yi - f32(0.5))).astype(np.int64)
    ok = (rx >= 0) & (rx < w) & (ry >= 0) & (ry < h)
    out = np.zeros_like(img)
    out[ok] = img[ry[ok], rx[ok]]
    return out


def _resize_nearest(img, out_h, out_w):
    """tf.image.resize_images(method=1) == ResizeNearestNeighbor, align_corners=False: src = min(floor(dst*in/out), in-1)."""
    in_h, in_w = img.shape
    ys = np.minimum(np.floor(np.arange(out_h, dtype=f32) * (f32(in_h) / f32(out_h))).astype(np.int64), in_h - 1)
    xs = np.minimum(np.floor(np.arange(out_w, dtype=f32) * (f32(in_w) / f32(out_w))).astype(np.int64), in_w - 1)
    return img[ys][:, xs]


def _crop_or_pad(img, th, tw):
    """tf.image.resize_image_with_crop_or_pad: centred crop / zero pad with floor-division offsets."""
    h, w = img.shape
    wd, hd = tw - w, th - h
    oc_w, op_w = max(-wd // 2, 0), max(wd // 2, 0)
    oc_h, op_h = max(-hd // 2, 0), max(hd // 2, 0)
    c = img[oc_h:oc_h + min(th, h), oc_w:oc_w + min(tw, w)]
    out = np.zeros((th, tw), img.dtype)
    out[op_h:op_h + c.shape[0], op_w:op_w + c.shape[1]] = c
    return out


def data_aug(dm, pose, cfg, com, cost, sint, edge_ratio):
    """One element of data_aug's map_fn.  dm (h,w) mm, pose (3J,) mm, cfg (6,), com (3,); the random draws of the reference
    (angle ~ U(-pi,pi) -> cost, sint; edge_ratio = clip(N(1,0.2),0.9,1.1) (2,)) are INPUTS so that the result is reproducible."""
    dm = np.asarray(dm, f32); pose = np.asarray(pose, f32).reshape(-1, 3); cfg = np.asarray(cfg, f32); com = np.asarray(com, f32)
    er = np.asarray(edge_ratio, f32); cost, sint = f32(cost), f32(sint)
    h, w = dm.shape
    rot = _rotate_nearest(dm, cost, sint)
    th, tw = int(f32(h) * er[0]), int(f32(w) * er[1])                       # tf.to_int32(tf.to_float(shape)*edge_ratio) :253-254
    out = _crop_or_pad(_resize_nearest(rot, th, tw), h, w)
    # pose: rotate / stretch in uvd about the projected centre of mass (:241-247, :258-262)
    ucom = (com[0] * cfg[0]) / com[2] + cfg[2]; vcom = (com[1] * cfg[1]) / com[2] + cfg[3]
    u = ((pose[:, 0] * cfg[0]) / pose[:, 2] + cfg[2]) - ucom
    v = ((pose[:, 1] * cfg[1]) / pose[:, 2] + cfg[3]) - vcom
    d = pose[:, 2] - com[2]
    ur = (u * cost + v * sint).astype(f32); vr = (u * (-sint) + v * cost).astype(f32)          # row-vector @ rot_mat
    ur = (ur * er[1] + ucom).astype(f32); vr = (vr * er[0] + vcom).astype(f32); dr = (d + com[2]).astype(f32)
    x = ((ur - cfg[2]) * dr / cfg[0]).astype(f32); y = ((vr - cfg[3]) * dr / cfg[1]).astype(f32)   # util.py:21 _bpro
    return out, np.stack([x, y, dr], axis=1).reshape(-1).astype(f32)
