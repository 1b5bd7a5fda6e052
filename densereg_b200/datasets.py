"""Dataset readers without TensorFlow: ICVL, NYU, MSRA15.

  data/dataset_base.py:28-240   BaseDataset: shard layout, TFRecord writer, queue readers   -> BaseDataset
  data/icvl.py:11-143           IcvlDataset  (labels.txt uvd -> xyz, PNG16)                  -> IcvlDataset
  data/nyu.py:11-220            NyuDataset   (joint_data.mat, y flip, G*256+B depth, 14 of 36 joints, test boxes) -> NyuDataset
  data/msra.py:12-219           MsraDataset  (joint.txt y/z flip, .bin -> PNG16)             -> MsraDataset

The reference parses `tf.train.Example` records inside the TF graph and crops with TF ops; here records are parsed on the host
(densereg_b200/tfrecord.py, png.py), whole frames go to the GPU once, and the crop / centre-of-mass kernels of the C-ABI
(dr_crop_from_xyz_pose, dr_crop_from_bbx) produce the (dms, poses, cfgs, coms) batch that `loss` / `test` consume.  Same class
names, attributes (`cfg`, `name`, `jnt_num`, `pose_dim`, `filenames`, `approximate_num`, `exact_num`, `tf_dir`) and shard file
names as the reference, so shards written by either side are read by the other.
"""
import collections
import os
import pickle
import struct
import threading

import numpy as np

from . import png, tfrecord

CameraConfig = collections.namedtuple("CameraConfig", "fx,fy,cx,cy,w,h")           # data/util.py:9
Annotation = collections.namedtuple("Annotation", "name,pose,bbx", defaults=(None,))  # dataset_base.py:16, nyu.py:10


def uvd2xyz(uvd, cfg):
    """data/util.py:20,33-39 (_bpro): x=(u-cx)*d/fx, y=(v-cy)*d/fy, z=d."""
    uvd = np.asarray(uvd, dtype=np.float64).reshape(-1, 3)
    return np.stack([(uvd[:, 0] - cfg[2]) * uvd[:, 2] / cfg[0], (uvd[:, 1] - cfg[3]) * uvd[:, 2] / cfg[1], uvd[:, 2]], 1)


def xyz2uvd(xyz, cfg):
    """data/util.py:19,24-31 (_pro)."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    return np.stack([xyz[:, 0] * cfg[0] / xyz[:, 2] + cfg[2], xyz[:, 1] * cfg[1] / xyz[:, 2] + cfg[3], xyz[:, 2]], 1)


def _decode_image(data):
    """PNG bytes -> array.  OpenCV when importable (what the reference's own loaders use, nyu.py:144), else densereg_b200/png.py;
    tests check the two agree."""
    try:
        import cv2
    except ImportError:
        return png.decode_png(data)
    img = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise png.PngError("cv2.imdecode failed")
    return img[..., ::-1] if img.ndim == 3 else img                                  # BGR -> RGB


def to_device(a, device):
    """Host array -> tensor on `device`: pinned staging + asynchronous copy for CUDA devices, plain tensor otherwise (CPU tests of the drivers)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    if getattr(device, "type", str(device)) == "cuda":
        return t.pin_memory().to(device, non_blocking=True)
    return t.to(device)


class EndOfData(Exception):
    """tf.errors.OutOfRangeError of a one-epoch reader (test_model.py:64)."""


class BaseDataset(object):
    cfg = None
    approximate_num_per_file = 0
    name = "base"
    directory = "."
    pose_dim = 0
    jnt_num = 0
    crop_pad = 20.0                                                                  # preprocess.py:10 default

    def __init__(self, subset, directory=None):
        self.subset = subset
        if directory is not None:
            self.directory = directory
        self._annotations = None
        self._epoch_iter = None

    # ---- reference interface ----------------------------------------------------------------------------------------
    @property
    def annotations(self):
        return self._annotations

    @property
    def is_train(self):                                                              # icvl.py:47-48: always True
        return True

    @property
    def filenames(self):
        raise NotImplementedError

    @property
    def approximate_num(self):
        return self.approximate_num_per_file * len(self.filenames)

    @property
    def exact_num(self):
        return self.approximate_num

    def available(self):
        """True when every shard of `filenames` exists on disk."""
        try:
            return all(os.path.exists(p) for p in self.filenames)
        except (OSError, AssertionError):
            return False

    def loadAnnotation(self):
        raise NotImplementedError

    def image_path(self, label):
        return os.path.join(self.img_dir, label.name)

    def convert_to_example(self, label):
        """icvl.py:118-128: {'name', 'xyz_pose', 'png16' (the PNG file's bytes, untouched)} [+ 'bbx', nyu.py:164-169]."""
        with open(self.image_path(label), "rb") as f:
            img_data = f.read()
        feat = {"name": label.name.encode("utf-8"), "xyz_pose": np.asarray(label.pose, np.float32), "png16": img_data}
        if label.bbx is not None:
            feat["bbx"] = np.asarray(label.bbx, np.float32).reshape(-1)
        return tfrecord.make_example(feat)

    def saveSampleToRecord(self, idx_list, tar_file_path):                           # dataset_base.py:52-66
        with tfrecord.TFRecordWriter(tar_file_path) as w:
            for idx in idx_list:
                w.write(self.convert_to_example(self.annotations[idx]))

    def shard_name(self, file_idx, num_shards):
        return "%s-%d-of-%d" % (self.subset, file_idx, num_shards)                   # dataset_base.py:82

    def write_TFRecord_multi_thread(self, num_threads, num_shards):
        """dataset_base.py:93-131: same thread / shard index ranges (np.linspace(...).astype(int)) and file names."""
        os.makedirs(self.tf_dir, exist_ok=True)
        assert not num_shards % num_threads, "please make the num_threads commensurate with file_shards"
        if self._annotations is None:
            self.loadAnnotation()
        per_thread = num_shards // num_threads
        spacing = np.linspace(0, len(self.annotations), num_threads + 1).astype(int)
        written = []

        def work(tidx):
            sp = np.linspace(spacing[tidx], spacing[tidx + 1], per_thread + 1).astype(int)
            for k in range(per_thread):
                path = os.path.join(self.tf_dir, self.shard_name(tidx * per_thread + k, num_shards))
                self.saveSampleToRecord(range(sp[k], sp[k + 1]), path)
                written.append(path)

        threads = [threading.Thread(target=work, args=(t,)) for t in range(num_threads)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        return sorted(written)

    # ---- record parsing (tf.parse_single_example + decode_png, icvl.py:131-143) -----------------------------------------------
    def _decode_depth(self, img_data):
        img = _decode_image(img_data)
        if img.ndim != 2:
            raise png.PngError("%s expects a single-channel 16-bit depth PNG" % self.name)
        return img.astype(np.float32)

    def _select_pose(self, pose):
        return pose

    def parse_example(self, example_serialized):
        """-> (image (h,w) float32 mm, pose (pose_dim,) float32 mm, name str, bbx (5,) float32 or None)."""
        feat = tfrecord.parse_example(example_serialized)
        for key in ("name", "xyz_pose", "png16"):
            if feat.get(key) is None:
                raise KeyError("record has no feature %r" % key)                      # FixedLenFeature without default
        image = self._decode_depth(feat["png16"][0])
        if image.shape != (self.cfg.h, self.cfg.w):
            raise ValueError("%s frame is %s, expected %s" % (self.name, image.shape, (self.cfg.h, self.cfg.w)))
        pose = self._select_pose(np.asarray(feat["xyz_pose"], np.float32))
        if pose.shape[0] != self.pose_dim:
            raise ValueError("xyz_pose has %d values, expected %d" % (pose.shape[0], self.pose_dim))
        bbx = feat.get("bbx")
        return image, pose, feat["name"][0].decode("utf-8"), (None if bbx is None else np.asarray(bbx, np.float32))

    # ---- iteration (replaces the TF queue runners, dataset_base.py:154-240) ---------------------------------------------------
    def examples(self, shuffle=False, seed=0, epochs=1):
        """Parsed examples shard by shard.  shuffle=True: shard order reshuffled every epoch and an example-level buffer of
        approximate_num_per_file*8 (the RandomShuffleQueue's min_after_dequeue, dataset_base.py:166-169); epochs=None = forever."""
        rng = np.random.RandomState(seed)
        files = list(self.filenames)
        buf, cap = [], max(self.approximate_num_per_file * 8, 1)
        ep = 0
        while epochs is None or ep < epochs:
            order = rng.permutation(len(files)) if shuffle else range(len(files))
            for fi in order:
                for rec in tfrecord.read_records(files[fi]):
                    if not shuffle:
                        yield self.parse_example(rec)
                        continue
                    buf.append(rec)
                    if len(buf) > cap:
                        k = rng.randint(len(buf))
                        buf[k], buf[-1] = buf[-1], buf[k]
                        yield self.parse_example(buf.pop())
            ep += 1
        while buf:
            k = rng.randint(len(buf))
            buf[k], buf[-1] = buf[-1], buf[k]
            yield self.parse_example(buf.pop())

    def frame_batch(self, batch_size, it, allow_partial=False):
        """Next `batch_size` parsed examples of iterator `it` as stacked host arrays; raises EndOfData when the epoch is over
        (tf.train.batch_join drops a trailing partial batch unless allow_partial)."""
        frames, poses, names, bbxs = [], [], [], []
        for image, pose, name, bbx in it:
            frames.append(image); poses.append(pose); names.append(name); bbxs.append(bbx)
            if len(frames) == batch_size:
                break
        if not frames or (len(frames) < batch_size and not allow_partial):
            raise EndOfData(self.name)
        bb = None if bbxs[0] is None else np.stack(bbxs).astype(np.float32)
        return np.stack(frames), np.stack(poses).astype(np.float32), names, bb

    def crop(self, engine, frames, poses, bbx):
        """preprocess_op (icvl.py:145-150): crop_from_xyz_pose + center_of_mass, on the GPU."""
        return engine.crop_from_xyz_pose(frames, poses, self.cfg, 128, self.crop_pad, icvl=(self.name == "icvl"))

    def batch_device(self, engine, batch_size, seed=0, lo=0, hi=None):
        """-> device tensors (dms (b,128,128,1), poses (b,3J), cfgs (b,6), coms (b,3)) and names; rows [lo,hi) of the global
        batch (the rank's shard, train_multi_gpu.py:63-64).  Training subsets shuffle forever, 'testing' is one ordered epoch."""
        if self._epoch_iter is None:
            train = self.subset != "testing"
            self._epoch_iter = self.examples(shuffle=train, seed=seed, epochs=None if train else 1)
        frames, poses, names, bbx = self.frame_batch(batch_size, self._epoch_iter, allow_partial=(self.subset == "testing"))
        hi = len(names) if hi is None else min(hi, len(names))
        up = lambda a: to_device(a[lo:hi], engine.device)
        f_d, p_d = up(frames), up(poses)
        dms, cfgs, coms = self.crop(engine, f_d, p_d, None if bbx is None else up(bbx))
        return dms, p_d, cfgs, coms, names[lo:hi]


class IcvlDataset(BaseDataset):
    cfg = CameraConfig(fx=241.42, fy=241.42, cx=160, cy=120, w=320, h=240)           # icvl.py:12
    approximate_num_per_file = 220
    name = "icvl"
    max_depth = 500.0
    pose_dim = 48
    jnt_num = 16
    directory = "./exp/data/icvl/"

    def __init__(self, subset, directory=None):
        if subset not in ("training", "training_small", "validation", "testing"):
            raise ValueError("unknown sub %s set to ICVL hand datset" % subset)
        super(IcvlDataset, self).__init__(subset, directory)
        test = subset == "testing"
        self.src_dir = os.path.join(self.directory, "Testing" if test else "Training")
        self.img_dir = os.path.join(self.src_dir, "Depth")
        self.tf_dir = os.path.join(self.directory, "tf_test" if test else "tf_train")

    @property
    def filenames(self):                                                             # icvl.py:54-76 (last shard listed twice)
        tr = lambda n: [os.path.join(self.tf_dir, "training-%d-of-100" % i) for i in range(n)]
        if self.subset == "training":
            files = tr(100)
            return files + [files[-1]]
        if self.subset == "training_small":
            return [f for i, f in enumerate(tr(10)) if i % 10 == 0]
        if self.subset == "validation":
            return [f for i, f in enumerate(tr(10)) if i % 21 == 0]
        files = [os.path.join(self.tf_dir, "testing-%d-of-4" % i) for i in range(4)]
        return files + [files[-1]]

    @property
    def exact_num(self):
        return 1596 if self.subset == "testing" else self.approximate_num           # icvl.py:82-87

    def shard_name(self, file_idx, num_shards):
        return "%s-%d-of-%d" % ("testing" if self.subset == "testing" else "training", file_idx, num_shards)

    def loadAnnotation(self):
        """icvl.py:92-116: labels.txt rows `name u v d ...`; rows not starting with '2014' are dropped (is_train is always True)."""
        self._annotations = []
        with open(os.path.join(self.src_dir, "labels.txt"), "r") as f:
            for line in f:
                if self.is_train and not line.startswith("2014"):
                    continue
                buf = line.split()
                if len(buf) < 2:
                    continue
                pose = np.array([float(d) for d in buf[1:]])
                self._annotations.append(Annotation(buf[0], uvd2xyz(pose, self.cfg).reshape(-1).tolist()))
        return self._annotations


class NyuDataset(BaseDataset):
    cfg = CameraConfig(fx=588.235, fy=587.084, cx=320, cy=240, w=640, h=480)         # nyu.py:13
    approximate_num_per_file = 730
    name = "nyu"
    max_depth = 1500.0
    directory = "./exp/data/nyu/"
    keep_joints = (0, 3, 6, 9, 12, 15, 18, 21, 24, 25, 27, 30, 31, 32)               # nyu.py:40
    orig_pose_dim = 108
    bbx_path = "data/nyu_bbx.pkl"                                                    # nyu.py:110

    def __init__(self, subset, directory=None, bbx_path=None):
        if subset not in ("training", "training_small", "validation", "testing"):
            raise ValueError("unknown sub %s set to NYU hand datset" % subset)
        super(NyuDataset, self).__init__(subset, directory)
        test = subset == "testing"
        self.src_dir = os.path.join(self.directory, "dataset/test" if test else "dataset/train")
        self.img_dir = self.src_dir
        self.tf_dir = os.path.join(self.directory, "tf_test" if test else "tf_train")
        if bbx_path is not None:
            self.bbx_path = bbx_path
        self.keep_pose_idx = np.array([j * 3 + c for j in self.keep_joints for c in range(3)])
        self.pose_dim = len(self.keep_pose_idx)
        self.jnt_num = self.pose_dim // 3

    @property
    def filenames(self):                                                             # nyu.py:62-82
        tr = lambda n: [os.path.join(self.tf_dir, "training-%d-of-300" % i) for i in range(n)]
        if self.subset == "training":
            files = tr(100)
            return files + [files[-1]]
        if self.subset == "training_small":
            return [f for i, f in enumerate(tr(30)) if i % 10 == 0]
        if self.subset == "validation":
            return [f for i, f in enumerate(tr(100)) if i % 21 == 0]
        files = [os.path.join(self.tf_dir, "testing-%d-of-16" % i) for i in range(16)]
        return files + [files[-1]]

    @property
    def exact_num(self):
        return 8252 if self.subset == "testing" else self.approximate_num           # nyu.py:88-93

    def shard_name(self, file_idx, num_shards):
        return "%s-%d-of-%d" % ("testing" if self.subset == "testing" else "training", file_idx, num_shards)

    def loadAnnotation(self, is_trun=False):
        """nyu.py:98-136: joint_data.mat['joint_xyz'][camera] (N,36,3), y negated; names depth_<cam>_<idx:07d>.png; the test set
        carries one detected box [top,left,bottom,right,d_th] per frame (nyu_bbx.pkl)."""
        import scipy.io as sio
        mat = sio.loadmat(os.path.join(self.src_dir, "joint_data.mat"))
        cams = 1 if self.subset == "testing" else 3
        bbxes = None
        if self.subset == "testing":
            with open(self.bbx_path, "rb") as f:
                bbxes = pickle.load(f, encoding="latin1")
        self._annotations = []
        for cam in range(cams):
            joints = np.array(mat["joint_xyz"][cam], dtype=np.float64)
            for idx, j in enumerate(joints):
                j = j.reshape(-1, 3).copy()
                j[:, 1] *= -1.0
                j = j.reshape(-1)
                if is_trun:
                    j = j[self.keep_pose_idx]
                b = None if bbxes is None else np.asarray(bbxes[idx], np.float32).reshape(-1)
                self._annotations.append(Annotation("depth_%d_%07d.png" % (cam + 1, idx + 1), j, b))
        return self._annotations

    def _decode_depth(self, img_data):
        """nyu.py:148-156: depth = G*256 | B of the 8-bit RGB PNG."""
        img = _decode_image(img_data)
        if img.ndim != 3 or img.shape[2] < 3 or img.dtype != np.uint8:
            raise png.PngError("NYU expects an 8-bit RGB depth PNG")
        g, b = img[..., 1].astype(np.uint16), img[..., 2].astype(np.uint16)
        return ((g << 8) | b).astype(np.float32)

    def _select_pose(self, pose):
        if pose.shape[0] != self.orig_pose_dim:
            raise ValueError("NYU xyz_pose has %d values, expected %d" % (pose.shape[0], self.orig_pose_dim))
        return pose[self.keep_pose_idx]                                              # tf.gather_nd, nyu.py:187

    def crop(self, engine, frames, poses, bbx):
        if self.subset == "testing":                                                 # nyu.py:208-214
            if bbx is None:
                raise KeyError("NYU test records need the 'bbx' feature (nyu.py:192-206)")
            return engine.crop_from_bbx(frames, bbx, self.cfg, 128)
        return engine.crop_from_xyz_pose(frames, poses, self.cfg, 128, self.crop_pad, icvl=False)


class MsraDataset(BaseDataset):
    cfg = CameraConfig(fx=241.42, fy=241.42, cx=160, cy=120, w=320, h=240)           # msra.py:13
    approximate_num_per_file = 85
    max_depth = 1000.0
    pose_dim = 63
    jnt_num = 21
    pose_list = "1 2 3 4 5 6 7 8 9 I IP L MP RP T TIP Y".split()
    directory = "./exp/data/msra15/"
    pid_num = [8499, 8492, 8412, 8488, 8500, 8497, 8497, 8498, 8492]

    def __init__(self, subset, pid, directory=None):
        if subset not in ("training", "testing"):
            raise ValueError("unknown sub %s set to MSRA hand datset" % subset)
        super(MsraDataset, self).__init__(subset, directory)
        self.src_dir = os.path.join(self.directory, "P%d" % pid)
        self.img_dir = self.src_dir
        self.tf_dir = os.path.join(self.directory, "tf")
        self.pid = pid
        self.name = "msra_P%d" % pid

    @property
    def filenames(self):
        """msra.py:49-64.  The reference's training list formats every entry with self.pid (so it trains on the held-out
        subject only -- `'P%d-%d-of-100'%(self.pid, i)` inside `for pid in range(9)`); the leave-one-out intent is implemented here
        (all subjects except pid) and recorded as a deviation in DESIGN.md."""
        shards = lambda p: [os.path.join(self.tf_dir, "P%d-%d-of-100" % (p, i)) for i in range(100)]
        if self.subset == "training":
            files = []
            for p in range(9):
                if p != self.pid:
                    files += shards(p)
            return files + [files[-1]]
        files = shards(self.pid)
        return files + [files[-1]]

    @property
    def exact_num(self):
        return self.pid_num[self.pid] if self.subset == "testing" else self.approximate_num

    def shard_name(self, file_idx, num_shards):
        return "P%d-%d-of-%d" % (self.pid, file_idx, num_shards)                     # msra.py:166

    def image_path(self, label):
        return os.path.join(self.img_dir, label.name + ".png")                       # msra.py:177

    def loadAnnotation(self):
        """msra.py:90-111: <gesture>/joint.txt, first line = frame count, then 63 floats per frame with y and z negated."""
        self._annotations = []
        for pose_name in self.pose_list:
            with open(os.path.join(self.src_dir, pose_name, "joint.txt"), "r") as f:
                for frm, line in enumerate(f):
                    if frm == 0:
                        continue
                    v = np.array([float(d) for d in line.split()], dtype=np.float64).reshape(-1, 3)
                    v[:, 1:] *= -1.0
                    self._annotations.append(Annotation(os.path.join(pose_name, "%06i_depth" % (frm - 1)), v.reshape(-1).tolist()))
        return self._annotations

    @staticmethod
    def read_bin(path):
        """msra.py:121-133: int32 {cols, rows, left, top, right, bottom} then the float32 crop -> full-size (rows, cols) frame."""
        with open(path, "rb") as f:
            cols, rows, left, top, right, bottom = struct.unpack("<6i", f.read(24))
            crop = np.frombuffer(f.read(), dtype="<f4")
        dm = np.zeros((rows, cols), np.float32)
        dm[top:bottom, left:right] = crop.reshape(bottom - top, right - left)
        return dm

    def cvtBin2Png(self):
        """msra.py:115-149: every .bin -> 16-bit PNG next to it; an empty frame (sum < 10) repeats the previous one."""
        if self._annotations is None:
            self.loadAnnotation()
        prev = None
        for anno in self._annotations:
            dm = self.read_bin(os.path.join(self.img_dir, anno.name + ".bin"))
            if dm.sum() < 10 and prev is not None:
                dm = prev
            prev = dm.copy()
            with open(os.path.join(self.img_dir, anno.name + ".png"), "wb") as f:
                f.write(png.encode_png(dm.astype(np.uint16), filter_type=2))


def open_dataset(name, subset, pid=0, directory=None):
    """The reference's dataset switch (hourglass_um_crop_tiny.py:886-906)."""
    if name == "icvl":
        return IcvlDataset(subset, directory)
    if name == "nyu":
        return NyuDataset(subset, directory)
    if name == "msra":
        return MsraDataset(subset, pid, directory)
    raise ValueError("unknown dataset %s" % name)
