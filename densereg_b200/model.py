"""Host-side mirror of the reference's model object, drivers and flag surface for the hot path.

  model/hourglass_um_crop_tiny.py:29-62   tf.app.flags             -> build_argparser (same names / defaults)
  model/hourglass_um_crop_tiny.py:66-191  JointDetectionModel      -> JointDetectionModel (same attribute names)
  model/train_single_gpu.py:37-177        train(model)             -> train(model): accumulate sub_batch micro-batches,
                                                                      ONE all-reduce, clip +-0.2, Adam, staircase lr
  model/train_multi_gpu.py:41-158         in-graph towers          -> one process per GPU + torch.distributed all_reduce
  model/test_model.py:14-94               test(model)              -> test(model): writes name\\t%.4f... rows, '/'->'\\\\'
  data/evaluation.py:9-18                 maxJntError/meanJntError -> evaluation helpers below

All arithmetic happens in libdensereg_sm100.so through DenseRegEngine; this file only moves buffers,
steps counters and writes text.  Datasets: the TFRecord shards of data/{icvl,nyu,msra}.py are read by
densereg_b200/datasets.py when they exist under the reference's directories (`--data_source tfrecord|auto`); none are on
the box, so by default SyntheticDataset generates seeded synthetic crops of each dataset's shape
(densereg_b200/synth.py).  Both kinds expose `batch_device(engine, batch_size, seed, lo, hi)`.
"""
import argparse
import os
import time
from datetime import datetime

import numpy as np
import torch

from . import datasets, synth, tf_checkpoint
from .engine import DenseRegEngine


def str2bool(v):
    # the reference parses 'True'/'False' strings (readme.md:19,36)
    return str(v).lower() in ("1", "true", "yes")


def build_argparser():
    p = argparse.ArgumentParser(description="densereg_b200 (flag surface of model/hourglass_um_crop_tiny.py:29-62)")
    p.add_argument("--num_gpus", type=int, default=1)
    p.add_argument("--batch_size", type=int, default=40)
    p.add_argument("--debug_level", type=int, default=1)
    p.add_argument("--sub_batch", type=int, default=5)
    p.add_argument("--pid", type=int, default=0)
    p.add_argument("--is_train", type=str2bool, default=True)
    p.add_argument("--net_module", type=str, default="um_v1")
    p.add_argument("--is_aug", type=str2bool, default=True)
    p.add_argument("--dataset", type=str, default="nyu")
    p.add_argument("--epoch", type=int, default=80)
    p.add_argument("--num_stack", type=int, default=2)
    p.add_argument("--num_fea", type=int, default=128)
    p.add_argument("--kernel_size", type=int, default=3)
    # additions (not in the reference): arithmetic mode and a bound on synthetic steps
    p.add_argument("--precision", type=str, default="tf32x3", choices=["fp32", "tf32", "tf32x3"])
    p.add_argument("--pipeline", type=int, default=2, choices=[1, 2],
                   help="2 = micro-batch pipeline: the forward pass of micro-batch i+1 runs next to the backward pass of micro-batch i (second activation arena)")
    p.add_argument("--max_steps", type=int, default=0, help="stop after this many optimiser steps (0 = epoch schedule)")
    p.add_argument("--test_num", type=int, default=0, help="number of synthetic test frames (0 = dataset's exact_num)")
    p.add_argument("--restore_step", type=int, default=None,
                   help="checkpoint step to restore from train_dir (testing defaults to -1 like run_test, hourglass_um_crop_tiny.py:909)")
    p.add_argument("--data_source", type=str, default="auto", choices=["auto", "synthetic", "tfrecord"],
                   help="tfrecord = the reference's shards under --data_dir; auto = tfrecord when every shard exists, else synthetic")
    p.add_argument("--data_dir", type=str, default=None, help="dataset root (default: the reference's ./exp/data/<dataset>/)")
    p.add_argument("--allow_random_init", type=str2bool, default=False,
                   help="--is_train False without a checkpoint: test the freshly initialised network instead of failing like saver.restore")
    return p


class SyntheticDataset:
    """Stands in for data/{icvl,nyu,msra}.py: same name / jnt_num / approximate_num / exact_num, synthetic crops."""
    _SPEC = {  # name: (jnt_num, approximate_num, exact_num)   data/icvl.py:13-17,85  nyu.py:14,40-45,92  msra.py:14,17,70
        "icvl": (16, 220 * 101, 1596),
        "nyu": (14, 730 * 101, 8252),
        "msra": (21, 85 * 801, 8499),
    }

    def __init__(self, name, subset, pid=0):
        self.name, self.subset, self.pid = ("msra_P%d" % pid if name == "msra" else name), subset, pid    # msra.py:34
        self.jnt_num, self.approximate_num, self.exact_num = self._SPEC[name]
        if name == "msra":
            self.exact_num = datasets.MsraDataset.pid_num[pid]                     # msra.py:66-73
        self._cursor = 0

    def batch(self, batch_size, seed):
        """-> dms (B,128,128,1) mm, poses (B,3J) mm, cfgs (B,6), coms (B,3), names"""
        dms, poses, cfgs, coms = synth.make_batch(batch_size, self.jnt_num, seed=seed)
        names = ["%s_seq/image_%06d.png" % (self.subset, self._cursor + i) for i in range(batch_size)]
        self._cursor += batch_size
        return dms, poses, cfgs, coms, names

    def batch_device(self, engine, batch_size, seed=0, lo=0, hi=None):
        """Rows [lo,hi) of the global batch as device tensors (same contract as datasets.BaseDataset.batch_device)."""
        dms, poses, cfgs, coms, names = self.batch(batch_size, seed)
        tens = [datasets.to_device(a[lo:hi], engine.device) for a in (dms, poses, cfgs, coms)]
        return tens[0], tens[1], tens[2], tens[3], names[lo:hi]


def meanJntError(skel1, skel2):      # data/evaluation.py:15-18
    d = np.asarray(skel1).reshape(-1, 3) - np.asarray(skel2).reshape(-1, 3)
    return float(np.mean(np.sqrt(np.sum(d ** 2, axis=1))))


def maxJntError(skel1, skel2):       # data/evaluation.py:9-12
    d = np.asarray(skel1).reshape(-1, 3) - np.asarray(skel2).reshape(-1, 3)
    return float(np.max(np.sqrt(np.sum(d ** 2, axis=1))))


def error_curve(max_errors):
    """data/evaluation.py:63-103 plotError: fraction of frames whose MAX joint error is below 10.5/20.5/30.5/40.5 mm and the
    (threshold, percentage) curve at 0.5, 5.5, ..., 80.5 mm that the reference writes to `<result>_error.txt`."""
    s = sorted(float(x) for x in max_errors)
    n = max(len(s), 1)
    within = {t: sum(1 for v in s if v <= t + 0.5) / n for t in (10, 20, 30, 40)}
    thresh = [t * 5.0 + 0.5 for t in range(17)]
    curve = [(t, 100.0 * sum(1 for v in s if v < t) / n) for t in thresh]
    return within, curve


def write_error_curve(max_errors, path):
    within, curve = error_curve(max_errors)
    with open(path, "w") as f:
        for t, p in curve:
            f.write("%f %f\n" % (t, p))                                                 # evaluation.py:99-101
    return within


def format_result_row(name, xyz_val):
    """model/test_model.py:74-75."""
    res_str = "%s\t%s\n" % (name, "\t".join(format(float(pt), ".4f") for pt in xyz_val))
    return res_str.replace("/", "\\")


def read_result_file(path):
    """Rows written by test() / model/test_model.py:70-76 (`name\\t%.4f\\t...`, '/' stored as '\\\\') -> ([names], (N, 3J) float64 array).  Reads the reference's
    published exp/result/{icvl,nyu,msra}.txt as well as this repo's files."""
    names, rows = [], []
    with open(path) as f:
        for ln, line in enumerate(f, 1):
            line = line.rstrip("\n")
            if not line:
                continue
            parts = line.split("\t")
            if len(parts) < 4 or (len(parts) - 1) % 3:
                raise ValueError("%s:%d: expected name + 3J tab-separated values, got %d fields" % (path, ln, len(parts)))
            names.append(parts[0].replace("\\", "/"))
            rows.append([float(v) for v in parts[1:]])
    if rows and len({len(r) for r in rows}) != 1:
        raise ValueError("%s: rows with different joint counts" % path)
    return names, np.asarray(rows, np.float64)


def compare_result_files(path_a, path_b):
    """Frame-by-frame comparison of two result files (e.g. the reference's exp/result/icvl.txt and a run of this engine restored from the authors'
    checkpoint): frames are matched by name; returns mean / max joint distance (mm) and the evaluation.py error curve of the per-frame maxima."""
    na, a = read_result_file(path_a)
    nb, b = read_result_file(path_b)
    ib = {n: i for i, n in enumerate(nb)}
    common = [(i, ib[n]) for i, n in enumerate(na) if n in ib]
    if not common:
        raise ValueError("no frame names in common")
    if a.shape[1] != b.shape[1]:
        raise ValueError("different joint counts: %d vs %d values per row" % (a.shape[1], b.shape[1]))
    mean_e = [meanJntError(a[i], b[j]) for i, j in common]
    max_e = [maxJntError(a[i], b[j]) for i, j in common]
    within, curve = error_curve(max_e)
    return {"frames": len(common), "only_in_a": len(na) - len(common), "only_in_b": len(nb) - len(common),
            "mean_joint_dist_mm": float(np.mean(mean_e)), "max_joint_dist_mm": float(np.max(max_e)), "within_mm": within, "curve": curve}


class JointDetectionModel:
    """Same public attributes as the reference object (hourglass_um_crop_tiny.py:66-191, 436-543)."""
    _init_lr = 0.001
    _lr_decay_factor = 0.1
    _adam_beta1 = 0.5
    _num_epochs_per_decay = {"nyu": 10, "msra": 20, "icvl": 10}   # icvl undefined in the reference (:70-73) -> 10 (deviation)
    _base_dir = "./exp/train_cache/"
    TOWER_NAME = "um_v1"

    def __init__(self, dataset, flags, val_dataset=None, device=0, world=1):
        self._dataset, self._val_dataset, self.flags = dataset, val_dataset, flags
        self._jnt_num = int(dataset.jnt_num)
        self._num_batches_per_epoch = dataset.approximate_num / (flags.batch_size * flags.sub_batch)   # :109
        self._max_steps = int(flags.epoch * self._num_batches_per_epoch)                                # :112
        self._model_desc = "%s_%s_s%d_f%d" % (dataset.name, dataset.subset, flags.num_stack, flags.num_fea)
        if flags.is_aug:
            self._model_desc += "_daug"
        self.world = world
        self.engine = DenseRegEngine(flags.num_stack, flags.num_fea, self._jnt_num, max_batch=flags.batch_size,
                                     precision=flags.precision, device=device, kernel_size=flags.kernel_size,
                                     training=bool(flags.is_train), pipeline=getattr(flags, "pipeline", 1))

    # ---- trainer/tester contract -----------------------------------------------------
    @property
    def init_lr(self): return self._init_lr
    @property
    def lr_decay_factor(self): return self._lr_decay_factor
    @property
    def decay_steps(self): return int(self._num_batches_per_epoch * self._num_epochs_per_decay[self.flags.dataset])
    @property
    def max_steps(self): return self._max_steps
    @property
    def name(self): return "%s_%s" % (self._model_desc, self.TOWER_NAME)
    @property
    def train_dir(self): return os.path.join(self._base_dir, self.name)
    @property
    def train_dataset(self): return self._dataset
    @property
    def val_dataset(self): return self._val_dataset
    @property
    def is_validate(self): return self._val_dataset is not None

    def lr_at(self, step):
        """tf.train.exponential_decay(staircase=True), train_single_gpu.py:45-49."""
        return self._init_lr * self._lr_decay_factor ** (step // max(self.decay_steps, 1))

    # ---- device-side calls --------------------------------------------------------------------------
    def loss(self, dms, poses, cfgs, coms, dropout_seed=0):
        """One micro-batch of model.loss + accum_op; device tensors in, device loss vector out."""
        return self.engine.loss_backward(dms, poses, cfgs, coms, dropout_seed=dropout_seed)

    def test(self, dms, cfgs, coms, out=None):
        """model.test: raw crops -> xyz mm (B,3J)."""
        return self.engine.infer(dms, cfgs, coms, out=out)

    # ---- checkpoints ----------------------------------------------------------------
    # `model.ckpt-<step>.{index,data-00000-of-00001}` is the reference's own format (tf.train.Saver V2 bundle with the um_v1
    # variable names, densereg_b200/tf_checkpoint.py): what train() of the reference writes, what the authors' pretrained
    # models ship as, and what restore() reads here.  save() writes it too, next to a flat `.pt` of the same buffers.
    def save(self, step, tf_bundle=True):
        os.makedirs(self.train_dir, exist_ok=True)
        path = os.path.join(self.train_dir, "model.ckpt-%d.pt" % step)
        e = self.engine
        torch.save(dict(step=step, params=e.params.cpu(), state=e.state.cpu(), adam_m=e.adam_m.cpu(), adam_v=e.adam_v.cpu(),
                        config=dict(num_stack=e.S, num_fea=e.F, num_jnt=e.J)), path)
        if tf_bundle:
            tf_checkpoint.export_checkpoint(e, os.path.join(self.train_dir, "model.ckpt-%d" % step), global_step=step)
        return path

    def restore(self, step):
        """saver.restore(sess, train_dir/model.ckpt-<step>) (test_model.py:31-35; the reference tests with step -1)."""
        prefix = os.path.join(self.train_dir, "model.ckpt-%d" % step)
        e = self.engine
        if os.path.exists(prefix + ".index"):
            got = tf_checkpoint.import_checkpoint(e, prefix)
            return got if got else step
        ck = torch.load(prefix + ".pt", map_location="cpu")
        e.load_flat(ck["params"], ck["state"])
        if e.adam_m is not None:
            e.adam_m.copy_(ck["adam_m"]); e.adam_v.copy_(ck["adam_v"])
        return ck["step"]

    def has_checkpoint(self, step):
        prefix = os.path.join(self.train_dir, "model.ckpt-%d" % step)
        return os.path.exists(prefix + ".index") or os.path.exists(prefix + ".pt")


def augment(engine, tens, rng):
    """data_aug (data/preprocess.py:234-267): angle ~ U(-pi,pi), edge_ratio = clip(N(1,0.2),0.9,1.1); draws on the host, warp on the GPU."""
    dms, poses, cfgs, coms = tens
    B = dms.shape[0]
    ang = rng.uniform(-np.pi, np.pi, size=B).astype(np.float32)
    cossin = torch.from_numpy(np.stack([np.cos(ang), np.sin(ang)], 1).astype(np.float32)).to(dms.device)
    er = torch.from_numpy(np.clip(rng.normal(1.0, 0.2, size=(B, 2)), 0.9, 1.1).astype(np.float32)).to(dms.device)
    return engine.data_aug(dms, poses, cfgs, coms, cossin, er)


def shard_batch(batch_size, rank, world):
    """tf.split of the global minibatch across towers (train_multi_gpu.py:63-64) -> [lo, hi) of this rank."""
    assert batch_size % world == 0, "batch_size must be divisible by the number of GPUs (train_multi_gpu.py:59)"
    per = batch_size // world
    return rank * per, (rank + 1) * per


def allreduce_gradients(grads, world):
    """The ONE collective of the path: sum over ranks of the flat gradient buffer (replaces
    train_multi_gpu.py:16-39 _average_gradients; the 1/world factor is folded into dr_optimizer_step)."""
    if world > 1:
        torch.distributed.all_reduce(grads, op=torch.distributed.ReduceOp.SUM)
    return grads


def train(model, rank=0, world=1, log=print, start_step=0):
    """model/train_single_gpu.py:37-177 (loop :138-175) with per-rank sharding of each micro-batch.  Writes the reference's
    `training_log.txt` (:135-158) and `validation_log.txt` (hourglass_um_crop_tiny.py:126,816-840) under train_dir."""
    f = model.flags
    eng = model.engine
    max_steps = f.max_steps or model.max_steps
    lo, hi = shard_batch(f.batch_size, rank, world)
    dev = eng.device
    rng = np.random.RandomState(1234 + rank)
    tlog = vlog = None
    if rank == 0:
        os.makedirs(model.train_dir, exist_ok=True)
        tlog = open(os.path.join(model.train_dir, "training_log.txt"), "a")
        vlog = open(os.path.join(model.train_dir, "validation_log.txt"), "a")
    bad = None                                                                     # device-side "some loss was not finite" flag (no sync per step)
    for step in range(start_step, max_steps):                                      # resume: train_single_gpu.py:125-128
        t_step = time.time()
        eng.zero_grads()                                                           # reset_op :139
        ave_loss = None
        def fetch(sub):                                                            # one micro-batch of device inputs (+ augmentation, hourglass_um_crop_tiny.py:333-334)
            t = list(model.train_dataset.batch_device(eng, f.batch_size, seed=step * f.sub_batch + sub, lo=lo, hi=hi)[:4])
            if f.is_aug:
                t[0], t[1] = augment(eng, t, rng)
            return t
        nxt = fetch(0)
        for sub in range(f.sub_batch):                                             # :140-148
            tens = nxt                                                                 # inputs are prepared one micro-batch ahead: after a pipelined loss() the
            if sub + 1 < f.sub_batch:                                                  # caller's stream is ordered behind that micro-batch's forward pass
                nxt = fetch(sub + 1)
            loss = model.loss(*tens, dropout_seed=(step * f.sub_batch + sub) * world + rank)
            ave_loss = loss.clone() if ave_loss is None else ave_loss + loss           # ave_loss += loss_value :147 (device add, no sync)
            nf = ~torch.isfinite(loss[0])
            bad = nf if bad is None else (bad | nf)
        if world > 1:
            eng.join()                                                             # micro-batch pipeline: every backward pass has written its gradients
        allreduce_gradients(eng.grads, world)
        eng.optimizer_step(step + 1, model.lr_at(step), accum_steps=f.sub_batch, world=world)   # train_op :150
        if step % 5 == 0 or step + 1 == max_steps:                                 # every rank checks its own micro-batches (:146), one sync per 5 steps
            assert not bool(bad), "Model diverged with loss = NaN (rank %d, step <= %d)" % (rank, step)
        if step % 5 == 0 and rank == 0:                                            # :154-158
            lv = (ave_loss / f.sub_batch).cpu().numpy()
            duration = time.time() - t_step
            msg = ("[model/train_multi_gpu] %s: step %d/%d, loss = %.3f, %.3f sec/batch, %.3f sec/sample"    # the reference's format string :155
                   % (datetime.now(), step, max_steps, lv[0], duration, duration / (f.batch_size * f.sub_batch)))
            msg += " (hm %.2f hm3 %.2f um %.2f reg %.3f, lr %.1e)" % (lv[1], lv[2], lv[3], lv[4], model.lr_at(step))
            log(msg); tlog.write(msg + "\n"); tlog.flush()
        if step % 40 == 0 and rank == 0 and model.is_validate:                     # do_test every 40 steps :165-166
            try:                                                                   # batch of 3 like the reference (:62-65)
                vd, vp, vc, vm, _ = model.val_dataset.batch_device(eng, 3, seed=900_000 + step)
            except datasets.EndOfData:
                continue
            xyz = model.test(vd, vc, vm).cpu().numpy()
            err = [meanJntError(x, g) for x, g in zip(xyz, vp.cpu().numpy())]
            vlog.write("step %d mean joint error (mm): %s\n" % (step, " ".join("%.3f" % e for e in err))); vlog.flush()
        if ((step + 1) % 100 == 0 or step + 1 == max_steps) and rank == 0:         # :168-175 (every 100 steps AND after the last one)
            model.save(step + 1)
            tlog.write("model has been saved to %s\n" % os.path.join(model.train_dir, "model.ckpt")); tlog.flush()
    if rank == 0:
        tlog.close(); vlog.close()
    return max_steps


def test(model, out_path=None, log=print):
    """model/test_model.py:14-94: loop batches, write result rows, return (mean, max) joint error vs the synthetic GT."""
    f = model.flags
    ds = model.val_dataset or model.train_dataset
    total = f.test_num or ds.exact_num
    out_path = out_path or os.path.join("exp", "result", "%s_b200.txt" % ds.name)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    errs, maxs, n, step = [], [], 0, 0
    dev = model.engine.device
    with open(out_path, "w") as fo:
        while n < total:
            try:
                dms, poses, cfgs, coms, names = ds.batch_device(model.engine, f.batch_size, seed=10_000 + step)
            except datasets.EndOfData:                                             # OutOfRangeError ends the loop, test_model.py:64
                break
            xyz = model.test(dms, cfgs, coms).cpu().numpy()
            for xyz_val, gt_val, name in zip(xyz, poses.cpu().numpy(), names):
                errs.append(meanJntError(xyz_val, gt_val)); maxs.append(maxJntError(xyz_val, gt_val))
                fo.write(format_result_row(name, xyz_val))
                n += 1
                if n >= total:
                    break
            step += 1
    within = write_error_curve([v for v in maxs if np.isfinite(v)], out_path.replace(".txt", "_error.txt"))   # test_model.py:82
    log("finish test: %d frames, mean joint err %.3f mm, mean max-joint err %.3f mm, <=10/20/30/40 mm: %s -> %s"
        % (n, float(np.nanmean(errs)), float(np.nanmean(maxs)), " ".join("%.3f" % within[t] for t in (10, 20, 30, 40)), out_path))
    return float(np.nanmean(errs)), float(np.nanmean(maxs))


def open_datasets(flags, log=print):
    """The dataset switch of hourglass_um_crop_tiny.py:886-906: always (Dataset('training'), Dataset('testing')) -- the model / checkpoint
    directory name comes from the TRAINING dataset also when testing (run_test, :873-884)."""
    if flags.data_source != "synthetic":
        real = [datasets.open_dataset(flags.dataset, s, flags.pid, flags.data_dir) for s in ("training", "testing")]
        need = real if flags.is_train else real[1:]
        if all(d.available() for d in need):
            log("[densereg_b200] reading TFRecord shards from %s" % need[0].tf_dir)
            return real[0], real[1]
        if flags.data_source == "tfrecord":
            missing = [p for d in need for p in d.filenames if not os.path.exists(p)]
            raise FileNotFoundError("TFRecord shards missing (first: %s)" % missing[0])
        log("[densereg_b200] no TFRecord shards under %s -- using synthetic %s-shaped crops" % (need[0].tf_dir, flags.dataset))
    return SyntheticDataset(flags.dataset, "training", flags.pid), SyntheticDataset(flags.dataset, "testing", flags.pid)


def main(argv=None):
    flags = build_argparser().parse_args(argv)
    if flags.net_module != "um_v1":
        raise SystemExit("only --net_module um_v1 exists (network/um_v1.py)")
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    if flags.is_train and flags.batch_size % world != 0:
        raise SystemExit("--batch_size %d is not divisible by the %d GPUs (train_multi_gpu.py:59)" % (flags.batch_size, world))
    ds, val = open_datasets(flags)
    model = JointDetectionModel(ds, flags, val_dataset=val, device=local, world=world)
    model.engine.init_params(seed=0)
    step = flags.restore_step if flags.restore_step is not None else (None if flags.is_train else -1)
    start_step = 0
    if step is not None and model.has_checkpoint(step):
        start_step = model.restore(step)
        print("[densereg_b200] restored %s/model.ckpt-%d" % (model.train_dir, step))
    elif flags.restore_step is not None:
        raise FileNotFoundError("no checkpoint %s/model.ckpt-%d(.index|.pt)" % (model.train_dir, step))
    elif not flags.is_train:
        if not flags.allow_random_init:                        # saver.restore would fail here (test_model.py:31-35)
            raise FileNotFoundError("no checkpoint %s/model.ckpt-%d(.index|.pt); pass --allow_random_init True to test an untrained network"
                                    % (model.train_dir, step))
        print("[densereg_b200] no checkpoint %s/model.ckpt-%d -- testing the freshly initialised network (--allow_random_init)" % (model.train_dir, step))
    if flags.is_train:
        train(model, rank, world, start_step=max(start_step, 0))
    else:
        test(model)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
