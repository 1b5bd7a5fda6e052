"""CPU test of the host-side drivers train() / test() / save() / restore() / main() of densereg_b200/model.py (the mirror of
model/train_single_gpu.py:37-177 and model/test_model.py:14-94) with a stand-in engine: the control flow, sharding, logs, result files and
the checkpoint round trip are host logic and must not need a GPU to be checked.  The stand-in does NO arithmetic of the hot path (it
returns constants); the real engine is exercised by the -m gpu tests."""
import os

import numpy as np
import pytest
import torch

from densereg_b200 import model as M
from densereg_b200 import tf_checkpoint as T
from oracle import um_v1_torch as O


class FakeEngine:
    """Same Python surface as densereg_b200.engine.DenseRegEngine, host tensors, call log instead of kernels."""
    instances = []

    def __init__(self, num_stack=2, num_fea=128, num_jnt=16, max_batch=40, precision="fp32", device=0, kernel_size=3, training=True, **kw):
        self.S, self.F, self.J, self.max_batch, self.precision = num_stack, num_fea, num_jnt, max_batch, precision
        self.device = torch.device("cpu")
        specs, self.n_params, self.n_state = O.build_specs(num_stack, num_fea, num_jnt)
        self._layers = [dict(name=c.name, k=c.k, cin=c.cin, cout=c.cout, brn=int(c.brn), w_off=c.w_off, p_off=c.p_off, s_off=c.s_off) for c in specs]
        self.params = torch.zeros(self.n_params); self.state = torch.zeros(self.n_state)
        self.grads = torch.zeros(self.n_params) if training else None
        self.adam_m = torch.zeros(self.n_params) if training else None
        self.adam_v = torch.zeros(self.n_params) if training else None
        self.calls = []
        FakeEngine.instances.append(self)

    def layers(self): return self._layers
    def init_params(self, seed=0, stddev=0.01):
        g = torch.Generator().manual_seed(seed); self.params.normal_(0, stddev, generator=g); self.calls.append(("init", seed))
    def load_flat(self, params, state=None):
        self.params.copy_(params)
        if state is not None: self.state.copy_(state)
    def zero_grads(self): self.grads.zero_(); self.calls.append(("zero",))
    def join(self): self.calls.append(("join",))
    def loss_backward(self, dms, poses, cfgs, coms, dropout_seed=0, update_state=True):
        assert dms.shape[1:] == (128, 128, 1) and poses.shape[1] == 3 * self.J and cfgs.shape[1] == 6 and coms.shape[1] == 3
        self.grads += 1.0; self.calls.append(("loss", dms.shape[0], dropout_seed))
        return torch.tensor([10.0, 1.0, 2.0, 3.0, 0.5])
    def optimizer_step(self, step, lr, accum_steps=1, world=1):
        self.params -= lr; self.adam_m += 1; self.calls.append(("opt", step, lr, accum_steps, world))
    def infer(self, dms, cfgs, coms, out=None, top5=None):
        self.calls.append(("infer", dms.shape[0]))
        return torch.arange(dms.shape[0] * 3 * self.J, dtype=torch.float32).reshape(dms.shape[0], 3 * self.J) * 0.5
    def data_aug(self, dms, poses, cfgs, coms, cossin, er):
        assert cossin.shape == (dms.shape[0], 2) and er.shape == (dms.shape[0], 2)
        assert torch.allclose(cossin.pow(2).sum(1), torch.ones(dms.shape[0]), atol=1e-5) and float(er.min()) >= 0.9 - 1e-6 and float(er.max()) <= 1.1 + 1e-6
        self.calls.append(("aug", dms.shape[0]))
        return dms, poses
    def close(self): pass


@pytest.fixture
def fake(monkeypatch, tmp_path):
    FakeEngine.instances = []
    monkeypatch.setattr(M, "DenseRegEngine", FakeEngine)
    monkeypatch.setattr(M.JointDetectionModel, "_base_dir", str(tmp_path / "train_cache"))
    monkeypatch.chdir(tmp_path)
    return tmp_path


def _flags(*extra):
    return M.build_argparser().parse_args(["--dataset", "icvl", "--num_stack", "1", "--num_fea", "64", "--batch_size", "4", "--sub_batch", "2",
                                           "--data_source", "synthetic"] + list(extra))


def test_train_loop_sharding_logs_and_checkpoints(fake, monkeypatch):
    reduced = []
    monkeypatch.setattr(M, "allreduce_gradients", lambda g, world: reduced.append((float(g[0]), world)))   # the collective itself: tests/test_host.py (gloo)
    flags = _flags("--max_steps", "101", "--is_aug", "True")
    ds, val = M.open_datasets(flags, log=lambda *_: None)
    model = M.JointDetectionModel(ds, flags, val_dataset=val)
    assert model.name == "icvl_training_s1_f64_daug_um_v1" and model.train_dir.endswith(model.name)      # hourglass_um_crop_tiny.py:95-104
    assert model.decay_steps == int(220 * 101 / (4 * 2) * 10) and abs(model.lr_at(model.decay_steps) - 1e-4) < 1e-12
    logs = []
    M.train(model, rank=1, world=2, log=logs.append)                      # rank 1 of 2: rows [2,4) of every micro-batch, no files written
    eng = model.engine
    losses = [c for c in eng.calls if c[0] == "loss"]
    assert len(losses) == 101 * 2 and all(c[1] == 2 for c in losses)      # batch 4 split over 2 ranks
    assert [c[2] for c in losses[:4]] == [1, 3, 5, 7]                     # dropout seed = micro-step * world + rank
    assert sum(1 for c in eng.calls if c[0] == "aug") == 202 and sum(1 for c in eng.calls if c[0] == "zero") == 101
    opts = [c for c in eng.calls if c[0] == "opt"]
    assert opts[0][1:] == (1, 1e-3, 2, 2) and opts[-1][1] == 101
    assert len(reduced) == 101 and reduced[0] == (2.0, 2)                 # ONE all-reduce per optimiser step, after both micro-batches
    # micro-batch pipeline: the host-side collective reads the gradient buffer itself, so the pipeline is joined first -- once per step,
    # after the last loss() and before the optimiser step
    joins = [i for i, c in enumerate(eng.calls) if c[0] == "join"]
    assert len(joins) == 101 and all(eng.calls[i - 1][0] == "loss" and eng.calls[i + 1][0] == "opt" for i in joins)
    assert model.flags.pipeline == 2                                      # default of the training driver
    assert not os.path.exists(model.train_dir) and logs == []             # only rank 0 logs / saves
    # rank 0, single process: logs every 5 steps, validation every 40, checkpoint at step 100 in BOTH formats
    M.train(model, rank=0, world=1, log=logs.append)
    import re
    # the reference's format string (train_single_gpu.py:155) with the MEAN loss over the sub_batch micro-batches
    assert len(logs) == 21 and re.match(r"\[model/train_multi_gpu\] .*: step 0/101, loss = 10\.000, \d+\.\d{3} sec/batch, \d+\.\d{3} sec/sample "
                                        r"\(hm 1\.00 hm3 2\.00 um 3\.00 reg 0\.500, lr 1\.0e-03\)$", logs[0]), logs[0]
    tl = open(os.path.join(model.train_dir, "training_log.txt")).read().splitlines()
    vl = open(os.path.join(model.train_dir, "validation_log.txt")).read().splitlines()
    assert len([l for l in tl if l.startswith("[model")]) == 21 and len([l for l in tl if l.startswith("model has been saved")]) == 2 and len(vl) == 3
    # checkpoints every 100 steps AND after the last step (train_single_gpu.py:168)
    assert os.path.exists(os.path.join(model.train_dir, "model.ckpt-101.pt")) and vl[1].startswith("step 40 mean joint error (mm):") and len(vl[1].split(":")[1].split()) == 3
    assert os.path.exists(os.path.join(model.train_dir, "model.ckpt-100.pt")) and os.path.exists(os.path.join(model.train_dir, "model.ckpt-100.index"))
    tensors = T.read_bundle(os.path.join(model.train_dir, "model.ckpt-100"))
    assert float(tensors["global_step"]) == 100.0 and tensors["hg_imgproc/Conv/weights"].shape == (7, 7, 1, 32)
    # restore into a fresh model: parameters, Adam slots and the step come back from the TF bundle
    m2 = M.JointDetectionModel(ds, flags, val_dataset=val)
    assert m2.has_checkpoint(100) and not m2.has_checkpoint(7)
    snap = T.load_into_flat(tensors, eng.layers(), eng.n_params, eng.n_state)[0]
    assert m2.restore(100) == 100
    assert np.array_equal(m2.engine.params.numpy(), snap) and float(m2.engine.adam_m.max()) > 0
    os.remove(os.path.join(model.train_dir, "model.ckpt-100.index"))      # without the bundle the flat .pt is used
    m3 = M.JointDetectionModel(ds, flags, val_dataset=val)
    assert m3.restore(100) == 100 and np.array_equal(m3.engine.params.numpy(), snap)


def test_test_driver_writes_reference_format(fake):
    flags = _flags("--is_train", "False", "--test_num", "10")
    ds, val = M.open_datasets(flags, log=lambda *_: None)
    model = M.JointDetectionModel(ds, flags, val_dataset=val)
    out = str(fake / "exp" / "result" / "icvl_b200.txt")
    mean_err, max_err = M.test(model, out_path=out, log=lambda *_: None)
    rows = open(out).read().splitlines()
    assert len(rows) == 10 and np.isfinite(mean_err) and max_err >= mean_err
    name, *vals = rows[0].split("\t")
    assert name == "testing_seq\\image_000000.png" and len(vals) == 48 and vals[1] == "0.5000"        # '/' -> '\\', %.4f (test_model.py:74-75)
    curve = open(out.replace(".txt", "_error.txt")).read().splitlines()
    assert len(curve) == 17 and curve[0].split()[0] == "0.500000"                                      # evaluation.py:99-101
    assert [c[1] for c in model.engine.calls if c[0] == "infer"] == [4, 4, 4]


def test_main_restores_reference_checkpoint_for_testing(fake, capsys):
    # a checkpoint in the reference's own format, named like the authors' pretrained models (model.ckpt--1), is picked up by `--is_train False`
    specs, n_params, n_state = O.build_specs(1, 64, 16)
    layers = [dict(name=c.name, k=c.k, cin=c.cin, cout=c.cout, brn=int(c.brn), w_off=c.w_off, p_off=c.p_off, s_off=c.s_off) for c in specs]
    params = np.full(n_params, 0.25, np.float32); state = np.full(n_state, 0.5, np.float32)
    d = os.path.join(M.JointDetectionModel._base_dir, "icvl_training_s1_f64_daug_um_v1")
    T.write_bundle(os.path.join(d, "model.ckpt--1"), T.flat_to_tensors(layers, params, state))
    M.main(["--dataset", "icvl", "--num_stack", "1", "--num_fea", "64", "--batch_size", "4", "--data_source", "synthetic", "--is_train", "False",
            "--test_num", "4"])
    eng = FakeEngine.instances[-1]
    assert float(eng.params.min()) == 0.25 and float(eng.state.max()) == 0.5
    assert "restored" in capsys.readouterr().out
    with pytest.raises(FileNotFoundError):
        M.main(["--dataset", "icvl", "--num_stack", "1", "--num_fea", "64", "--batch_size", "4", "--data_source", "synthetic", "--is_train", "False",
                "--test_num", "4", "--restore_step", "123"])


def test_short_run_saves_final_checkpoint_and_untrained_test_needs_opt_in(fake):
    """a run shorter than 100 steps still leaves a checkpoint (train_single_gpu.py:168 saves after the last step), and
    `--is_train False` without any checkpoint fails like saver.restore unless --allow_random_init is given."""
    flags = _flags("--max_steps", "7", "--is_aug", "False")
    ds, val = M.open_datasets(flags, log=lambda *_: None)
    model = M.JointDetectionModel(ds, flags, val_dataset=val)
    M.train(model, rank=0, world=1, log=lambda *_: None)
    assert model.has_checkpoint(7) and not model.has_checkpoint(100)
    base = ["--dataset", "icvl", "--num_stack", "1", "--num_fea", "64", "--batch_size", "4", "--data_source", "synthetic", "--is_train", "False",
            "--test_num", "4", "--is_aug", "True"]                          # another model directory (…_daug): no checkpoint there
    with pytest.raises(FileNotFoundError):
        M.main(base)
    M.main(base + ["--allow_random_init", "True"])
    with pytest.raises(AssertionError):
        M.shard_batch(10, 0, 4)                                              # train_multi_gpu.py:59
