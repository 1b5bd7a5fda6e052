"""CPU tests of the host-side mirror (densereg_b200/model.py): flag surface, result-file format (fixture rows taken
from the reference's exp/result/*.txt -- data, format only), lr schedule, batch sharding, and the N>1 data-parallel
path on world-size-2 gloo: all-reduce(sum) of the flat gradient + identical Adam on every rank == single-process
accumulation over the same micro-batches (model/train_multi_gpu.py:16-39 semantics)."""
import os
import re
import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_flag_surface_matches_reference_defaults():
    from densereg_b200.model import build_argparser
    f = build_argparser().parse_args([])
    # model/hourglass_um_crop_tiny.py:29-62
    assert (f.num_gpus, f.batch_size, f.debug_level, f.sub_batch, f.pid, f.is_train, f.net_module, f.is_aug, f.dataset,
            f.epoch, f.num_stack, f.num_fea, f.kernel_size) == (1, 40, 1, 5, 0, True, "um_v1", True, "nyu", 80, 2, 128, 3)
    f = build_argparser().parse_args("--dataset icvl --batch_size 40 --num_stack 2 --num_fea 128 --is_train False".split())
    assert f.dataset == "icvl" and f.is_train is False      # readme.md:19,36 string booleans


def test_result_row_format_matches_reference_files():
    from densereg_b200.model import format_result_row
    jn = {"icvl": 16, "nyu": 14, "msra": 21}
    for line in open(os.path.join(GOLD, "result_format_rows.txt")):
        ds, row = line.split("|", 1)
        name, *vals = row.rstrip("\n").split("\t")
        assert len(vals) == 3 * jn[ds]
        assert all(re.fullmatch(r"-?\d+\.\d{4}", v) for v in vals)
        assert "/" not in name
        # round trip through our writer reproduces the reference row byte for byte
        assert format_result_row(name.replace("\\", "/"), [float(v) for v in vals]) == row


def test_lr_schedule_and_sharding():
    from densereg_b200.model import shard_batch
    from oracle import um_v1_torch as U
    assert U.lr_at(0, 100) == 1e-3 and abs(U.lr_at(100, 100) - 1e-4) < 1e-12 and abs(U.lr_at(250, 100) - 1e-5) < 1e-12
    for world in (1, 2, 4, 8):
        cover = []
        for r in range(world):
            lo, hi = shard_batch(64, r, world)
            cover += list(range(lo, hi))
        assert cover == list(range(64))


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from densereg_b200.model import allreduce_gradients, shard_batch
    from densereg_b200 import synth
    from oracle import um_v1_torch as U
    torch.set_num_threads(2)
    net = U.Net(1, 16, 4)
    p, s = net.init_params(0, stddev=0.05), net.init_state()
    dms, poses, cfgs, coms = synth.make_batch(2, 4, seed=3)
    lo, hi = shard_batch(2, rank, world)
    _, g, _ = U.loss_and_grads(net, p, s, dms[lo:hi, ..., 0], poses[lo:hi], cfgs[lo:hi], coms[lo:hi], dropout_seed=rank)
    g = allreduce_gradients(g.clone(), world)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    U.adam_step(p, g, m, v, step=1, lr=1e-3, accum_steps=1, world=world)
    if rank == 0:
        torch.save(dict(p=p, g=g), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_data_parallel_world2_gloo_equals_accumulation(tmp_path):
    import torch.multiprocessing as mp
    from densereg_b200 import synth
    from oracle import um_v1_torch as U
    out = str(tmp_path / "rank0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    # single process: accumulate the two shards as two micro-batches, divide by 2
    net = U.Net(1, 16, 4)
    p, s = net.init_params(0, stddev=0.05), net.init_state()
    dms, poses, cfgs, coms = synth.make_batch(2, 4, seed=3)
    gsum = torch.zeros_like(p)
    for r in range(2):
        _, g, _ = U.loss_and_grads(net, p, s.clone(), dms[r:r + 1, ..., 0], poses[r:r + 1], cfgs[r:r + 1], coms[r:r + 1], dropout_seed=r)
        gsum += g
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    U.adam_step(p, gsum, m, v, step=1, lr=1e-3, accum_steps=2, world=1)
    assert float((got["g"] - gsum).abs().max() / gsum.abs().max()) < 1e-5
    assert float((got["p"] - p).abs().max()) < 1e-6


def test_error_curve_matches_reference_definition(tmp_path):
    from densereg_b200.model import error_curve, write_error_curve
    errs = [3.0, 10.4, 10.6, 25.0, 39.0, 41.0, 90.0, 0.2]
    within, curve = error_curve(errs)
    assert within[10] == 3 / 8 and within[20] == 4 / 8 and within[30] == 5 / 8 and within[40] == 6 / 8      # <= t + 0.5
    assert curve[0] == (0.5, 100.0 * 1 / 8) and curve[-1] == (80.5, 100.0 * 7 / 8) and len(curve) == 17   # strict <
    write_error_curve(errs, str(tmp_path / "e.txt"))
    rows = open(tmp_path / "e.txt").read().strip().split("\n")
    assert len(rows) == 17 and rows[0].startswith("0.500000 12.5")


def test_bench_reference_arm_contract():
    """bench.py --impl reference prints ONE JSON line with the contract keys (the CPU restatement, time-boxed)."""
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DENSEREG_REF_WALL_S="60")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=280, env=env, cwd=root)
    line = [l for l in out.stdout.strip().split("\n") if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "crops/s" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_dataset_switch_and_checkpoint_flags(tmp_path):
    """open_datasets: synthetic unless the reference's TFRecord shards exist (hourglass_um_crop_tiny.py:886-906); --restore_step / --data_* flags parse."""
    from densereg_b200 import model as M
    flags = M.build_argparser().parse_args(["--dataset", "icvl", "--is_train", "False", "--data_dir", str(tmp_path / "nope")])
    assert flags.data_source == "auto" and flags.restore_step is None
    logs = []
    ds, val = M.open_datasets(flags, log=logs.append)
    assert isinstance(ds, M.SyntheticDataset) and ds.subset == "training" and val.subset == "testing" and val.jnt_num == 16 and "synthetic" in logs[-1]
    flags.data_source = "tfrecord"
    with pytest.raises(FileNotFoundError):
        M.open_datasets(flags, log=logs.append)
    flags.data_source = "synthetic"
    assert isinstance(M.open_datasets(flags, log=logs.append)[0], M.SyntheticDataset)
    f2 = M.build_argparser().parse_args(["--restore_step", "-1", "--dataset", "msra", "--pid", "3"])
    assert f2.restore_step == -1 and f2.pid == 3
    tr, te = M.open_datasets(f2, log=logs.append)
    assert tr.name == "msra_P3" and tr.jnt_num == 21 and te.exact_num == 8488


def test_result_file_reader_and_comparison(tmp_path):
    """read_result_file parses the reference's exp/result format (fixture rows) and this repo's writer; compare_result_files matches frames by name."""
    from densereg_b200 import model as M
    rows = {}
    for line in open(os.path.join(GOLD, "result_format_rows.txt")):
        ds, row = line.rstrip("\n").split("|", 1)
        rows.setdefault(ds, []).append(row)
    for ds, J in (("icvl", 16), ("nyu", 14), ("msra", 21)):
        p = tmp_path / (ds + ".txt")
        p.write_text("\n".join(rows[ds]) + "\n")
        names, arr = M.read_result_file(str(p))
        assert arr.shape == (len(rows[ds]), 3 * J) and "\\" not in names[0] and (ds == "nyu" or "/" in names[0])       # NYU names have no directory
        # our writer round-trips the same rows byte for byte
        q = tmp_path / (ds + "_b200.txt")
        q.write_text("".join(M.format_result_row(n, v) for n, v in zip(names, arr)))
        assert q.read_text() == p.read_text()
        # shift every joint of the second file by (3, 4, 0) mm -> 5 mm joint distance everywhere; drop the last frame
        shifted = arr + np.tile([3.0, 4.0, 0.0], J)
        q.write_text("".join(M.format_result_row(n, v) for n, v in zip(names[:-1], shifted[:-1])))
        c = M.compare_result_files(str(p), str(q))
        assert c["frames"] == len(names) - 1 and c["only_in_a"] == 1 and c["only_in_b"] == 0
        assert abs(c["mean_joint_dist_mm"] - 5.0) < 1e-3 and abs(c["max_joint_dist_mm"] - 5.0) < 1e-3 and c["within_mm"][10] == 1.0
        assert c["curve"][1] == (5.5, 100.0) and c["curve"][0] == (0.5, 0.0)
    bad = tmp_path / "bad.txt"; bad.write_text("name\t1.0\t2.0\n")
    with pytest.raises(ValueError):
        M.read_result_file(str(bad))
