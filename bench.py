#!/usr/bin/env python
"""bench.py -- depth-crops/sec of the denseReg hot path on N B200s, one rank per GPU.

  python bench.py --gpus N --steps K --warmup W                 default workload = BASELINE.json configs[1] (icvl_train)
  python bench.py --config {icvl_train,nyu64_dp,msra_infer,vote}  the other BASELINE.json configs (2: strong-scaling DP, 3, 4)
  python bench.py --impl reference ...                           CPU restatement of the reference (oracle port), same config

icvl_train / nyu64_dp: one "step" == one optimiser step of model/train_single_gpu.py:138-150: sub_batch micro-batches of forward+backward
(BRN batch statistics, dropout, loss, all gradients), ONE gradient all-reduce across ranks (inside libdensereg_sm100.so, NCCL, overlapped
with the last backward pass), clip +-0.2, Adam.  msra_infer: one step == dr_infer on a batch (forward + vote -> xyz mm).  vote: one step ==
dr_vote on B dense map sets.  `value` = crops/s with inputs resident in HBM; `e2e` = the same step through the public Python API with
pinned HOST inputs (H2D inside the timed region) and the result read back (D2H).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before the CUDA context exists: see densereg_b200/__init__.py

# SURVEY.md 8d / BASELINE.md section 2: algorithmic conv FLOPs per crop (2 x MACs of the conv table); training = 3 x forward
FWD_GFLOP_PER_CROP = {16: 9.790, 14: 9.732, 21: 9.939}
TRAIN_GFLOP_PER_CROP = {16: 29.37, 14: 29.19, 21: 29.82}
CONFIGS = {
    # name: (kind, J, per-step global batch, sub_batch, scaling, description)
    "icvl_train": ("train", 16, 40, 5, "weak", "BASELINE configs[1]: ICVL J=16 2-stack fea=128 training, batch 40 per GPU x sub_batch 5"),
    "nyu64_dp": ("train", 14, 64, 5, "strong", "BASELINE configs[2]: NYU J=14 2-stack fea=128 training, GLOBAL batch 64 split over the GPUs x sub_batch 5"),
    "msra_infer": ("infer", 21, 256, 1, "weak", "BASELINE configs[3]: MSRA J=21 2-stack fea=128 inference (forward + vote), batch 256 per GPU"),
    "vote": ("vote", 21, 4096, 1, "weak", "BASELINE configs[4]: offset-vote microbench, 4096 x 128x128 heat-map + 3-D offset maps, J=21"),
}
METRICS = {"train": "depth-crops/sec (128x128, 2-stack fea=128) training step", "infer": "depth-crops/sec (128x128, 2-stack fea=128) inference",
           "vote": "depth-crops/sec offset-vote (128x128 maps, J=21)"}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernels named in `roofline`, from the `ncu --set full` captures summarised
# under profiles/ (r1_final.md / r2_kernels.md: CTA-pair conv on um_comb/c2 at B=40 = 46.8 MB read + 4.3 MB written; r2_kernels_pair_wgrad.md: CTA-pair wgrad); None = not captured
NCU_TRAFFIC_BYTES = {"conv": 51.1e6, "dgrad": 51.1e6, "wgrad": 90.7e6}      # wgrad_tc_pair_kernel on um_comb/c2, B=40: 86.3 MB read + 4.4 MB written (profiles/r2_kernels_pair_wgrad.md)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], hbm=d["hbm_gbs"], src="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU leg (the ONLY place where bench.py executes oracle/): the reference arm and the cpu_baseline object
# ------------------------------------------------------------------------------------------------
def usable_cpus():
    """Hardware threads this process may really use: affinity mask capped by the cgroup CPU quota (a shared GPU host reports
    128 CPUs but the container may be throttled to a fraction; oversubscribing a quota makes PyTorch-CPU 10-100x slower)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            q, p = open(path).read().split()[:2]
            if q != "max":
                n = min(n, max(1, int(int(q) / int(p))))
        except Exception:
            pass
    try:
        q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            n = min(n, max(1, q // p))
    except Exception:
        pass
    return max(1, n)


PARITY_SAMPLE = dict(n=4, seed=4242, init_seed=0, stddev=0.05)      # crops whose xyz the CPU leg also returns (mean_joint_err_mm)


def _cpu_worker(conn, kind, J, steps, warmup, target_s, max_batch, want_xyz):
    """Child process: CPU restatement of the reference path (oracle port).  Reports after every timed step so that the parent can
    enforce a wall-clock box and still use what was measured.  With want_xyz it first evaluates the reference pipeline (crops -> xyz mm)
    on PARITY_SAMPLE so that the parent can report the mean joint error of the GPU path against it."""
    try:
        import numpy as np
        import torch
        from oracle import um_v1_torch as U, vote_numpy as V
        from densereg_b200 import synth
        threads = usable_cpus()
        torch.set_num_threads(threads)
        net = U.Net(2, 128, J)
        if want_xyz:
            ps = PARITY_SAMPLE
            p, s = net.init_params(ps["init_seed"], stddev=ps["stddev"]), net.init_state()
            dms, poses, cfgs, coms = synth.make_batch(ps["n"], J, seed=ps["seed"])
            x0n = V.norm_dm(dms[..., 0], coms)
            hms, hm3s, ums = net.forward(p, s, torch.from_numpy(x0n[..., None]), training=False)
            ref, _ = V.xyz_estimation(hms[-1].numpy(), hm3s[-1].numpy(), ums[-1].numpy(), V.tiny_dm(x0n), cfgs, coms)
            import tempfile
            wpath = os.path.join(tempfile.gettempdir(), "densereg_parity_weights_%d.npz" % os.getpid())
            np.savez(wpath, params=p.numpy(), state=s.numpy())        # the weights the reference xyz was computed with
            conn.send({"phase": "xyz", "xyz": ref.tolist(), "weights": wpath})
        p, s = net.init_params(0), net.init_state()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        state = {"n": 0}
        if kind == "train":
            def step(batch):
                state["n"] += 1
                dms, poses, cfgs, coms = batch
                L, g, _ = U.loss_and_grads(net, p, s, dms[..., 0], poses, cfgs, coms, dropout_seed=state["n"])
                U.adam_step(p, g, m, v, step=state["n"], lr=1e-3, accum_steps=1, world=1)
            what = "fwd+bwd+Adam on one micro-batch"
        elif kind == "infer":
            def step(batch):
                dms, poses, cfgs, coms = batch
                x0n = V.norm_dm(dms[..., 0], coms)
                with torch.no_grad():
                    hms, hm3s, ums = net.forward(p, s, torch.from_numpy(x0n[..., None]), training=False)
                V.xyz_estimation(hms[-1].numpy(), hm3s[-1].numpy(), ums[-1].numpy(), V.tiny_dm(x0n), cfgs, coms)
            what = "forward + vote on one batch"
        else:
            maps = {}

            def step(batch):
                n = batch[0].shape[0]
                if n not in maps:
                    maps[n] = synth.make_vote_maps(n, J, hw=128, seed=3)
                V.xyz_estimation(*maps[n])
            what = "NumPy vote on 128x128 maps"
            threads = 1
        one = synth.make_batch(1, J, seed=0)
        step(one)                                             # primitive creation / first touch
        t0 = time.perf_counter(); step(one); t1 = time.perf_counter() - t0
        bsz = int(max(1, min(max_batch, round(target_s / max(t1, 1e-3)))))
        sample = ("%s of %d crop(s) per step, %d thread(s) (usable CPUs of %d reported; 1-crop calibration step %.2f s)"
                  % (what, bsz, threads, os.cpu_count() or 0, t1))
        conn.send({"phase": "calib", "value": 1.0 / t1, "steps": 1, "dt": t1, "threads": threads, "sample": sample, "batch": 1})
        batch = synth.make_batch(bsz, J, seed=1)
        for _ in range(warmup):
            step(batch)
        t0 = time.perf_counter()
        for i in range(steps):
            step(batch)
            dt = time.perf_counter() - t0
            conn.send({"phase": "timed", "value": bsz * (i + 1) / dt, "steps": i + 1, "dt": dt, "threads": threads, "sample": sample,
                       "batch": bsz})
        conn.send({"phase": "done"})
    except Exception as e:                                     # pragma: no cover
        conn.send({"phase": "error", "error": repr(e)})


def run_cpu_reference(kind, J, steps, warmup, wall_s, target_s=3.0, max_batch=8, want_xyz=False):
    """-> dict(value, steps, dt, threads, sample, complete[, xyz]) or None.  Never takes longer than wall_s seconds."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe(duplex=False)
    proc = ctx.Process(target=_cpu_worker, args=(child, kind, J, steps, warmup, target_s, max_batch, want_xyz), daemon=True)
    proc.start()
    deadline = time.time() + wall_s
    last, complete, xyz = None, False, None
    while time.time() < deadline:
        if parent.poll(0.5):
            msg = parent.recv()
            if msg.get("phase") == "done":
                complete = True
                break
            if msg.get("phase") == "xyz":
                xyz = {"xyz": msg["xyz"], "weights": msg["weights"]}
                continue
            if msg.get("phase") == "error":
                last = last or {"value": None, "steps": 0, "dt": 0.0, "threads": usable_cpus(), "sample": "oracle failed: " + msg["error"]}
                break
            last = msg
        elif not proc.is_alive():
            break
    if proc.is_alive():
        proc.kill()                                            # exact PID of the child we started
    proc.join(5)
    if last is not None:
        last["complete"] = complete
        last["xyz"] = xyz
        if not complete:
            last["sample"] += " [stopped by the %d s wall-clock box after %d timed step(s)]" % (int(wall_s), last.get("steps", 0))
    return last


def run_reference(args):
    """--impl reference: each step is a bounded sample of the workload (one batch sized by calibration); the whole run is boxed to a few
    minutes.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    kind, J, B, SUB, scaling, desc = CONFIGS[args.config]
    r = run_cpu_reference(kind, J, args.steps, args.warmup, wall_s=float(os.environ.get("DENSEREG_REF_WALL_S", "240")))
    if r is None or r.get("value") is None:
        r = {"value": 0.0, "steps": 0, "dt": 0.0, "threads": usable_cpus(), "sample": "CPU restatement produced no step inside the time box"}
    val = r["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRICS[kind], "value": val, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (r["dt"] / max(r["steps"], 1)) * 1e3, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": desc + " -- CPU restatement of the reference TF graph (PyTorch-CPU fp32 / NumPy; TF 1.3 cannot run here)",
                   "name": args.config, "sample": r["sample"], "timed_steps": r["steps"]},
        "cpu_baseline": {"value": val, "unit": "crops/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": val, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# GPU legs
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def setup(args):
    import torch
    import torch.distributed as dist
    c = Ctx()
    c.torch, c.dist = torch, dist
    c.rank = int(os.environ.get("RANK", 0)); c.world = int(os.environ.get("WORLD_SIZE", 1)); c.local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(c.local)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", c.local))
    c.dev = torch.device("cuda", c.local)
    return c


def timed(c, fn, first, steps, eng=None):
    torch, dist = c.torch, c.dist
    if c.world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count if eng is not None else 0
    e0.record()
    for i in range(steps):
        fn(first + i)
    e1.record()
    torch.cuda.synchronize()
    if c.world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=c.dev)
    if c.world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), (eng.launch_count - l0 if eng is not None else 0)


def cpu_leg(args, kind, J, want_xyz):
    """cpu_baseline object (+ the reference xyz of PARITY_SAMPLE) -- rank 0, N=1 only."""
    r = run_cpu_reference(kind, J, steps=4, warmup=0, wall_s=float(os.environ.get("DENSEREG_CPU_WALL_S", "75")), target_s=4.0, want_xyz=want_xyz)
    if r is not None and r.get("value") is not None:
        return {"value": r["value"], "unit": "crops/s", "cores": r["threads"], "kind": "port",
                "sample": "%d timed step(s): %s" % (r["steps"], r["sample"])}, r.get("xyz")
    return {"value": None, "unit": "crops/s", "cores": usable_cpus(), "kind": "port",
            "sample": "CPU restatement did not finish a step inside the 75 s box"}, (r or {}).get("xyz")


def joint_error_vs_reference(c, J, precision, ref_xyz):
    """mean joint error (mm) of the GPU path (crops -> xyz through dr_infer) against the CPU leg's reference xyz on PARITY_SAMPLE
    (data/evaluation.py:15-18 meanJntError, averaged over crops)."""
    import numpy as np
    from densereg_b200.engine import DenseRegEngine
    from densereg_b200 import synth
    ps = PARITY_SAMPLE
    e = DenseRegEngine(2, 128, J, max_batch=ps["n"], precision=precision, device=c.local, training=False)
    w = np.load(ref_xyz["weights"])                   # the flat parameter / BRN-state vectors the CPU leg used
    e.load_flat(c.torch.from_numpy(w["params"]), c.torch.from_numpy(w["state"]))
    try:
        os.remove(ref_xyz["weights"])
    except OSError:
        pass
    dms, poses, cfgs, coms = synth.make_batch(ps["n"], J, seed=ps["seed"])
    xyz = e.infer(*[c.torch.from_numpy(x).to(c.dev) for x in (dms, cfgs, coms)]).cpu().numpy()
    e.close()
    ref = np.asarray(ref_xyz["xyz"], np.float64).reshape(ps["n"], J, 3)
    err = np.linalg.norm(xyz.reshape(ps["n"], J, 3) - ref, axis=-1)
    ok = np.isfinite(err)
    return {"value": float(err[ok].mean()), "max": float(err[ok].max()), "unit": "mm", "joints": int(ok.sum()),
            "vs": "CPU restatement of the reference (oracle) on %d seeded synthetic crops, same weights (trunc-normal sigma %.2f)" % (ps["n"], ps["stddev"]),
            "bar_mm": 1e-3}


def class_roofline(eng, B, peaks, run_traced):
    """Per-class roofline of the conv-type kernels from ONE extra traced pass (every launch timed alone with CUDA events on its own
    stream -- serialised, outside the timed region): achieved = sum of algorithmic FLOPs / sum of launch times for forward conv, dgrad,
    wgrad.  Returns (roofline of the time-dominant class, list of all classes, share of the traced micro-batch)."""
    eng.trace(True)
    run_traced()
    eng.torch_sync()
    recs = eng.trace_records()
    eng.trace(False)
    cls = {}
    for r in recs:
        a = cls.setdefault(r["kind"], {"ms": 0.0, "flops": 0.0, "n": 0, "kernels": {}})
        a["ms"] += r["ms"]; a["n"] += 1
        a["flops"] += 2.0 * r["B"] * r["hw"] * r["hw"] * r["k"] * r["k"] * r["cin"] * r["cout"]
        a["kernels"][r["kernel"]] = a["kernels"].get(r["kernel"], 0) + 1
    out = []
    for k, a in cls.items():
        ach = a["flops"] / (a["ms"] * 1e-3) / 1e12 if a["ms"] > 0 else 0.0
        out.append({"class": k, "launches": a["n"], "avg_launch_ms": a["ms"] / max(a["n"], 1), "total_ms": a["ms"], "achieved": ach, "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_burst"], "kernels": a["kernels"]})
    out.sort(key=lambda d: -d["total_ms"])
    return out


def run_train(args, c):
    torch, dist = c.torch, c.dist
    from densereg_b200.engine import DenseRegEngine
    from densereg_b200 import synth
    kind, J, GB, SUB, scaling, desc = CONFIGS[args.config]
    if args.batch_size:
        GB = args.batch_size
    if args.sub_batch:
        SUB = args.sub_batch
    world, rank, dev = c.world, c.rank, c.dev
    if scaling == "strong":
        assert GB % world == 0, "global batch %d not divisible by %d GPUs" % (GB, world)
        B = GB // world
    else:
        B = GB
    S, F = 2, 128
    eng = DenseRegEngine(S, F, J, max_batch=B, precision=args.precision, device=c.local, training=True, pipeline=args.pipeline)
    eng.torch_sync = torch.cuda.synchronize
    eng.init_params(seed=0)                       # same seed on every rank -> identical replicas without a broadcast
    eng.comm_init(rank, world)                    # in-library NCCL communicator (id broadcast through torch.distributed)
    NROT = 4                                      # distinct input batches rotated through (host + device copies)
    host = [synth.make_batch(B, J, seed=1000 * rank + i) for i in range(NROT)]
    pinned = [[torch.from_numpy(a).pin_memory() for a in hb] for hb in host]
    resident = [[t.to(dev) for t in hb] for hb in pinned]
    h2d_bytes = SUB * sum(t.numel() * 4 for t in pinned[0])
    loss_host = torch.zeros(5).pin_memory()

    def step_resident(i):
        eng.zero_grads()
        for sub in range(SUB):
            d, po, cf, co = resident[(i * SUB + sub) % NROT]
            if sub == SUB - 1:
                eng.comm_overlap_next_backward()
            eng.loss_backward(d, po, cf, co, dropout_seed=i * SUB + sub)
        eng.optimizer_step(i + 1, 1e-3, accum_steps=SUB, world=world)

    # end to end: every micro-batch's inputs are copied from pinned host memory inside the timed region, ONE micro-batch ahead on a copy
    # stream into three rotating device staging sets (the usual loader prefetch).  After a pipelined dr_loss_backward the caller's stream is
    # ordered behind that micro-batch's forward pass and loss, so copies issued on it would start that late and delay the next forward pass.
    NST = 3
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [[torch.empty_like(t, device=dev) for t in pinned[0]] for _ in range(NST)]
    ev_ready = [torch.cuda.Event() for _ in range(NST)]
    ev_consumed = [torch.cuda.Event() for _ in range(NST)]
    fetched = {"n": -1}

    def prefetch(n):                                  # n = running micro-batch index
        if n <= fetched["n"]:
            return
        fetched["n"] = n
        j = n % NST
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed[j])    # micro-batch n - 3 has read this set (its loss is computed)
            for a, b in zip(staging[j], pinned[n % NROT]):
                a.copy_(b, non_blocking=True)
            ev_ready[j].record(copy_stream)

    def step_e2e(i):
        cur = torch.cuda.current_stream()
        eng.zero_grads()
        prefetch(i * SUB)
        for sub in range(SUB):
            n = i * SUB + sub
            prefetch(n + 1)                           # also across the optimiser step: the copy does not touch parameters or gradients
            cur.wait_event(ev_ready[n % NST])
            if sub == SUB - 1:
                eng.comm_overlap_next_backward()
            loss = eng.loss_backward(*staging[n % NST], dropout_seed=n)
            ev_consumed[n % NST].record(cur)          # the call returns with the stream ordered behind this micro-batch's loss kernel
        eng.optimizer_step(i + 1, 1e-3, accum_steps=SUB, world=world)
        loss_host.copy_(loss, non_blocking=True)

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(c.local)
    if rank == 0:
        sampler.start()
    ar0 = eng.allreduce_count
    ms, launches = timed(c, step_resident, args.warmup, args.steps, eng)
    n_allreduce = eng.allreduce_count - ar0
    n_e2e_warm = max(1, args.warmup // 2)
    for i in range(n_e2e_warm):
        step_e2e(i)
    ms_e2e, _ = timed(c, step_e2e, n_e2e_warm, args.steps, eng)
    clocks = sampler.stop() if rank == 0 else None
    crops = B * SUB * args.steps * world
    value = crops / (ms * 1e-3)
    value_e2e = crops / (ms_e2e * 1e-3)

    # ---- roofline.  (1) per class of conv-type kernel over one traced micro-batch (time-dominant class first);
    #      (2) the best single layer (conv on s0/um_comb/c2, 3x3 256->256 @32x32, 12.3 % of the MACs per stack) timed alone over
    #      rotating > L2 inputs -- what the kernel reaches when the main loop dominates.
    peaks = measured_peaks()
    classes = class_roofline(eng, B, peaks, lambda: eng.loss_backward(*resident[0], dropout_seed=12345, update_state=False))
    traced_ms = sum(cl["total_ms"] for cl in classes)
    names = [l["name"] for l in eng.layers()]
    li = names.index("s0/um_comb/c2")
    xs = [torch.randn(B, 32, 32, 256, device=dev) for _ in range(4)]       # 4 x 42 MB inputs + outputs > 126 MB L2 at B=40
    yb = eng.debug_conv(li, xs[0], args.precision)
    for x in xs:
        eng.debug_conv(li, x, args.precision, reuse_weights=True, out=yb)
    torch.cuda.synchronize()
    reps = 12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        eng.debug_conv(li, xs[r % 4], args.precision, reuse_weights=True, out=yb)
    e1.record(); torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    k_flops = 2.0 * B * 1024 * 9 * 256 * 256
    best = k_flops / (k_ms * 1e-3) / 1e12
    step_tflops = value / world * TRAIN_GFLOP_PER_CROP[J] / 1e3
    top = classes[0] if classes else {"class": "wgrad", "achieved": 0.0, "frac": 0.0, "avg_launch_ms": 0.0, "launches": 0, "total_ms": 0.0}
    # the time-dominant class (wgrad) on ITS heaviest layer, timed alone like best_layer: per-launch FLOPs / per-launch time, with the
    # per-launch DRAM traffic of the same launch from the ncu capture
    dyb = torch.randn(B, 32, 32, 256, device=dev)
    for r in range(3):
        eng.debug_conv_bwd(li, xs[r % 4], dyb, args.precision, want_dx=False)
    torch.cuda.synchronize()
    e0.record()
    for r in range(reps):
        eng.debug_conv_bwd(li, xs[r % 4], dyb, args.precision, want_dx=False)
    e1.record(); torch.cuda.synchronize()
    w_ms = e0.elapsed_time(e1) / reps
    w_ach = k_flops / (w_ms * 1e-3) / 1e12
    # top-level numbers: the kernel of the TIME-DOMINANT class (conv / dgrad share one kernel) on the heaviest layer of the path, timed alone;
    # `classes` has every class over all its layers, `other_kernel` the same layer through the other kernel family
    per_kernel = {
        "wgrad": {"name": "wgrad_tc_pair_kernel (filter gradient, CTA pairs)", "ms": w_ms, "achieved": w_ach, "traffic": NCU_TRAFFIC_BYTES["wgrad"],
                  "note": " (includes the memset of the 2.4 MB gradient)"},
        "conv": {"name": "conv_tc_pair_kernel (implicit-GEMM conv, CTA pairs; dgrad runs the same kernel on rotated weights)", "ms": k_ms, "achieved": best,
                 "traffic": NCU_TRAFFIC_BYTES["conv"], "note": ""}}
    dom = "wgrad" if top["class"] == "wgrad" else "conv"
    oth = "conv" if dom == "wgrad" else "wgrad"
    pk = per_kernel[dom]
    roofline = {"bound": "tensor", "achieved": pk["achieved"], "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": pk["achieved"] / peaks["bf16_burst"],
                "traffic": pk["traffic"] if B == 40 and args.precision == "tf32x3" else None,
                "kernel": "%s on the heaviest layer s0/um_comb/c2 3x3 256->256, B=%d, timed alone%s.  %s is the TIME-DOMINANT kernel class of the step: %d "
                          "launches per micro-batch, %.2f ms = %.0f %% of the conv-type time at %.0f TFLOP/s over all its layers (`classes`)"
                          % (pk["name"], B, pk["note"], top["class"], top["launches"], top["total_ms"], 100.0 * top["total_ms"] / max(traced_ms, 1e-9), top["achieved"]),
                "kernel_ms": pk["ms"], "algorithmic_bytes": 4.0 * B * 1024 * 256 * 2 + 4.0 * 9 * 256 * 256,
                "algorithmic_flops": k_flops,
                "peak_source": peaks["src"] + " dense bf16 cuBLAS burst (tf32 kind is nominally half of bf16; 3xTF32 issues 3 MMAs per algorithmic MAC: ceiling = peak / 6)",
                "dominant_class": top["class"], "classes": classes,
                "other_kernel": {"kernel": "%s on the same layer, timed alone%s" % (per_kernel[oth]["name"], per_kernel[oth]["note"]),
                                 "kernel_ms": per_kernel[oth]["ms"], "achieved": per_kernel[oth]["achieved"], "frac": per_kernel[oth]["achieved"] / peaks["bf16_burst"],
                                 "traffic": per_kernel[oth]["traffic"] if B == 40 and args.precision == "tf32x3" else None},
                "whole_step": {"achieved": step_tflops, "peak": peaks["bf16_sustained"], "frac": step_tflops / peaks["bf16_sustained"],
                               "note": "%.2f GFLOP/crop (fwd+dgrad+wgrad conv FLOPs) x crops/s per GPU vs sustained measured peak" % TRAIN_GFLOP_PER_CROP[J]}}

    if rank == 0:
        cpu_baseline, mje = None, None
        if not args.no_cpu_baseline and world == 1:
            cpu_baseline, ref_xyz = cpu_leg(args, "train", J, want_xyz=True)
            if ref_xyz is not None:
                mje = joint_error_vs_reference(c, J, args.precision, ref_xyz)
        others = None
        eng_depth = eng.pipeline_depth
        ws_gb = eng.workspace_bytes / 2**30 / eng_depth
        if world == 1 and args.config == "icvl_train" and not args.no_other_configs and not args.batch_size and not args.sub_batch:
            eng.close(); del eng; torch.cuda.empty_cache()
            others = other_configs(args)
        print(json.dumps({
            "metric": METRICS["train"], "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": {"fp32": "fp32", "tf32": "tf32", "tf32x3": "fp32 (3xTF32 split, two-level accumulation)"}[args.precision], "data": "synthetic",
            "config": {"workload": desc + ": optimiser step = %d micro-batches x batch %d per GPU, fwd+bwd+allreduce+clip+Adam" % (SUB, B),
                       "name": args.config, "global_batch": B * world, "sub_batch": SUB, "parallelism": "dp%d" % world,
                       "micro_batch_pipeline": "depth %d%s" % (eng_depth, " (forward of micro-batch i+1 overlaps backward of micro-batch i; forward passes in "
                                                                "order, one backward at a time, second activation arena)" if eng_depth == 2 else ""),
                       "collective": "1 all-reduce(sum) of the 23.4 MB flat gradient per step inside libdensereg_sm100.so (NCCL, %d bucket call(s) per step "
                                     "overlapped with the last backward pass)" % (n_allreduce // max(args.steps, 1)) if world > 1 else "none (1 GPU)",
                       "l2": "working set ~%.1f GB per micro-batch >> 126 MB L2; inputs rotated over %d batches" % (ws_gb, NROT)},
            "e2e": {"value": value_e2e, "unit": "crops/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 20,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "mean_joint_err_mm": mje,
            "other_configs": others,
        }))


def other_configs(args):
    """BASELINE.json configs[2..4] next to the headline, so that the driver's one default run also carries them: each is this same script in a child
    process (`--config ...`, its own CUDA context, a few steps), reduced to the numbers that matter.  Rank 0, N=1, default config only."""
    out = {}
    for name, extra in (("msra_infer", ["--steps", "10"]), ("vote", ["--steps", "5"]), ("nyu64_dp", ["--steps", "3"])):
        cmd = [sys.executable, os.path.abspath(__file__), "--config", name, "--no_cpu_baseline", "--no_other_configs", "--precision", args.precision,
               "--warmup", "3", "--pipeline", str(args.pipeline)] + extra
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if not line:
                out[name] = {"error": (r.stderr or "no output")[-300:]}
                continue
            d = json.loads(line[-1])
            out[name] = {"metric": d["metric"], "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "steps": d["steps"],
                         "workload": d["config"]["workload"], "e2e": d["e2e"], "gpu_launches": d["gpu_launches"],
                         "roofline": {k: d["roofline"].get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "frac_convention")
                                      if k in d["roofline"]}}
        except Exception as e:                                   # a side measurement must never take the headline line down
            out[name] = {"error": repr(e)[:300]}
    return out


def run_infer(args, c):
    torch, dist = c.torch, c.dist
    import numpy as np
    from densereg_b200.engine import DenseRegEngine
    from densereg_b200 import synth
    kind, J, B, SUB, scaling, desc = CONFIGS[args.config]
    if args.batch_size:
        B = args.batch_size
    world, rank, dev = c.world, c.rank, c.dev
    eng = DenseRegEngine(2, 128, J, max_batch=B, precision=args.precision, device=c.local, training=False, infer_graph=True)
    eng.init_params(0, 0.05)
    NROT = 4
    base = [synth.make_batch(min(B, 64), J, seed=100 * rank + i) for i in range(NROT)]
    rep = (B + base[0][0].shape[0] - 1) // base[0][0].shape[0]
    pinned = [[torch.from_numpy(np.concatenate([x] * rep)[:B]).pin_memory() for x in (hb[0], hb[2], hb[3])] for hb in base]
    # the CUDA graph is keyed on the buffers: fixed device input / output buffers, contents refreshed per step
    d, cf, co = [t.to(dev) for t in pinned[0]]
    resident = [[t.to(dev) for t in hb] for hb in pinned]
    xyz = torch.empty(B, 3 * J, device=dev)
    res = torch.empty(B, 3 * J).pin_memory()
    h2d = sum(t.numel() * 4 for t in pinned[0]); d2h = res.numel() * 4

    def step_resident(i):
        src = resident[i % NROT]
        d.copy_(src[0], non_blocking=True)            # device-to-device refresh of the graph's input buffer (168 MB/s-class work, in the timed region)
        eng.infer(d, cf, co, out=xyz)

    def step_e2e(i):
        src = pinned[i % NROT]
        d.copy_(src[0], non_blocking=True); cf.copy_(src[1], non_blocking=True); co.copy_(src[2], non_blocking=True)
        eng.infer(d, cf, co, out=xyz)
        res.copy_(xyz, non_blocking=True)

    steps = max(args.steps, 10)
    l_before = eng.launch_count
    step_resident(0)                                  # first call: the kernels are launched into the capturing stream and counted
    graph_nodes = eng.launch_count - l_before
    for i in range(1, max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(c.local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(c, step_resident, 0, steps, eng)
    for i in range(3):
        step_e2e(i)
    ms_e2e, _ = timed(c, step_e2e, 0, steps, eng)
    clocks = sampler.stop() if rank == 0 else None
    value = B * steps * world / (ms * 1e-3)
    value_e2e = B * steps * world / (ms_e2e * 1e-3)
    peaks = measured_peaks()
    tfl = value / world * FWD_GFLOP_PER_CROP[J] / 1e3
    if rank == 0:
        cpu_baseline, mje = None, None
        if not args.no_cpu_baseline and world == 1:
            cpu_baseline, ref_xyz = cpu_leg(args, "infer", J, want_xyz=True)
            if ref_xyz is not None:
                mje = joint_error_vs_reference(c, J, args.precision, ref_xyz)
        print(json.dumps({
            "metric": METRICS["infer"], "value": value, "unit": "crops/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "fp32", "tf32": "tf32", "tf32x3": "fp32 (3xTF32 split, two-level accumulation)"}[args.precision], "data": "synthetic",
            "config": {"workload": desc + " (eval-mode BRN folded into the conv epilogue, CUDA-graph replay), replicas only -- no collective",
                       "name": args.config, "batch_per_gpu": B, "parallelism": "replicas x%d" % world,
                       "l2": "activation arena %.1f GB >> 126 MB L2; inputs rotated over %d batches" % (eng.workspace_bytes / 2**30, NROT)},
            "e2e": {"value": value_e2e, "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / steps},
            "gpu_launches": launches if launches else graph_nodes * steps, "graph": "one CUDA graph replay per step, %d kernel nodes" % graph_nodes, "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": tfl, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": tfl / peaks["bf16_sustained"],
                         "traffic": None, "kernel": "whole forward pass: %.3f GFLOP/crop of conv FLOPs x crops/s vs %s sustained bf16 peak" % (FWD_GFLOP_PER_CROP[J], peaks["src"])},
            "cpu_baseline": cpu_baseline, "mean_joint_err_mm": mje,
        }))


def run_vote(args, c):
    torch, dist = c.torch, c.dist
    from densereg_b200.engine import DenseRegEngine
    kind, J, B, SUB, scaling, desc = CONFIGS[args.config]
    if args.batch_size:
        B = args.batch_size
    H = 128
    world, rank, dev = c.world, c.rank, c.dev
    eng = DenseRegEngine(1, 64, 16, max_batch=1, device=c.local, training=False)
    g = torch.Generator(device=dev).manual_seed(rank)
    hm = torch.rand(B, H, H, J, device=dev, generator=g) * 1.2 - 0.1
    hm3 = torch.rand(B, H, H, J, device=dev, generator=g).clamp_(0.05, 1)
    um = torch.randn(B, H, H, 3 * J, device=dev, generator=g)
    dmn = torch.where(torch.rand(B, H, H, device=dev, generator=g) < 0.6, torch.full((), -1.0, device=dev),
                      torch.rand(B, H, H, device=dev, generator=g) * 1.3 - 0.4)
    cfgs = torch.tensor([[240., 240., 64., 64., 128., 128.]], device=dev).repeat(B, 1)
    coms = torch.tensor([[0., 0., 400.]], device=dev).repeat(B, 1)
    # end to end: the maps live in pinned HOST memory in chunks of CH samples and stream through two device staging sets
    CH = min(B, 256)
    hchunk = [t[:CH].cpu().pin_memory() for t in (hm, hm3, um, dmn, cfgs, coms)]
    dstage = [[torch.empty_like(t[:CH]) for t in (hm, hm3, um, dmn, cfgs, coms)] for _ in range(2)]
    res = torch.empty(CH, 3 * J).pin_memory()
    h2d = sum(t.numel() * 4 for t in hchunk) * (B // CH); d2h = res.numel() * 4 * (B // CH)

    def step_resident(i):
        eng.vote(hm, hm3, um, dmn, cfgs, coms)

    def step_e2e(i):
        for k in range(B // CH):
            st = dstage[k & 1]
            for a, b in zip(st, hchunk):
                a.copy_(b, non_blocking=True)
            x = eng.vote(*st)
            res.copy_(x, non_blocking=True)

    steps = max(args.steps, 5)
    for i in range(3):
        step_resident(i)
    sampler = ClockSampler(c.local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(c, step_resident, 0, steps, eng)
    step_e2e(0)
    ms_e2e, _ = timed(c, step_e2e, 0, 2, eng)
    clocks = sampler.stop() if rank == 0 else None
    value = B * steps * world / (ms * 1e-3)
    value_e2e = B * 2 * world / (ms_e2e * 1e-3)
    peaks = measured_peaks()
    bytes_alg = 4.0 * H * H * (5 * J + 1) * B            # (5J+1)*4*H*W per crop
    bytes_streamed = 4.0 * H * H * (2 * J + 1) * B       # what the kernel must read: hm, hm3, dm (um is gathered at 5 winners per joint)
    k_ms = ms / steps
    if rank == 0:
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            cpu_baseline, _ = cpu_leg(args, "vote", J, want_xyz=False)
        print(json.dumps({
            "metric": METRICS["vote"], "value": value, "unit": "crops/s", "n_gpus": world, "steps": steps, "warmup": 3,
            "ms_per_step": k_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 (no FMA contraction; int32 indices)",
            "data": "synthetic",
            "config": {"workload": desc, "name": args.config, "batch_per_gpu": B, "parallelism": "replicas x%d" % world,
                       "l2": "inputs %.1f GB per step >> 126 MB L2" % (bytes_alg / 2**30)},
            "e2e": {"value": value_e2e, "unit": "crops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / 2,
                    "note": "maps streamed from pinned host memory in chunks of %d crops" % CH},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": bytes_streamed / (k_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": bytes_streamed / (k_ms * 1e-3) / 1e9 / peaks["hbm"], "traffic": None,
                         "kernel": "vote_kernel; bytes = what the kernel must stream, (2J+1)*4*H*W per crop (hm, hm3, dm; um is gathered at the 5 winners per joint)",
                         "survey_convention": {"bytes_per_crop": 4.0 * H * H * (5 * J + 1), "achieved": bytes_alg / (k_ms * 1e-3) / 1e9,
                                               "frac": bytes_alg / (k_ms * 1e-3) / 1e9 / peaks["hbm"],
                                               "note": "SURVEY.md 8d counts all of um as read; the kernel never streams it, so this figure exceeds the HBM peak"}},
            "cpu_baseline": cpu_baseline,
        }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=str, default="icvl_train", choices=sorted(CONFIGS))
    ap.add_argument("--batch_size", type=int, default=0, help="override the config's batch")
    ap.add_argument("--sub_batch", type=int, default=0, help="override the config's micro-batches per optimiser step")
    ap.add_argument("--precision", type=str, default="tf32x3", choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--pipeline", type=int, default=2, choices=[1, 2], help="training: micro-batch pipeline depth (2 = forward of micro-batch i+1 next to backward of i)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_other_configs", action="store_true", help="default config at N=1 only: do not also measure msra_infer / vote / nyu64_dp in child processes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    c = setup(args)
    kind = CONFIGS[args.config][0]
    {"train": run_train, "infer": run_infer, "vote": run_vote}[kind](args, c)
    if c.world > 1:
        c.dist.destroy_process_group()


if __name__ == "__main__":
    main()
