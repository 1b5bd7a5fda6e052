#!/usr/bin/env python
"""bench.py -- depth-crops/sec of the denseReg training step on N B200s (BASELINE.json configs[1]:
ICVL 16-joint, 2-stack, fea=128, batch 40 per GPU, sub_batch 5), one rank per GPU.

  python bench.py --gpus N --steps K --warmup W            (N>1 under torchrun)
  python bench.py --impl reference ...                     CPU restatement of the reference graph (oracle port)

One "step" == one optimiser step of model/train_single_gpu.py:138-150: sub_batch micro-batches of
forward+backward (BRN batch statistics, dropout, loss, all gradients), ONE gradient all-reduce across ranks,
clip +-0.2, Adam.  `value` = crops/s with inputs resident in HBM; `e2e` = the same step through the public
Python API with pinned HOST inputs (H2D inside the timed region) and the loss read back (D2H).
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

# SURVEY.md 8d / BASELINE.md section 2: algorithmic conv FLOPs per crop, training = 3 x forward
TRAIN_GFLOP_PER_CROP = {16: 29.37, 14: 29.19, 21: 29.82}
METRIC = "depth-crops/sec (128x128, 2-stack fea=128) training step"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel (conv on s0/um_comb/c2, B=40), from the
# `ncu --set full` captures summarised in profiles/r1_final.md (tf32x3 = the CTA-pair kernel: 46.8 MB read + 4.3 MB written) and
# profiles/r1_tensor_core_path.md (algorithmic: 42 MB in + 42 MB out + 2.4 MB weights; the output is still L2-resident when the capture ends)
NCU_TRAFFIC_BYTES = {"tf32x3": 51.1e6, "tf32": 47.4e6, "fp32": 51.0e6}


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
def cpu_reference_step(net, U, params, state, m, v, batch, step, J):
    dms, poses, cfgs, coms = batch
    L, g, _ = U.loss_and_grads(net, params, state, dms[..., 0], poses, cfgs, coms, dropout_seed=step)
    U.adam_step(params, g, m, v, step=step, lr=1e-3, accum_steps=1, world=1)
    return L["total"]


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


def _cpu_worker(conn, steps, warmup, target_s, max_batch):
    """Child process: CPU restatement of the reference graph (oracle port).  Reports after every timed step so that the
    parent can enforce a wall-clock box and still use what was measured."""
    try:
        import torch
        from oracle import um_v1_torch as U
        from densereg_b200 import synth
        J = 16
        threads = usable_cpus()
        torch.set_num_threads(threads)
        net = U.Net(2, 128, J)
        p, s = net.init_params(0), net.init_state()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        state = {"n": 0}

        def step(batch):
            state["n"] += 1
            return cpu_reference_step(net, U, p, s, m, v, batch, state["n"], J)

        one = synth.make_batch(1, J, seed=0)
        step(one)                                             # primitive creation / first touch
        t0 = time.perf_counter(); step(one); t1 = time.perf_counter() - t0
        bsz = int(max(1, min(max_batch, round(target_s / max(t1, 1e-3)))))
        sample = ("fwd+bwd+Adam on one micro-batch of %d crop(s) per step (instead of 5x40), %d threads "
                  "(usable CPUs of %d reported; 1-crop calibration step %.2f s)" % (bsz, threads, os.cpu_count() or 0, t1))
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


def run_cpu_reference(steps, warmup, wall_s, target_s=3.0, max_batch=8):
    """-> dict(value, steps, dt, threads, sample, complete) or None.  Never takes longer than wall_s seconds."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe(duplex=False)
    proc = ctx.Process(target=_cpu_worker, args=(child, steps, warmup, target_s, max_batch), daemon=True)
    proc.start()
    deadline = time.time() + wall_s
    last, complete = None, False
    while time.time() < deadline:
        if parent.poll(0.5):
            msg = parent.recv()
            if msg.get("phase") == "done":
                complete = True
                break
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
        if not complete:
            last["sample"] += " [stopped by the %d s wall-clock box after %d timed step(s)]" % (int(wall_s), last.get("steps", 0))
    return last


def run_reference(args):
    """--impl reference: each step is a bounded sample of the workload (one micro-batch sized by calibration); the whole run is
    boxed to a few minutes."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    r = run_cpu_reference(args.steps, args.warmup, wall_s=float(os.environ.get("DENSEREG_REF_WALL_S", "240")))
    if r is None or r.get("value") is None:
        r = {"value": 0.0, "steps": 0, "dt": 0.0, "threads": usable_cpus(), "sample": "CPU restatement produced no step inside the time box"}
    val = r["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (r["dt"] / max(r["steps"], 1)) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "ICVL J=16 2-stack fea=128 training step, CPU restatement of the reference TF graph (PyTorch-CPU fp32)",
                   "sample": r["sample"], "timed_steps": r["steps"]},
        "cpu_baseline": {"value": val, "unit": "crops/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": val, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch_size", type=int, default=40)
    ap.add_argument("--sub_batch", type=int, default=5)
    ap.add_argument("--precision", type=str, default="tf32x3", choices=["fp32", "tf32", "tf32x3"])
    ap.add_argument("--no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from densereg_b200.engine import DenseRegEngine
    from densereg_b200 import synth
    from densereg_b200.model import allreduce_gradients

    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    S, F, J, B, SUB = 2, 128, 16, args.batch_size, args.sub_batch
    eng = DenseRegEngine(S, F, J, max_batch=B, precision=args.precision, device=local, training=True)
    eng.init_params(seed=0)                       # same seed on every rank -> identical replicas without a broadcast
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
            eng.loss_backward(d, po, cf, co, dropout_seed=i * SUB + sub)
        allreduce_gradients(eng.grads, world)
        eng.optimizer_step(i + 1, 1e-3, accum_steps=SUB, world=world)

    def step_e2e(i):
        eng.zero_grads()
        for sub in range(SUB):
            d, po, cf, co = [t.to(dev, non_blocking=True) for t in pinned[(i * SUB + sub) % NROT]]
            loss = eng.loss_backward(d, po, cf, co, dropout_seed=i * SUB + sub)
        allreduce_gradients(eng.grads, world)
        eng.optimizer_step(i + 1, 1e-3, accum_steps=SUB, world=world)
        loss_host.copy_(loss, non_blocking=True)

    def timed(fn, first):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count
        e0.record()
        for i in range(args.steps):
            fn(first + i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), eng.launch_count - l0

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed(step_resident, args.warmup)
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    crops = B * SUB * args.steps * world
    value = crops / (ms * 1e-3)
    value_e2e = crops / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: the conv implicit GEMM on its largest layer (s*/um_comb/c2, 3x3 256->256 @32x32,
    #      12.3 % of the MACs per stack), timed alone with CUDA events; inputs rotated over > L2 worth of buffers ----------
    peaks = measured_peaks()
    names = [l["name"] for l in eng.layers()]
    li = names.index("s0/um_comb/c2")
    xs = [torch.randn(B, 32, 32, 256, device=dev) for _ in range(4)]       # 4 x 42 MB inputs + outputs > 126 MB L2
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
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    step_tflops = value / world * TRAIN_GFLOP_PER_CROP[J] / 1e3
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_burst"], "traffic": NCU_TRAFFIC_BYTES.get(args.precision),
                "kernel": "conv implicit-GEMM (%s path%s) on s0/um_comb/c2 3x3 256->256, B=%d"
                          % (args.precision, ", tcgen05 cta_group::2 CTA pairs" if args.precision == "tf32x3" and os.environ.get("DENSEREG_TC_PAIR", "1") != "0" else "", B),
                "kernel_ms": k_ms, "peak_source": peaks["src"] + " dense bf16 cuBLAS burst (tf32 kind nominally half)",
                "whole_step": {"achieved": step_tflops, "peak": peaks["bf16_sustained"], "frac": step_tflops / peaks["bf16_sustained"],
                               "note": "29.37 GFLOP/crop (fwd+dgrad+wgrad conv FLOPs) x crops/s per GPU vs sustained measured peak"}}

    if rank == 0:
        cpu_baseline = None
        if not args.no_cpu_baseline and world == 1:
            r = run_cpu_reference(steps=4, warmup=0, wall_s=float(os.environ.get("DENSEREG_CPU_WALL_S", "75")), target_s=4.0)
            if r is not None and r.get("value") is not None:
                cpu_baseline = {"value": r["value"], "unit": "crops/s", "cores": r["threads"], "kind": "port",
                                "sample": "%d timed step(s): %s" % (r["steps"], r["sample"])}
            else:
                cpu_baseline = {"value": None, "unit": "crops/s", "cores": usable_cpus(), "kind": "port",
                                "sample": "CPU restatement did not finish a step inside the 75 s box"}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "fp32", "tf32": "tf32", "tf32x3": "fp32 (3xTF32 split)"}[args.precision], "data": "synthetic",
            "config": {"workload": "ICVL J=16 2-stack fea=128 training: optimiser step = %d micro-batches x batch %d per GPU, fwd+bwd+allreduce+clip+Adam"
                                   % (SUB, B), "global_batch": B * world, "sub_batch": SUB, "parallelism": "dp%d" % world,
                       "l2": "working set ~%.1f GB per micro-batch >> 126 MB L2; inputs rotated over %d batches" % (eng.workspace_bytes / 2**30, NROT)},
            "e2e": {"value": value_e2e, "unit": "crops/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 20,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
