#!/usr/bin/env python
"""Where the end-to-end training step loses time against the resident one: the same optimiser step (5 micro-batches, batch 40) timed in
several variants and orders inside one process.  python tools/e2e_probe.py [--pipeline 2]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--pipeline", type=int, default=2); ap.add_argument("--batch", type=int, default=40); ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
from densereg_b200.engine import DenseRegEngine
from densereg_b200 import synth
B, J, SUB, NROT, NST = a.batch, 16, 5, 4, 3
dev = torch.device("cuda", 0)
eng = DenseRegEngine(2, 128, J, max_batch=B, precision="tf32x3", training=True, pipeline=a.pipeline)
eng.init_params(0)
pinned = [[torch.from_numpy(x).pin_memory() for x in synth.make_batch(B, J, seed=i)] for i in range(NROT)]
resident = [[t.to(dev) for t in hb] for hb in pinned]
staging = [[torch.empty_like(t, device=dev) for t in pinned[0]] for _ in range(NST)]
copy_stream = torch.cuda.Stream(device=dev)
ev_ready = [torch.cuda.Event() for _ in range(NST)]; ev_consumed = [torch.cuda.Event() for _ in range(NST)]
loss_host = torch.zeros(5).pin_memory()
state = {"n": -1}

def prefetch(n, do_copy=True):
    if n <= state["n"]:
        return
    state["n"] = n
    j = n % NST
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(ev_consumed[j])
        if do_copy:
            for x, y in zip(staging[j], pinned[n % NROT]):
                x.copy_(y, non_blocking=True)
        ev_ready[j].record(copy_stream)

def make_step(mode):
    def step(i):
        cur = torch.cuda.current_stream()
        eng.zero_grads()
        if mode in ("e2e", "e2e_nod2h", "events_only"):
            prefetch(i * SUB, mode != "events_only")
        for sub in range(SUB):
            n = i * SUB + sub
            if mode in ("e2e", "e2e_nod2h", "events_only"):
                prefetch(n + 1, mode != "events_only")
                cur.wait_event(ev_ready[n % NST])
                loss = eng.loss_backward(*staging[n % NST], dropout_seed=n)
                ev_consumed[n % NST].record(cur)
            elif mode == "same_stream":
                d = [t.to(dev, non_blocking=True) for t in pinned[n % NROT]]
                loss = eng.loss_backward(*d, dropout_seed=n)
            else:
                loss = eng.loss_backward(*resident[n % NROT], dropout_seed=n)
        eng.optimizer_step(i + 1, 1e-3, accum_steps=SUB)
        if mode in ("e2e", "resident_d2h", "same_stream"):
            loss_host.copy_(loss, non_blocking=True)
    return step

def timed(fn, first, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(first + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

out = {"pipeline": eng.pipeline_depth, "B": B}
k = 0
for mode in ("resident", "e2e", "resident", "e2e_nod2h", "events_only", "resident_d2h", "same_stream", "resident"):
    fn = make_step(mode)
    for w in range(2):
        fn(k); k += 1
    ms = timed(fn, k, a.steps); k += a.steps
    out.setdefault(mode, []).append(round(ms, 3))
print(json.dumps(out))
