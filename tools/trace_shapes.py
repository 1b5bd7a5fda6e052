#!/usr/bin/env python
"""Where the conv-type time of a training micro-batch goes, by layer SHAPE: one traced pass (every conv / dgrad / wgrad launch timed alone with CUDA
events, dr_trace), grouped by (kind, hw, cin, cout, k, kernel) -> launches, total ms, TFLOP/s, share.  Prints a markdown table sorted by time.
  python tools/trace_shapes.py [--batch 40] [--J 16] > gpurun_out/trace_shapes_b40.md"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=40); ap.add_argument("--J", type=int, default=16); ap.add_argument("--precision", default="tf32x3")
a = ap.parse_args()
from densereg_b200.engine import DenseRegEngine
from densereg_b200 import synth
eng = DenseRegEngine(2, 128, a.J, max_batch=a.batch, precision=a.precision, training=True)
eng.init_params(0, 0.05)
data = [torch.from_numpy(x).cuda() for x in synth.make_batch(a.batch, a.J, seed=3)]
eng.zero_grads()
for i in range(2):
    eng.loss_backward(*data, dropout_seed=i)
torch.cuda.synchronize()
eng.trace(True)
eng.loss_backward(*data, dropout_seed=7)
torch.cuda.synchronize()
recs = eng.trace_records()
eng.trace(False)
g = {}
for r in recs:
    key = (r["kind"], r["hw"], r["cin"], r["cout"], r["k"], r["kernel"])
    e = g.setdefault(key, [0, 0.0, 0.0])
    e[0] += 1; e[1] += r["ms"]; e[2] += 2.0 * r["B"] * r["hw"] ** 2 * r["k"] ** 2 * r["cin"] * r["cout"]
tot = sum(e[1] for e in g.values())
print("conv-type launches of one micro-batch, B=%d J=%d %s: %d launches, %.2f ms if serialised\n" % (a.batch, a.J, a.precision, len(recs), tot))
print("| kind | hw | cin | cout | k | kernel | launches | total ms | share | us / launch | TFLOP/s |")
print("|---|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|")
for key, e in sorted(g.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %d | %d | %d | %s | %d | %.3f | %.1f %% | %.1f | %.0f |" % (*key, e[0], e[1], 100 * e[1] / tot, 1e3 * e[1] / e[0], e[2] / (e[1] * 1e-3) / 1e12))
for kind in ("conv", "dgrad", "wgrad"):
    for lo, hi in ((32, 32), (16, 16), (1, 8), (64, 64)):
        sel = [(k, e) for k, e in g.items() if k[0] == kind and lo <= k[1] <= hi]
        if sel:
            ms = sum(e[1] for _, e in sel); fl = sum(e[2] for _, e in sel)
            print("\n%s at %d..%d px: %d launches, %.3f ms, %.0f TFLOP/s" % (kind, lo, hi, sum(e[0] for _, e in sel), ms, fl / (ms * 1e-3) / 1e12), end="")
print()
