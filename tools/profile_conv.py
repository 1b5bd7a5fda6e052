#!/usr/bin/env python
"""Run one conv layer of the table (forward, dgrad, wgrad) a few times -- target for `ncu -k regex:conv`."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--layer", default="s0/um_comb/c2")
ap.add_argument("--batch", type=int, default=40)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--bwd", action="store_true")
a = ap.parse_args()
from densereg_b200.engine import DenseRegEngine
eng = DenseRegEngine(2, 128, 16, max_batch=a.batch, training=False)
eng.init_params(0, 0.05)
L = eng.layers(); li = [l["name"] for l in L].index(a.layer); l = L[li]
x = torch.randn(a.batch, l["in_hw"], l["in_hw"], l["cin"], device="cuda")
dy = torch.randn(a.batch, l["out_hw"], l["out_hw"], l["cout"], device="cuda")
for _ in range(a.iters):
    y = eng.debug_conv(li, x, a.precision)
    if a.bwd:
        eng.debug_conv_bwd(li, x, dy, a.precision)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y = eng.debug_conv(li, x, a.precision)
for _ in range(10):
    eng.debug_conv(li, x, a.precision, reuse_weights=True, out=y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2.0 * a.batch * l["out_hw"] ** 2 * l["k"] ** 2 * l["cin"] * l["cout"]
print("%s B=%d %s: %.3f ms  %.1f TFLOP/s" % (a.layer, a.batch, a.precision, ms, fl / ms / 1e9))
