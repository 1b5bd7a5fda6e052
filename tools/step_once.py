#!/usr/bin/env python
"""Run N training micro-batches (fwd+bwd) + one optimiser step -- a light target for `ncu` launch lists."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="tf32x3"); ap.add_argument("--batch", type=int, default=40); ap.add_argument("--micro", type=int, default=2)
ap.add_argument("--infer", action="store_true")
a = ap.parse_args()
from densereg_b200.engine import DenseRegEngine
from densereg_b200 import synth
eng = DenseRegEngine(2, 128, 16, max_batch=a.batch, precision=a.precision, training=not a.infer)
eng.init_params(0)
d, po, cf, co = [torch.from_numpy(x).cuda() for x in synth.make_batch(a.batch, 16, seed=0)]
if a.infer:
    for i in range(a.micro):
        eng.infer(d, cf, co)
else:
    eng.zero_grads()
    for i in range(a.micro):
        eng.loss_backward(d, po, cf, co, dropout_seed=i)
    eng.optimizer_step(1, 1e-3, accum_steps=a.micro)
torch.cuda.synchronize()
print("launches", eng.launch_count, "tc", eng.tc_launch_count)
