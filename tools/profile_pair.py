#!/usr/bin/env python
"""ncu target: one launch of the one-CTA and one of the CTA-pair 3xTF32 conv kernel on the same layer (after a warm-up of each).
  ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 2 --launch-count 2 -o gpurun_out/pair python tools/profile_pair.py"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--layer", default="s0/um_comb/c2"); ap.add_argument("--batch", type=int, default=40)
a = ap.parse_args()
from densereg_b200.engine import DenseRegEngine
eng = DenseRegEngine(2, 128, 16, max_batch=a.batch, precision="tf32x3", training=False)
eng.init_params(0, 0.05)
L = eng.layers(); li = [l["name"] for l in L].index(a.layer); l = L[li]
x = torch.randn(a.batch, l["in_hw"], l["in_hw"], l["cin"], device="cuda")
y = eng.debug_conv(li, x, "tf32x3")                                   # warm-up (also builds the split weight copies)
eng.debug_conv(li, x, "tf32x3", reuse_weights=True, out=y, pair=True)
torch.cuda.synchronize()
eng.debug_conv(li, x, "tf32x3", reuse_weights=True, out=y)            # profiled: one-CTA
eng.debug_conv(li, x, "tf32x3", reuse_weights=True, out=y, pair=True) # profiled: CTA pair
torch.cuda.synchronize()
print("done")
