#!/usr/bin/env python
"""Run ONE layer's conv / wgrad a few times through the per-layer debug entry points (isolated kernel, B=40) -- the target of the
per-kernel `ncu --set full` captures in tools/ncu_kernels.sh.   python tools/profile_layer.py --layer s0/um_comb/c2 --what fwd|fwd_chunk|wgrad|dgrad"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--layer", default="s0/um_comb/c2"); ap.add_argument("--what", default="fwd"); ap.add_argument("--batch", type=int, default=40)
ap.add_argument("--reps", type=int, default=8)
a = ap.parse_args()
from densereg_b200.engine import DenseRegEngine
from densereg_b200 import _ffi
eng = DenseRegEngine(2, 128, 16, max_batch=a.batch, precision="tf32x3", training=False)
eng.init_params(0, 0.05)
L = eng.layers(); li = [l["name"] for l in L].index(a.layer); l = L[li]
xs = [torch.randn(a.batch, l["in_hw"], l["in_hw"], l["cin"], device="cuda") for _ in range(4)]        # rotate > L2 worth of inputs
dy = torch.randn(a.batch, l["out_hw"], l["out_hw"], l["cout"], device="cuda")
y = None
for r in range(a.reps):
    x = xs[r % 4]
    if a.what == "fwd":
        y = eng.debug_conv(li, x, "tf32x3", reuse_weights=r > 0, out=y)
    elif a.what == "fwd_chunk":
        import ctypes as C
        y = y if y is not None else torch.empty(a.batch, l["out_hw"], l["out_hw"], l["cout"], device="cuda")
        eng._check(eng.lib.dr_debug_conv(eng._h, li, a.batch, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()),
                                         _ffi.PRECISIONS["tf32x3"] | (0x100 if r > 0 else 0) | 0x400, eng._stream()))
    elif a.what == "wgrad":
        eng.debug_conv_bwd(li, x, dy, "tf32x3", want_dx=False)
    else:
        eng.debug_conv_bwd(li, x, dy, "tf32x3", want_dx=True)
torch.cuda.synchronize()
print("done", a.layer, a.what)
