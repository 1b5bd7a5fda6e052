#!/usr/bin/env python
"""Filter-gradient (wgrad) timing of single layers through dr_debug_conv_bwd (dw only), B=40: the largest item of the training step after
round 1 (profiles/r1_final.md section 5).  python tools/time_wgrad.py [--precision tf32x3]"""
import argparse, json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from densereg_b200.engine import DenseRegEngine
ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="tf32x3"); ap.add_argument("--batch", type=int, default=40); ap.add_argument("--reps", type=int, default=8)
a = ap.parse_args()
B = a.batch
eng = DenseRegEngine(2, 128, 16, max_batch=B, precision=a.precision, training=False)
eng.init_params(0, 0.05)
L = eng.layers(); names = [l["name"] for l in L]
LAYERS = ["s0/um_comb/c2", "s0/um_res2/c2", "s0/um_full1", "s0/um_res1/c2", "s0/um_full2", "s0/hg/n4/upper1/c2", "s0/um_res2/c3", "s0/um_res1/skip",
          "s0/um_res2/c1", "stem/conv_2/c2", "s0/um_res1/c3", "s0/hg/n4/upper1/c3", "s0/um_comb/c3", "s0/um_comb/c1", "s0/hg/n3/upper1/c2",
          "s0/hg/n1/upper1/c2", "s0/hm3_res/c2", "s0/um_out"]
for name in LAYERS:
    li = names.index(name); l = L[li]
    x = torch.randn(B, l["in_hw"], l["in_hw"], l["cin"], device="cuda")
    dy = torch.randn(B, l["out_hw"], l["out_hw"], l["cout"], device="cuda")
    for _ in range(2):
        eng.debug_conv_bwd(li, x, dy, a.precision, want_dx=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        eng.debug_conv_bwd(li, x, dy, a.precision, want_dx=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps                    # includes the memset of dw (tiny)
    fl = 2.0 * B * l["out_hw"] ** 2 * l["k"] ** 2 * l["cin"] * l["cout"]
    byts = 4.0 * B * l["out_hw"] ** 2 * (l["cin"] + l["cout"])
    print(json.dumps(dict(layer=name, hw=l["in_hw"], k=l["k"], cin=l["cin"], cout=l["cout"], us=round(ms * 1e3, 1), tflops=round(fl / ms / 1e9, 1),
                          hbm_gbs_min=round(byts / ms / 1e6, 0), cin_rows_used=round(l["cin"] / (math.ceil(l["cin"] / 128) * 128), 2))), flush=True)
