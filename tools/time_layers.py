#!/usr/bin/env python
"""Eval-mode (plain scale/shift/ReLU epilogue) timing of single conv layers through dr_debug_conv: separates per-tile epilogue cost
from main-loop cost.  python tools/time_layers.py [--pair]"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from densereg_b200.engine import DenseRegEngine
B = 40
pair = "--pair" in sys.argv
eng = DenseRegEngine(2, 128, 16, max_batch=B, precision="tf32x3", training=False)
eng.init_params(0, 0.05)
L = eng.layers(); names = [l["name"] for l in L]
for name in ["s0/um_res1/skip", "s0/um_res1/c3", "s0/um_res2/c3", "s0/um_res2/c1", "s0/hg/n4/upper1/c3", "s0/hg/n4/upper1/c1", "s0/ll", "s0/um_comb/c3",
             "s0/um_comb/c1", "s0/um_full2", "s0/um_res2/c2", "s0/um_comb/c2", "s0/hg/n4/upper1/c2"]:
    li = names.index(name); l = L[li]
    xs = [torch.randn(B, l["in_hw"], l["in_hw"], l["cin"], device="cuda") for _ in range(3)]
    y = eng.debug_conv(li, xs[0], "tf32x3", pair=pair)
    for x in xs:
        eng.debug_conv(li, x, "tf32x3", reuse_weights=True, out=y, pair=pair)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(12):
        eng.debug_conv(li, xs[r % 3], "tf32x3", reuse_weights=True, out=y, pair=pair)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    BN = min((l["cout"] + 15) // 16 * 16, 128)
    tiles = math.ceil(B * l["out_hw"] ** 2 / 128) * math.ceil(l["cout"] / BN)
    rounds = math.ceil(tiles / 148); kb = l["k"] ** 2 * math.ceil(l["cin"] / 32)
    fl = 2.0 * B * l["out_hw"] ** 2 * l["k"] ** 2 * l["cin"] * l["cout"]
    byts = 4.0 * B * l["out_hw"] ** 2 * (l["cin"] + l["cout"])
    print(json.dumps(dict(layer=name, k=l["k"], cin=l["cin"], cout=l["cout"], us=round(ms * 1e3, 1), tiles=tiles, rounds=rounds, kb=kb,
                          us_per_round=round(ms * 1e3 / rounds, 2), tflops=round(fl / ms / 1e9, 1), hbm_gbs=round(byts / ms / 1e6, 0))), flush=True)
