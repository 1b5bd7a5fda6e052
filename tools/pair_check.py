#!/usr/bin/env python
"""Bring-up check of the experimental CTA-pair (tcgen05 cta_group::2) 3xTF32 conv kernel (densereg_b200/csrc/conv_tc_pair.cu)
against the default one-CTA kernel: per-layer agreement, kernel timing, whole micro-step agreement and timing.
Every result is appended to gpurun_out/pair_check.jsonl as soon as it exists (run it under `timeout`: a protocol bug shows up as a hang).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "pair_check.jsonl")
os.makedirs(os.path.dirname(OUT), exist_ok=True)


def emit(**kw):
    kw["t"] = round(time.time() - T0, 2)
    with open(OUT, "a") as f:
        f.write(json.dumps(kw) + "\n")
    print(json.dumps(kw), flush=True)


T0 = time.time()
from densereg_b200.engine import DenseRegEngine  # noqa: E402
from densereg_b200 import synth  # noqa: E402

stage = sys.argv[1] if len(sys.argv) > 1 else "all"
emit(step="start", stage=stage)

if stage in ("all", "conv"):
    eng = DenseRegEngine(2, 128, 16, max_batch=40, precision="tf32x3", training=False)
    eng.init_params(0, 0.05)
    L = eng.layers(); names = [l["name"] for l in L]
    cases = [("s0/um_comb/c2", 40), ("s0/um_full2", 40), ("s0/um_res2/c2", 40), ("s0/um_res1/skip", 40), ("s0/um_comb/c3", 3),
             ("s0/um_out", 7), ("s0/hg/n2/lower3/c1", 5), ("s0/hg/n3/upper1/c2", 3), ("s0/um_full1", 2)]
    for name, B in cases:
        li = names.index(name); l = L[li]
        g = torch.Generator(device="cuda").manual_seed(li)
        x = torch.randn(B, l["in_hw"], l["in_hw"], l["cin"], device="cuda", generator=g)
        y0 = eng.debug_conv(li, x, "tf32x3")
        torch.cuda.synchronize()
        emit(step="conv_ref_done", layer=name, B=B)
        t0 = eng.tc_launch_count
        y1 = eng.debug_conv(li, x, "tf32x3", reuse_weights=True, pair=True)
        torch.cuda.synchronize()
        d = (y1 - y0).abs().max().item(); sc = y0.abs().max().item()
        emit(step="conv_pair", layer=name, B=B, k=l["k"], cin=l["cin"], cout=l["cout"], hw=l["in_hw"], max_abs_diff=d, scale=sc, rel=d / max(sc, 1e-30),
             finite=bool(torch.isfinite(y1).all().item()), tc_launches=eng.tc_launch_count - t0)

    # kernel timing on the roofline layer, inputs rotated over > L2
    for name in ("s0/um_comb/c2", "s0/um_full2", "s0/um_res2/c2"):
        li = names.index(name); l = L[li]; B = 40
        xs = [torch.randn(B, l["in_hw"], l["in_hw"], l["cin"], device="cuda") for _ in range(4)]
        res = {}
        for pair in (False, True):
            y = eng.debug_conv(li, xs[0], "tf32x3", pair=pair)
            for x in xs:
                eng.debug_conv(li, x, "tf32x3", reuse_weights=True, out=y, pair=pair)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(12):
                eng.debug_conv(li, xs[r % 4], "tf32x3", reuse_weights=True, out=y, pair=pair)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 12
            fl = 2.0 * B * l["out_hw"] ** 2 * l["k"] ** 2 * l["cin"] * l["cout"]
            res["pair" if pair else "single"] = dict(ms=ms, tflops=fl / ms / 1e9)
        emit(step="conv_time", layer=name, **res)
    eng.close(); del eng

if stage in ("all", "step"):
    B, J = 40, 16
    d, po, cf, co = [torch.from_numpy(a).cuda() for a in synth.make_batch(B, J, seed=0)]
    out = {}
    for pair in (False, True):
        eng = DenseRegEngine(2, 128, J, max_batch=B, precision="tf32x3", training=True, tc_pair=pair)
        eng.init_params(0)
        eng.zero_grads()
        loss = eng.loss_backward(d, po, cf, co, dropout_seed=1).clone()
        torch.cuda.synchronize()
        xyz = eng.infer(d, cf, co).clone()
        grads = eng.grads.clone()
        # timing: 3 micro-steps
        for _ in range(2):
            eng.loss_backward(d, po, cf, co, dropout_seed=2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(4):
            eng.loss_backward(d, po, cf, co, dropout_seed=3 + i)
        e1.record(); torch.cuda.synchronize()
        out[pair] = dict(loss=loss.cpu(), xyz=xyz.cpu(), grads=grads.cpu(), ms=e0.elapsed_time(e1) / 4)
        emit(step="micro_step", pair=pair, ms=out[pair]["ms"], crops_per_s=B / out[pair]["ms"] * 1e3, loss=[float(v) for v in out[pair]["loss"]])
        eng.close(); del eng
    a, b = out[False], out[True]
    gd = (a["grads"] - b["grads"]).abs().max().item(); gs = a["grads"].abs().max().item()
    emit(step="step_agreement", loss_rel=float(((a["loss"] - b["loss"]).abs() / a["loss"].abs().clamp_min(1e-30)).max()),
         grad_max_abs_diff=gd, grad_scale=gs, xyz_max_abs_diff_mm=float((a["xyz"] - b["xyz"]).abs().max()),
         speedup=a["ms"] / b["ms"])
emit(step="done")
