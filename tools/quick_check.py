#!/usr/bin/env python
"""One-process check of the training micro-step under the current environment switches: gradients / loss / outputs of the 3xTF32
tensor-core engine against the fp32 FFMA engine of the same library on the same inputs (GPU vs GPU: seconds, no CPU oracle), then
CUDA-event timing of N micro-batches.  Prints one JSON line.  Used by tools/r2_sweep.py (one child process per switch setting,
because the switches are read once per process).
  python tools/quick_check.py [--batch 40] [--micro 6] [--ref_cache gpurun_out/qc_ref.pt] [--tag name]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=40); ap.add_argument("--micro", type=int, default=6)
ap.add_argument("--precision", default="tf32x3"); ap.add_argument("--ref_cache", default=os.path.join(ROOT, "gpurun_out", "qc_ref.pt"))
ap.add_argument("--tag", default="base"); ap.add_argument("--J", type=int, default=16); ap.add_argument("--no_parity", action="store_true")
a = ap.parse_args()
from densereg_b200.engine import DenseRegEngine
from densereg_b200 import synth
B, J = a.batch, a.J
out = {"tag": a.tag, "B": B, "env": {k: v for k, v in os.environ.items() if k.startswith("DENSEREG_")}}
try:
    data = [torch.from_numpy(x).cuda() for x in synth.make_batch(B, J, seed=77)]
    eng = DenseRegEngine(2, 128, J, max_batch=B, precision=a.precision, training=True)
    eng.init_params(0, 0.05)
    if not a.no_parity:
        key = "B%d_J%d" % (B, J)
        ref = torch.load(a.ref_cache) if os.path.exists(a.ref_cache) else {}
        if key not in ref:
            r = DenseRegEngine(2, 128, J, max_batch=B, precision="fp32", training=True)
            r.load_flat(eng.params, eng.state)
            r.zero_grads()
            l = r.loss_backward(*data, dropout_seed=5, update_state=False).clone()
            o = r.forward(data[0], data[3], is_training=True, update_state=False, dropout_seed=5)
            ref[key] = dict(grads=r.grads.cpu(), loss=l.cpu(), um=o["um_outs"][-1].cpu(), hm=o["hm_outs"][-1].cpu(),
                            xyz=r.infer(data[0], data[2], data[3]).cpu())
            torch.save(ref, a.ref_cache)
            r.close(); del r; torch.cuda.empty_cache()
        ref = ref[key]
        eng.zero_grads()
        l = eng.loss_backward(*data, dropout_seed=5, update_state=False).cpu()
        o = eng.forward(data[0], data[3], is_training=True, update_state=False, dropout_seed=5)
        g, gr = eng.grads.cpu(), ref["grads"]
        worst, wname = 0.0, None
        for L in eng.layers():
            n = L["k"] * L["k"] * L["cin"] * L["cout"]
            sl = slice(L["w_off"], L["w_off"] + n)
            e = float((g[sl] - gr[sl]).norm() / (gr[sl].norm() + 1e-20))
            if e > worst:
                worst, wname = e, L["name"]
        out.update(loss_rel=float((l[0] - ref["loss"][0]).abs() / ref["loss"][0].abs()),
                   grad_total_rel=float((g - gr).norm() / gr.norm()), grad_worst_rel=worst, grad_worst_layer=wname,
                   um_rel=float((o["um_outs"][-1].cpu() - ref["um"]).abs().max() / ref["um"].abs().max()),
                   hm_rel=float((o["hm_outs"][-1].cpu() - ref["hm"]).abs().max() / ref["hm"].abs().max()),
                   finite=bool(torch.isfinite(g).all()))
        xyz = eng.infer(data[0], data[2], data[3]).cpu()
        d = (xyz - ref["xyz"]).abs(); d = d[torch.isfinite(d)]
        out.update(xyz_max_mm=float(d.max()), xyz_mean_mm=float(d.mean()), xyz_p99_mm=float(d.flatten().kthvalue(max(1, int(0.99 * d.numel())))[0]))
        out["parity_ok"] = bool(out["finite"] and out["loss_rel"] < 1e-3 and out["grad_worst_rel"] < 5e-2 and out["um_rel"] < 1e-3
                                and out["xyz_max_mm"] < 2e-2)          # inference at the full batch (several tiles per CTA) against the fp32 engine
    # ---- timing: full optimiser steps of `micro` micro-batches
    def step(i):
        eng.zero_grads()
        for s in range(a.micro):
            eng.loss_backward(*data, dropout_seed=i * a.micro + s)
        eng.optimizer_step(i + 1, 1e-3, accum_steps=a.micro)
    step(0); step(1)
    torch.cuda.synchronize()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    reps = 2
    for i in range(reps):
        step(2 + i)
    e1.record()
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * a.micro)
    # host cost of enqueueing ONE micro-batch into an empty queue (no back-pressure from the GPU)
    torch.cuda.synchronize(); t0 = time.perf_counter(); eng.loss_backward(*data, dropout_seed=99); t_one = time.perf_counter() - t0
    torch.cuda.synchronize()
    # per-class conv-type time of one micro-batch, every launch timed alone (serialised)
    eng.trace(True); eng.loss_backward(*data, dropout_seed=98, update_state=False); torch.cuda.synchronize()
    cls = {}
    for r in eng.trace_records():
        c = cls.setdefault(r["kind"] + ":" + r["kernel"], [0, 0.0]); c[0] += 1; c[1] += r["ms"]
    eng.trace(False)
    lay = {}
    for r in eng.trace_records():
        k = "%s hw%d k%d %d->%d %s" % (r["kind"], r["hw"], r["k"], r["cin"], r["cout"], r["kernel"])
        c = lay.setdefault(k, [0, 0.0, 2.0 * r["B"] * r["hw"] ** 2 * r["k"] ** 2 * r["cin"] * r["cout"]]); c[0] += 1; c[1] += r["ms"]
    out["trace_layers"] = [[k, v[0], round(v[1], 4), round(v[2] * v[0] / max(v[1], 1e-9) / 1e9, 1)] for k, v in sorted(lay.items(), key=lambda kv: -kv[1][1])[:60]]
    out["trace_ms"] = {k: [v[0], round(v[1], 3)] for k, v in sorted(cls.items())}
    out["cpu_enqueue_one_micro_ms"] = t_one * 1e3
    out.update(ms_per_micro=ms, crops_per_s=B / ms * 1e3, launches_per_micro=(eng.launch_count - l0) / (reps * a.micro),
               cpu_enqueue_ms_per_micro=t_cpu * 1e3 / (reps * a.micro), tc_launches=eng.tc_launch_count)
except Exception as e:  # noqa: BLE001
    out["error"] = repr(e)[:600]
print(json.dumps(out), flush=True)
