#!/usr/bin/env python
"""Inference throughput (BASELINE.json configs[3]: MSRA15 21-joint, 2-stack fea=128, eval-mode BRN folded into the conv epilogue,
forward + vote -> xyz mm), batch sweep, with the mean joint error vs the CPU oracle on a small sample."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="tf32x3"); ap.add_argument("--jnt", type=int, default=21)
ap.add_argument("--batches", default="1,8,64,256"); ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--check", type=int, default=2, help="crops compared with the CPU oracle (0 = skip)")
a = ap.parse_args()
from densereg_b200.engine import DenseRegEngine
from densereg_b200 import synth
Bs = [int(x) for x in a.batches.split(",")]
eng = DenseRegEngine(2, 128, a.jnt, max_batch=max(Bs), precision=a.precision, training=False, infer_graph=bool(a.graph))
eng.init_params(0, 0.05)
out = {"workload": "MSRA-shape J=%d 2-stack fea=128 inference (forward + vote), %s, cuda_graph=%d" % (a.jnt, a.precision, a.graph), "rows": []}
for B in Bs:
    dms, poses, cfgs, coms = synth.make_batch(min(B, 64), a.jnt, seed=B)
    rep = (B + dms.shape[0] - 1) // dms.shape[0]
    hd = [torch.from_numpy(np.concatenate([x] * rep)[:B]).pin_memory() for x in (dms, cfgs, coms)]
    d, cf, co = [t.cuda() for t in hd]
    xyz = torch.empty(B, 3 * a.jnt, device="cuda")
    for _ in range(3):
        eng.infer(d, cf, co, out=xyz)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        eng.infer(d, cf, co, out=xyz)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    # end to end: pinned host -> device -> xyz -> host
    res = torch.empty(B, 3 * a.jnt).pin_memory()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.iters):
        d.copy_(hd[0], non_blocking=True); cf.copy_(hd[1], non_blocking=True); co.copy_(hd[2], non_blocking=True)   # same device buffers
        eng.infer(d, cf, co, out=xyz)
        res.copy_(xyz, non_blocking=True)
    torch.cuda.synchronize(); ms_e2e = (time.perf_counter() - t0) * 1e3 / a.iters
    out["rows"].append({"batch": B, "ms": ms, "crops_per_s": B / ms * 1e3, "e2e_crops_per_s": B / ms_e2e * 1e3,
                        "conv_tflops": B / ms * 1e3 * 9.939e9 / 1e12})
if a.check:
    from oracle import um_v1_torch as U, vote_numpy as V
    n = a.check
    dms, poses, cfgs, coms = synth.make_batch(n, a.jnt, seed=77)
    net = U.Net(2, 128, a.jnt)
    p = eng.params.cpu(); s = eng.state.cpu()
    x0n = V.norm_dm(dms[..., 0], coms)
    hms, hm3s, ums = net.forward(p, s, torch.from_numpy(x0n[..., None]), training=False)
    ref, _ = V.xyz_estimation(hms[-1].numpy(), hm3s[-1].numpy(), ums[-1].numpy(), V.tiny_dm(x0n), cfgs, coms)
    got = eng.infer(*[torch.from_numpy(x).cuda() for x in (dms, cfgs, coms)]).cpu().numpy()
    ok = np.isfinite(ref) & np.isfinite(got)
    err = np.linalg.norm((got - ref).reshape(n, a.jnt, 3), axis=-1)
    out["mean_joint_err_mm_vs_oracle"] = float(np.nanmean(err)); out["max_joint_err_mm_vs_oracle"] = float(np.nanmax(err))
print(json.dumps(out))
