#!/usr/bin/env python
"""Offset-vote microbench (BASELINE.json configs[4]): B x H x W heat-map + 3-D offset maps, J joints.
Reports GB/s on the algorithmic-bytes convention 4*H*W*(5J+1) per sample vs the measured HBM
copy peak, next to the NumPy oracle timed on a bounded sample."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--jnt", type=int, default=21)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--cpu_samples", type=int, default=2)
    a = ap.parse_args()
    from densereg_b200.engine import DenseRegEngine
    eng = DenseRegEngine(1, 64, 16, max_batch=1, training=False)
    B, H, J = a.batch, a.hw, a.jnt
    g = torch.Generator(device="cuda").manual_seed(0)
    hm = torch.rand(B, H, H, J, device="cuda", generator=g) * 1.2 - 0.1
    hm3 = torch.rand(B, H, H, J, device="cuda", generator=g).clamp_(0.05, 1)
    um = torch.randn(B, H, H, 3 * J, device="cuda", generator=g)
    dmn = torch.where(torch.rand(B, H, H, device="cuda", generator=g) < 0.6, torch.full((), -1.0, device="cuda"),
                      torch.rand(B, H, H, device="cuda", generator=g) * 1.3 - 0.4)
    cfgs = torch.tensor([[240., 240., 64., 64., 128., 128.]], device="cuda").repeat(B, 1)
    coms = torch.tensor([[0., 0., 400.]], device="cuda").repeat(B, 1)
    for _ in range(3):
        eng.vote(hm, hm3, um, dmn, cfgs, coms)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        xyz = eng.vote(hm, hm3, um, dmn, cfgs, coms)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    bytes_alg = 4.0 * H * H * (5 * J + 1) * B
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    out = {"workload": "vote B=%d %dx%d J=%d" % (B, H, H, J), "ms": ms, "samples_per_s": B / (ms * 1e-3),
           "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": bytes_alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        "note": "algorithmic bytes (5J+1)*4*H*W per sample; the kernel only gathers um at the 5 winners per joint, "
                                "so its real DRAM traffic is ~(2J+1)/(5J+1) of this"}}
    if a.cpu_samples:
        from oracle import vote_numpy as V
        n = a.cpu_samples
        args = [t[:n].cpu().numpy() for t in (hm, hm3, um, dmn, cfgs, coms)]
        t0 = time.perf_counter(); ref, _ = V.xyz_estimation(*args); dt = time.perf_counter() - t0
        ok = np.isfinite(ref)
        out["cpu_baseline"] = {"value": n / dt, "unit": "samples/s", "cores": 1, "kind": "port", "sample": "%d samples, NumPy oracle" % n}
        out["max_err_mm_vs_oracle"] = float(np.abs(xyz[:n].cpu().numpy() - ref)[ok].max())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
