#!/usr/bin/env python
"""Bring-up script for the tcgen05 conv path: layer by layer vs torch (GPU fp32 reference for speed), prints as it goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from densereg_b200.engine import DenseRegEngine
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = DenseRegEngine(2, 128, 14, max_batch=B, training=False)
eng.init_params(0, 0.1)
L = eng.layers()
names = sys.argv[3].split(",") if len(sys.argv) > 3 else ["s0/hg/n4/upper1/c1", "s0/hg/n4/upper1/c2", "s0/um_comb/c2", "s0/um_full2", "s0/hm_out",
        "s0/um_res1/c1", "s0/um_res1/c2", "stem/conv_2/c2", "s0/hg/n1/lower1/c2", "s0/hg/n2/lower3/c1", "s0/hg/n3/upper1/c3", "s0/um_out"]
for name in names:
    li = [l["name"] for l in L].index(name); l = L[li]
    k, cin, cout, hw = l["k"], l["cin"], l["cout"], l["in_hw"]
    x = torch.randn(B, hw, hw, cin, device="cuda")
    w = eng.params[l["w_off"]:l["w_off"] + k * k * cin * cout].view(k, k, cin, cout)
    ref = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1).contiguous()
    print("%-22s k%d %3d->%3d @%2d ..." % (name, k, cin, cout, hw), end="", flush=True)
    t0 = eng.tc_launch_count
    y = eng.debug_conv(li, x, prec)
    torch.cuda.synchronize()
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    dy = torch.randn_like(ref)
    t1 = eng.tc_launch_count
    dx, dw = eng.debug_conv_bwd(li, x, dy, prec)
    torch.cuda.synchronize()
    nbw = eng.tc_launch_count - t1
    xr = x.permute(0, 3, 1, 2)
    dwr = torch.nn.grad.conv2d_weight(xr, (cout, cin, k, k), dy.permute(0, 3, 1, 2).contiguous(), padding=k // 2).permute(2, 3, 1, 0).reshape(-1)
    werr = (dw - dwr).abs().max().item() / dwr.abs().max().item()
    dref = F.conv_transpose2d(dy.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
    derr = (dx - dref).abs().max().item() / dref.abs().max().item()
    print(" fwd %.2e  dgrad %.2e  wgrad %.2e  tc fwd/bwd launches %d/%d" % (err, derr, werr, t1 - t0, nbw), flush=True)
    if err > 1e-2:
        d = (y - ref).abs()
        idx = torch.nonzero(d > 1e-2 * ref.abs().max())[:5]
        print("   first bad (b,y,x,c):", idx.tolist(), " y:", y.flatten()[:6].tolist(), " ref:", ref.flatten()[:6].tolist(), flush=True)
