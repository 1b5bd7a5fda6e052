"""Generates the committed golden vectors with the CPU oracle (the reference itself -- TF 1.3 / py2 --
cannot run here; see oracle/__init__.py: PARITY UNPINNED).  Run from the repo root:
    python tests/golden/make_golden.py
Fixtures: vote_J{16,14,21}.npz (inputs + oracle xyz/top5), net_S1F64J16.npz (sub-sampled oracle outputs),
train_S1F64J16.npz (loss values + gradient norms of one micro-batch)."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vote_numpy as V, um_v1_torch as U   # noqa: E402
from densereg_b200 import synth                        # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

for J, seed in [(16, 10), (14, 11), (21, 12)]:
    hm, hm3, um, dmn, cfgs, coms = synth.make_vote_maps(3, J, hw=32, seed=seed)
    xyz, top5 = V.xyz_estimation(hm, hm3, um, dmn, cfgs, coms)
    np.savez_compressed(os.path.join(OUT, "vote_J%d.npz" % J), hm=hm.astype(np.float16).astype(np.float32) if False else hm,
                        hm3=hm3, um=um, dmn=dmn, cfgs=cfgs, coms=coms, xyz=xyz, top5=top5)

net = U.Net(1, 64, 16)
seed, stddev, data_seed = 3, 0.05, 5
p, s = net.init_params(seed, stddev=stddev), net.init_state()
dms, poses, cfgs, coms = synth.make_batch(1, 16, seed=data_seed)
x0 = torch.from_numpy(V.norm_dm(dms[..., 0], coms)[..., None])
hms, hm3s, ums = net.forward(p, s, x0, training=False)
np.savez_compressed(os.path.join(OUT, "net_S1F64J16.npz"), seed=seed, stddev=stddev, data_seed=data_seed,
                    hm_sub=hms[0].numpy()[0, ::4, ::4], um_sub=ums[0].numpy()[0, ::4, ::4])
dms, poses, cfgs, coms = synth.make_batch(2, 16, seed=data_seed)
L, g, _ = U.loss_and_grads(net, p, s.clone(), dms[..., 0], poses, cfgs, coms, dropout_seed=9)
norms = np.array([float(g[c.w_off:c.w_off + c.k * c.k * c.cin * c.cout].norm()) for c in net.specs], np.float64)
np.savez_compressed(os.path.join(OUT, "train_S1F64J16.npz"), seed=seed, stddev=stddev, data_seed=data_seed, dropout_seed=9,
                    loss=np.array([L["total"], L["hm"], L["hm3"], L["um"], L["reg"]]), grad_norms=norms)
print("golden vectors written to", OUT)
