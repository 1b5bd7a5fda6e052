"""smoke(): small invocations of the hot path on cuda:0, checked against the CPU oracle: forward + vote and one training micro-step on the
1-stack / 64-feature net, then one training micro-step of the benchmarked configuration (2-stack, 128 features, batch 32) so that the
tcgen05 kernels bench.py times (one-CTA and CTA-pair conv, one-CTA and CTA-pair wgrad) are exercised here too -- with the micro-batch
pipeline on, like bench.py: a second micro-batch of the same crops runs in the second arena and must double the accumulated gradient.
Everything runs in the library's default arithmetic (3xTF32 on the tensor cores)."""
import numpy as np
import torch


def run():
    from .engine import DenseRegEngine
    from . import synth
    from oracle import um_v1_torch as U, vote_numpy as V   # checker only
    S, F, J, B = 1, 64, 16, 2
    eng = DenseRegEngine(S, F, J, max_batch=B, training=True)
    net = U.Net(S, F, J)
    p, s = net.init_params(0, stddev=0.05), net.init_state()
    eng.load_flat(p, s)
    dms, poses, cfgs, coms = synth.make_batch(B, J, seed=1)
    cu = lambda a: torch.from_numpy(a).cuda()
    d, po, cf, co = cu(dms), cu(poses), cu(cfgs), cu(coms)
    # inference: crops -> xyz mm
    xyz = eng.infer(d, cf, co).cpu().numpy()
    x0n = V.norm_dm(dms[..., 0], coms)
    hms, hm3s, ums = net.forward(p, s, torch.from_numpy(x0n[..., None]), training=False)
    out = eng.forward(d, co)
    e_map = float((out["um_outs"][-1].cpu() - ums[-1]).abs().max() / ums[-1].abs().max())
    # vote on the GPU's own maps vs oracle vote on the same maps: indices exact, xyz <= 1e-3 mm
    hm_g, hm3_g, um_g = (out[k][-1].cpu().numpy() for k in ("hm_outs", "hm3_outs", "um_outs"))
    ref_xyz, ref_top5 = V.xyz_estimation(hm_g, hm3_g, um_g, V.tiny_dm(x0n), cfgs, coms)
    ok = np.isfinite(ref_xyz)
    e_xyz = float(np.abs(xyz - ref_xyz)[ok].max())
    # one training micro-step
    eng.zero_grads()
    loss = eng.loss_backward(d, po, cf, co, dropout_seed=3).cpu().numpy()
    L, g_ref, _ = U.loss_and_grads(net, p, s.clone(), dms[..., 0], poses, cfgs, coms, dropout_seed=3)
    e_loss = abs(loss[0] - L["total"]) / abs(L["total"])
    e_grad = float((eng.grads.cpu() - g_ref).norm() / g_ref.norm())
    eng.optimizer_step(step=1, lr=1e-3)
    torch.cuda.synchronize()
    print("smoke: map relerr %.2e | vote xyz err %.2e mm | loss relerr %.2e | grad relerr %.2e | launches %d"
          % (e_map, e_xyz, e_loss, e_grad, eng.launch_count))
    # gradient bar: 2e-2 -- two fp32 evaluations of this graph differ by ~3e-3 (ReLU/BRN sign flips; see tests/test_gpu_net.py)
    assert e_map < 1e-4 and e_xyz <= 1e-3 and e_loss < 1e-4 and e_grad < 2e-2
    assert eng.tc_launch_count > 0, "the tensor-core path did not run"
    eng.close()
    # ---- the benchmarked configuration: 2-stack fea=128, batch 32 (>= 64 work items per big layer -> CTA-pair kernel)
    S, F, J, B = 2, 128, 16, 32
    eng = DenseRegEngine(S, F, J, max_batch=B, training=True, precision="tf32x3", pipeline=2)
    net = U.Net(S, F, J)
    p, s = net.init_params(0, stddev=0.05), net.init_state()
    eng.load_flat(p, s)
    dms, poses, cfgs, coms = synth.make_batch(B, J, seed=2)
    d, po, cf, co = cu(dms), cu(poses), cu(cfgs), cu(coms)
    eng.zero_grads()
    tc0 = eng.tc_launch_count
    loss = eng.loss_backward(d, po, cf, co, dropout_seed=4).cpu().numpy()
    L, g_ref, _ = U.loss_and_grads(net, p, s.clone(), dms[..., 0], poses, cfgs, coms, dropout_seed=4)
    e_loss = abs(loss[0] - L["total"]) / abs(L["total"])
    e_grad = float((eng.grads.cpu() - g_ref).norm() / g_ref.norm())
    n_tc = eng.tc_launch_count - tc0
    # second micro-batch in the pipeline's second arena: same crops, same dropout seed, BRN state put back to what the first one read (its
    # update moved d_max from 0 to 1e-3, ops.py:146-149) and not updated again -> the same forward pass, so the accumulated gradient doubles
    g1 = eng.grads.clone()
    eng.state.copy_(s.to(eng.device))
    loss2 = eng.loss_backward(d, po, cf, co, dropout_seed=4, update_state=False).cpu().numpy()
    e_twin = float((eng.grads - 2.0 * g1).norm() / g1.norm())
    print("smoke: 2x128 B=%d 3xTF32 micro-step: loss relerr %.2e | grad relerr %.2e | tensor-core launches %d of %d | pipelined second micro-batch: "
          "grad(2) - 2 grad(1) = %.2e, loss diff %.2e" % (B, e_loss, e_grad, n_tc, eng.launch_count, e_twin, abs(loss2[0] - loss[0]) / abs(loss[0])))
    assert e_loss < 1e-4 and e_grad < 2e-2 and n_tc > 300
    assert eng.pipeline_depth == 2 and e_twin < 1e-3 and abs(loss2[0] - loss[0]) <= 1e-5 * abs(loss[0])
