"""Micro-batch pipeline (dr_config.reserved[2] == 2, include/densereg.h dr_pipeline_join): the forward pass of micro-batch i+1 runs next to
the backward pass of micro-batch i on a second activation arena.  Reference semantics (model/train_single_gpu.py:138-150): sub_batch
sequential `sess.run([loss, accum_op])` calls, then ONE apply.  The pipelined engine must give the sequential engine's results: the same
losses, the same BRN state sequence (forward passes stay in micro-batch order), the same accumulated gradients up to the order of the
fp32 atomics of the filter-gradient kernels, hence the same parameters after the update."""
import numpy as np
import pytest
import torch

from gpu_util import cu, dump

pytestmark = pytest.mark.gpu


def _run(pipeline, S, F, J, B, micro, steps, precision):
    from densereg_b200.engine import DenseRegEngine
    from densereg_b200 import synth
    eng = DenseRegEngine(S, F, J, max_batch=B, training=True, precision=precision, pipeline=pipeline)
    assert eng.pipeline_depth == pipeline
    eng.init_params(seed=3, stddev=0.05)
    batches = [[cu(a) for a in synth.make_batch(B, J, seed=100 + i)] for i in range(micro)]
    losses, grads = [], None
    for step in range(steps):
        eng.zero_grads()
        for i, (d, po, cf, co) in enumerate(batches):
            losses.append(eng.loss_backward(d, po, cf, co, dropout_seed=step * micro + i).clone())    # stream-ordered after this micro-batch's loss
        if step == 0:
            eng.join()                                   # the caller reads the gradient buffer itself
            grads = eng.grads.clone()
        eng.optimizer_step(step + 1, 1e-3, accum_steps=micro)
    torch.cuda.synchronize()
    out = dict(loss=torch.stack(losses).cpu().numpy(), grads=grads.cpu().numpy(), params=eng.params.cpu().numpy(), state=eng.state.cpu().numpy(),
               launches=eng.launch_count, ws=eng.workspace_bytes)
    eng.close()
    return out


@pytest.mark.parametrize("S,F,J,B,precision", [(1, 64, 16, 4, "fp32"), (2, 128, 16, 8, "tf32x3")])
def test_pipeline_matches_sequential(built_lib, S, F, J, B, precision):
    micro, steps = 5, 2
    seq = _run(1, S, F, J, B, micro, steps, precision)
    pip = _run(2, S, F, J, B, micro, steps, precision)
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b.astype(np.float64)) + 1e-30))
    rep = dict(loss=float(np.abs(pip["loss"] - seq["loss"]).max() / np.abs(seq["loss"]).max()), grads=rel(pip["grads"], seq["grads"]),
               params=rel(pip["params"], seq["params"]), state=rel(pip["state"], seq["state"]),
               launches=(pip["launches"], seq["launches"]), workspace=(pip["ws"], seq["ws"]))
    dump("pipeline_vs_sequential_S%dF%dJ%d_%s.json" % (S, F, J, precision), rep)
    assert np.isfinite(pip["loss"]).all() and np.isfinite(pip["params"]).all()
    # step 1 is identical up to the order of fp32 / fp64 atomics; step 2 starts from parameters that differ by that noise times Adam's
    # sign-like first updates, so the bars are those of two runs of the SAME sequential engine
    assert rep["loss"] < 1e-4, rep
    assert rep["grads"] < 1e-3, rep
    assert rep["state"] < 1e-4, rep
    assert rep["params"] < 1e-3, rep
    assert pip["launches"] == seq["launches"], rep                  # the same kernels ran, half of them in the second arena
    assert pip["ws"] > 1.5 * seq["ws"], rep


def test_pipeline_forward_order_and_join(built_lib):
    """BRN moving statistics after k micro-batches are those of k sequential forward passes (ops.py:141-162 update order), and entry points
    that read shared buffers join the pipeline themselves: dr_forward right after a pipelined micro-batch sees the final state."""
    from densereg_b200.engine import DenseRegEngine
    from densereg_b200 import synth
    S, F, J, B = 1, 64, 16, 4
    res = {}
    for pipeline in (1, 2):
        eng = DenseRegEngine(S, F, J, max_batch=B, training=True, precision="fp32", pipeline=pipeline)
        eng.init_params(seed=5, stddev=0.05)
        eng.zero_grads()
        for i in range(3):
            d, po, cf, co = [cu(a) for a in synth.make_batch(B, J, seed=7 + i)]
            eng.loss_backward(d, po, cf, co, dropout_seed=i)
        o = eng.forward(d, co, is_training=False)                  # eval-mode BRN reads the moving statistics: must see all three updates
        torch.cuda.synchronize()
        res[pipeline] = (eng.state.cpu().numpy(), o["um_outs"][-1].cpu().numpy(), eng.grads.cpu().numpy())
        eng.close()
    assert float(np.abs(res[2][0] - res[1][0]).max() / np.abs(res[1][0]).max()) < 1e-5
    assert float(np.abs(res[2][1] - res[1][1]).max() / np.abs(res[1][1]).max()) < 1e-4
    assert float(np.linalg.norm(res[2][2] - res[1][2]) / np.linalg.norm(res[1][2])) < 1e-3
