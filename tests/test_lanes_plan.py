"""CPU test of the lane scheduler (engine.cu: Builder::plan_lanes / OpPlan).  The library runs independent branches of um_v1 on separate CUDA
streams; which events an op waits for is decided at build time by a hazard analysis.  Here every hazard is re-derived by brute force from the op
table (dr_debug_op) -- any two ops on DIFFERENT lanes that touch overlapping channel ranges of one buffer, at least one of them writing -- and the
plan must order each such pair: the later op has to be reachable from the earlier one through "same lane, earlier" and "waits for" edges.  Also
checked: every wait refers to an op that records an event and comes earlier in the pass, lanes really contain what network/um_v1.py allows to run
in parallel (hourglass `upper1` blocks :54-65, the masked um branch :143-149, projection skips :31-47), and nothing else was moved off lane 0."""
import ctypes as C

import pytest

from densereg_b200 import _ffi

CONV, POOL, UPADD, MASKCOPY = 0, 1, 2, 3


def _ops(lib, S, F, J):
    cfg = _ffi.DrConfig(num_stack=S, num_fea=F, kernel_size=3, num_jnt=J, in_hw=128, out_hw=32, max_batch=2, precision=2, device=0)
    h = C.c_void_p()
    assert lib.dr_create(C.byref(h), C.byref(cfg)) == 0
    ops = []
    for i in range(lib.dr_num_ops(h)):
        o = _ffi.DrOpInfo()
        assert lib.dr_debug_op(h, i, C.byref(o)) == 0
        name = ""
        if o.kind == CONV:
            li = _ffi.DrLayerInfo(); lib.dr_get_layer(h, o.layer, C.byref(li)); name = li.name.decode()
        ops.append(dict(i=i, kind=o.kind, lane=o.lane, name=name, need_dgrad=o.need_dgrad, raw=o.raw_buf,
                        inn=(o.in_buf, o.in_c0, o.in_c0 + o.in_c), out=(o.out_buf, o.out_c0, o.out_c0 + o.out_c),
                        res=(o.res_buf, o.res_c0, o.res_c0 + o.res_c) if o.res_buf >= 0 else None,
                        gsrc=(o.gsrc_buf, o.gsrc_c0, o.gsrc_c0 + o.gsrc_c) if o.gsrc_buf >= 0 else None,
                        dres=(o.dres_buf, o.dres_c0, o.dres_c0 + o.dres_c) if o.dres_buf >= 0 else None,
                        res_grad_fused=o.res_grad_fused, in_grad_fused=o.in_grad_fused,
                        waits=[[o.wait_op[p][k] for k in range(o.nwait[p])] for p in (0, 1)], record=[o.record[0], o.record[1]]))
    lib.dr_destroy(h)
    return ops


def _accesses(o, backward):
    """[(arena, buffer, c0, c1, is_write)] of one op in one pass (forward: activations; backward: gradients)."""
    if not backward:
        acc = [("a", *o["inn"], False), ("a", *o["out"], True)]
        if o["res"]:
            acc.append(("a", *o["res"], False))
        if o["raw"] >= 0:
            acc.append(("a", o["raw"], 0, 1 << 20, True))
        return acc
    acc = [("g", *(o["gsrc"] or o["out"]), False)]           # gradient aliasing: d(out) may be read from another view
    if o["kind"] == CONV:
        if o["res"] and not o["res_grad_fused"]:
            acc.append(("g", *o["res"], True))
        if o["dres"]:
            acc.append(("g", *o["dres"], False))
        if o["need_dgrad"]:
            acc.append(("g", *o["inn"], True))
    else:
        if not (o["kind"] == UPADD and o["in_grad_fused"]):
            acc.append(("g", *o["inn"], True))
        if o["kind"] == UPADD:
            acc.append(("g", *o["res"], True))
    return acc


@pytest.mark.parametrize("S,F,J", [(2, 128, 16), (1, 64, 16), (2, 128, 21)])
def test_every_cross_lane_hazard_is_ordered(built_lib, S, F, J):
    ops = _ops(built_lib, S, F, J)
    n = len(ops)
    for p, backward in ((0, False), (1, True)):
        order = list(range(n)) if not backward else list(range(n - 1, -1, -1))
        pos = {oi: k for k, oi in enumerate(order)}
        # happens-before: reach[k] = set of positions known to complete before position k starts
        last_on_lane, reach = {}, []
        for k, oi in enumerate(order):
            o = ops[oi]
            before = set()
            if o["lane"] in last_on_lane:
                q = last_on_lane[o["lane"]]; before |= reach[q] | {q}
            for w in o["waits"][p]:
                assert ops[w]["record"][p] == 1, (oi, w)            # the awaited op records its event ...
                assert pos[w] < k and ops[w]["lane"] != o["lane"]   # ... was enqueued earlier, on another lane
                before |= reach[pos[w]] | {pos[w]}
            reach.append(before)
            last_on_lane[o["lane"]] = k
        hazards = 0
        for k2 in range(n):
            a2 = _accesses(ops[order[k2]], backward)
            for k1 in range(k2):
                if ops[order[k1]]["lane"] == ops[order[k2]]["lane"]:
                    continue
                a1 = _accesses(ops[order[k1]], backward)
                clash = any(x[0] == y[0] and x[1] == y[1] and x[2] < y[3] and y[2] < x[3] and (x[4] or y[4]) for x in a1 for y in a2)
                if clash:
                    hazards += 1
                    assert k1 in reach[k2], "pass %d: op %d (%s) must wait for op %d (%s)" % (p, order[k2], ops[order[k2]]["name"], order[k1], ops[order[k1]]["name"])
        assert hazards > 20                                          # the test really exercised cross-lane dependencies


def test_lane_assignment_follows_the_graph(built_lib):
    ops = _ops(built_lib, 2, 128, 16)
    off0 = [o for o in ops if o["lane"] != 0]
    names = [o["name"] for o in off0 if o["kind"] == CONV]
    assert all(("/upper1/" in nm) or ("um_mask_res" in nm) or nm.endswith("/skip") for nm in names), names
    assert sum("/upper1/" in nm for nm in names) == 2 * 4 * 3                     # 4 hourglass levels x 3 convs per stack
    assert sum("um_mask_res" in nm for nm in names) == 2 * 7                      # two residual blocks (one with a projection skip) per stack
    assert [o["kind"] for o in off0 if o["kind"] != CONV] == [MASKCOPY, MASKCOPY]  # tf.where mask copy starts the masked branch
    assert {o["lane"] for o in ops} == {0, 1, 2}
    lower = [o for o in ops if "/lower" in o["name"] or o["kind"] in (POOL, UPADD)]
    assert lower and all(o["lane"] == 0 for o in lower)


@pytest.mark.parametrize("S,F,J", [(2, 128, 16), (1, 64, 14)])
def test_backward_reads_only_written_gradients(built_lib, S, F, J):
    """Gradient aliasing removed the copies of d(residual sum) (engine.cu: Op::gsrc / dres / *_grad_fused).  Walk the reverse schedule and check
    that every gradient view an op reads has been completely written before -- by the loss kernel (the outputs hm / hm3 / um of every stack) or by
    ops earlier in the walk -- and that the fused copies really are gone (no op writes the gradient of a skip buffer or of `up1`)."""
    ops = _ops(built_lib, S, F, J)
    written = {}
    def mark(buf, c0, c1):
        written.setdefault(buf, set()).update(range(c0, c1))
    for o in ops:
        if o["kind"] == CONV and o["name"].endswith(("/hm_out", "/hm3_out", "/um_out")):
            mark(*o["out"])                                              # dL/d(output maps): loss_kernel
    n_alias = 0
    for o in reversed(ops):
        for (_, buf, c0, c1, is_write) in _accesses(o, True):
            if not is_write:
                missing = set(range(c0, c1)) - written.get(buf, set())
                assert not missing, "op %d (%s) reads an unwritten gradient: buffer %d channels %s" % (o["i"], o["name"], buf, sorted(missing)[:4])
        for (_, buf, c0, c1, is_write) in _accesses(o, True):
            if is_write:
                mark(buf, c0, c1)
        n_alias += int(o["gsrc"] is not None) + int(o["dres"] is not None)
    n_blocks = sum(1 for o in ops if o["name"].endswith("/c3"))
    assert n_alias >= n_blocks and all(o["res_grad_fused"] for o in ops if o["kind"] == CONV and o["name"].endswith("/c3"))
    assert all(o["in_grad_fused"] for o in ops if o["kind"] == UPADD)
