"""Micro-batch pipeline, host side, WITHOUT a GPU.  dr_loss_backward / dr_pipeline_join of a pipelined handle (include/densereg.h) execute lists
of stream operations -- record / wait / forward pass / backward pass on the caller's stream and on the two slot streams -- that the library
builds as data (engine.cu: pipe_plan_micro_batch, pipe_plan_join, pipe_choose_slot).  dr_debug_pipeline_plan exports the same lists in a dry run.
This test replays them on a model of CUDA's stream / event semantics (a stream is a FIFO; cudaStreamWaitEvent waits for the event's most
recent record AT THE TIME OF THE CALL, none = no-op) with random pass durations and random call sequences, and checks what the reference's
training loop (model/train_single_gpu.py:138-150: sequential sess.run([loss, accum_op]) per micro-batch, then one apply) needs:

  1. forward passes run in micro-batch order (BRN moving statistics, network/slim/ops.py:141-162);
  2. backward passes never overlap (exclusive accumulation into the gradient buffer, train_single_gpu.py:69-84);
  3. a pass starts after everything the caller enqueued before the call (input copies, zero_grads, the optimiser step);
  4. when the call returns, the caller's stream is ordered behind that micro-batch's loss (loss_out valid; inputs no longer read);
  5. dr_pipeline_join / dr_zero_grads / dr_optimizer_step order the caller's stream behind every backward pass;
  6. a pass that rebuilds the weight copies (first after an optimiser step) or all-reduces gradient buckets runs in slot 0, the former only
     after the second arena's last pass has finished;
  7. the point of it all: forward(i+1) is NOT ordered behind backward(i) -- with short forward passes it starts before backward(i) ends.
"""
import ctypes as C
import random

import pytest

RECORD, WAIT, FORWARD, BACKWARD = 0, 1, 2, 3
CALLER = -1


@pytest.fixture()
def plan(built_lib):
    from densereg_b200 import _ffi
    lib = built_lib
    cfg = _ffi.DrConfig(num_stack=1, num_fea=64, kernel_size=3, num_jnt=16, in_hw=128, out_hw=32, max_batch=4, precision=2, device=0)
    cfg.reserved[2] = 2
    h = C.c_void_p()
    assert lib.dr_create(C.byref(h), C.byref(cfg)) == 0 and lib.dr_pipeline_depth(h) == 2
    buf = (_ffi.DrPipeOp * 16)()

    def call(what):
        n = lib.dr_debug_pipeline_plan(h, what, buf, 16)
        assert n >= 0, lib.dr_last_error(h)
        return [(buf[i].kind, buf[i].stream, buf[i].event) for i in range(n)]
    yield call
    lib.dr_destroy(h)


class Sim:
    """Times of stream operations under CUDA semantics.  Ops are issued in host order; every dependency points to an earlier op, so one pass
    in issue order fixes all start / end times."""

    def __init__(self):
        self.tail = {}          # stream -> end time of its last op
        self.latest = {}        # event -> completion time of its most recent record
        self.passes = []        # dicts: kind, slot, mb, start, end, loss (backward only)

    def op(self, stream, dur=0.0, dep=None):
        start = max(self.tail.get(stream, 0.0), dep if dep is not None else 0.0)
        self.tail[stream] = start + dur
        return start, start + dur

    def run(self, ops, mb, t_fwd, t_bwd, t_loss):
        slot = None
        for kind, stream, event in ops:
            if kind == RECORD:
                self.latest[event] = self.op(stream)[1]
            elif kind == WAIT:
                self.op(stream, dep=self.latest.get(event))          # never recorded: no-op
            elif kind == FORWARD:
                s, e = self.op(stream, t_fwd)
                slot = stream
                self.passes.append(dict(kind="fwd", slot=stream, mb=mb, start=s, end=e))
            elif kind == BACKWARD:
                s, e = self.op(stream, t_bwd)
                self.latest[event] = s + t_loss                      # recorded right after the loss kernels
                self.passes.append(dict(kind="bwd", slot=stream, mb=mb, start=s, end=e, loss=s + t_loss))
        return slot


def drive(plan, seed, fwd_range, bwd_range, steps=6):
    """Random optimiser steps; returns (sim, log) where log has per micro-batch: slot, dirty, armed, caller time before / after the call."""
    rng = random.Random(seed)
    sim, log, mb = Sim(), [], 0
    dirty = True                                                     # nothing prepared yet
    for step in range(steps):
        sim.op(CALLER, 0.01)                                         # dr_zero_grads (its join emitted nothing: nothing pending after an optimiser step)
        assert plan(1) == []
        sub = rng.randint(1, 6)
        for i in range(sub):
            armed = (i == sub - 1) and rng.random() < 0.5
            if armed:
                plan(3)
            joined = rng.random() < 0.15 and i > 0
            if joined:                                               # a caller that reads the gradients mid-step
                ops = plan(1)
                sim.run(ops, None, 0, 0, 0)
                assert all(p["end"] <= sim.tail[CALLER] + 1e-12 for p in sim.passes if p["kind"] == "bwd")
            sim.op(CALLER, rng.uniform(0.0, 0.3))                    # the caller's input copy for this micro-batch
            before = sim.tail[CALLER]
            ops = plan(0)
            t_f, t_b = rng.uniform(*fwd_range), rng.uniform(*bwd_range)
            slot = sim.run(ops, mb, t_f, t_b, 0.05 * t_b)
            log.append(dict(mb=mb, slot=slot, dirty=dirty, armed=armed, joined=joined, before=before, after=sim.tail[CALLER], ops=ops))
            dirty = False
            mb += 1
        ops = plan(2)                                                # dr_optimizer_step: join + parameters changed
        sim.run(ops, None, 0, 0, 0)
        assert all(p["end"] <= sim.tail[CALLER] + 1e-12 for p in sim.passes)           # (5)
        sim.op(CALLER, 0.05)                                         # the Adam kernel
        dirty = True
    return sim, log


@pytest.mark.parametrize("seed", range(12))
def test_pipeline_plan_orders_what_the_reference_needs(plan, seed):
    sim, log = drive(plan, seed, fwd_range=(0.5, 3.0), bwd_range=(0.5, 4.0))
    fwd = sorted((p for p in sim.passes if p["kind"] == "fwd"), key=lambda p: p["mb"])
    bwd = sorted((p for p in sim.passes if p["kind"] == "bwd"), key=lambda p: p["mb"])
    assert len(fwd) == len(bwd) == len(log)
    eps = 1e-12
    for a, b in zip(fwd, fwd[1:]):
        assert b["start"] >= a["end"] - eps                                           # (1)
    for a, b in zip(bwd, bwd[1:]):
        assert b["start"] >= a["end"] - eps                                           # (2)
    for f, b, l in zip(fwd, bwd, log):
        assert f["slot"] == b["slot"] == l["slot"] and b["start"] >= f["end"] - eps
        assert f["start"] >= l["before"] - eps                                        # (3)
        assert l["after"] >= b["loss"] - eps                                          # (4): loss written, inputs consumed (forward + loss kernels)
        if l["dirty"] or l["armed"]:
            assert l["slot"] == 0                                                     # (6)
        if l["dirty"]:
            others = [p["end"] for p in sim.passes if p["slot"] == 1 and p["mb"] < l["mb"]]
            assert not others or f["start"] >= max(others) - eps
    # same arena: a pass starts after the previous micro-batch of that slot has left it
    for slot in (0, 1):
        seq = sorted((p for p in sim.passes if p["slot"] == slot), key=lambda p: (p["mb"], p["kind"] == "bwd"))
        for a, b in zip(seq, seq[1:]):
            assert b["start"] >= a["end"] - eps
    # inside a step the slots alternate unless a pass is pinned to slot 0 (a join empties the pipeline: the next pass starts over in slot 0)
    for a, b in zip(log, log[1:]):
        if b["joined"]:
            assert b["slot"] == 0
        elif not (b["dirty"] or b["armed"]):
            assert b["slot"] == 1 - a["slot"]


def test_pipeline_plan_lets_forward_overlap_backward(plan):
    """(7) with forward passes much shorter than backward passes, forward(i+1) starts while backward(i) is still running -- and a step of n
    micro-batches takes about fwd + n * bwd instead of n * (fwd + bwd)."""
    sim, log = drive(plan, 3, fwd_range=(1.0, 1.0), bwd_range=(3.0, 3.0), steps=4)
    fwd = {p["mb"]: p for p in sim.passes if p["kind"] == "fwd"}
    bwd = {p["mb"]: p for p in sim.passes if p["kind"] == "bwd"}
    overlapped = sequential = 0
    for a, b in zip(log, log[1:]):
        if b["dirty"] or b["joined"]:
            continue                                                  # first pass after an optimiser step / a join: nothing left to overlap with
        if b["slot"] != a["slot"]:
            assert fwd[b["mb"]]["start"] < bwd[a["mb"]]["end"] - 0.5, (a, b)
            overlapped += 1
        else:
            sequential += 1                                           # pinned to slot 0 right after a slot-0 pass: same stream
    assert overlapped >= 3 and overlapped > sequential


def test_pipeline_plan_refuses_live_or_unpipelined_handles(built_lib):
    from densereg_b200 import _ffi
    lib = built_lib
    cfg = _ffi.DrConfig(num_stack=1, num_fea=64, kernel_size=3, num_jnt=16, in_hw=128, out_hw=32, max_batch=4, precision=2, device=0)
    h = C.c_void_p()
    assert lib.dr_create(C.byref(h), C.byref(cfg)) == 0
    buf = (_ffi.DrPipeOp * 16)()
    assert lib.dr_debug_pipeline_plan(h, 0, buf, 16) == -3 and b"pipeline" in lib.dr_last_error(h)      # depth 1
    assert lib.dr_debug_pipeline_plan(h, 0, buf, 4) == -1 and lib.dr_debug_pipeline_plan(h, 7, buf, 16) == -1
    assert lib.dr_debug_pipeline_plan(None, 0, buf, 16) == -1
    assert lib.dr_destroy(h) == 0
