#!/bin/bash
# Round-2 GPU call 21: micro-batch pipeline (forward of micro-batch i+1 next to backward of i): parity tests, A/B at batch 40 and batch 8, per-shape trace.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x > gpurun_out/c21_pytest.log 2>&1
SWEEP_ARGS="--micro 5" timeout -s KILL 420 python tools/r2_sweep.py base pipe2 pipe2_prio pipe2_wgrad_a_tmem pipe2_no_lanes chain_prio > gpurun_out/c21_sweep_b40.log 2>&1
SWEEP_ARGS="--batch 8 --J 14 --micro 5" timeout -s KILL 200 python tools/r2_sweep.py base pipe2 pipe2_prio > gpurun_out/c21_sweep_b8.log 2>&1
timeout -s KILL 120 python tools/trace_shapes.py --batch 40 > gpurun_out/trace_shapes_b40.md 2> gpurun_out/trace_shapes_b40.err
tail -15 gpurun_out/c21_pytest.log | cut -c1-400; cut -c1-330 gpurun_out/c21_sweep_b40.log; cut -c1-330 gpurun_out/c21_sweep_b8.log; head -30 gpurun_out/trace_shapes_b40.md | cut -c1-200
