#!/bin/bash
# Round-2 GPU call 17: forward BRN statistics from the transposed tile; BRN passes walking the maps end to start (L2 reuse); full tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/c17_pytest.log 2>&1
timeout -s KILL 400 python tools/r2_sweep.py base ew_reverse_0 > gpurun_out/c17_sweep_b40.log 2>&1
SWEEP_ARGS="--batch 8 --J 14" timeout -s KILL 300 python tools/r2_sweep.py base ew_reverse_0 > gpurun_out/c17_sweep_b8.log 2>&1
SWEEP_ARGS="--batch 64 --J 14" timeout -s KILL 300 python tools/r2_sweep.py base ew_reverse_0 > gpurun_out/c17_sweep_b64.log 2>&1
tail -4 gpurun_out/c17_pytest.log | cut -c1-400; cut -c1-200 gpurun_out/c17_sweep_b40.log; cut -c1-200 gpurun_out/c17_sweep_b8.log; cut -c1-200 gpurun_out/c17_sweep_b64.log
