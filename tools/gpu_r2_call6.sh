#!/bin/bash
# Round-2 GPU call 6: tests + short sweep on the current build, then one `ncu --set full` capture per kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/c6_pytest.log 2>&1
timeout -s KILL 600 python tools/r2_sweep.py > gpurun_out/c6_sweep.log 2>&1
timeout -s KILL 1500 bash tools/ncu_kernels.sh > gpurun_out/c6_ncu.log 2>&1
tail -4 gpurun_out/c6_pytest.log; cut -c1-300 gpurun_out/c6_sweep.log; tail -30 gpurun_out/c6_ncu.log
