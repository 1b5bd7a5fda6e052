#!/bin/bash
# Round-2 GPU call 22: wgrad splitters without the hi write (the tensor core truncates the landed fp32 tile itself): conv parity tests with the
# switch on, A/B with the micro-batch pipeline; bench.py line with the pipeline (no CPU leg, no other configs).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl
DENSEREG_SPLIT_TRUNC=1 timeout -s KILL 300 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "tensor_core_path or forward_and_backward" > gpurun_out/c22_pytest_trunc.log 2>&1
SWEEP_ARGS="--micro 5" timeout -s KILL 420 python tools/r2_sweep.py pipe2 pipe2_trunc trunc pipe2_trunc_waves2 pipe2_waves2 > gpurun_out/c22_sweep_b40.log 2>&1
SWEEP_ARGS="--batch 8 --J 14 --micro 5" timeout -s KILL 200 python tools/r2_sweep.py pipe2 pipe2_trunc > gpurun_out/c22_sweep_b8.log 2>&1
timeout -s KILL 300 python bench.py --no_cpu_baseline --no_other_configs > gpurun_out/c22_bench.json 2> gpurun_out/c22_bench.err
tail -6 gpurun_out/c22_pytest_trunc.log | cut -c1-600; cut -c1-250 gpurun_out/c22_sweep_b40.log; cut -c1-250 gpurun_out/c22_sweep_b8.log; cut -c1-700 gpurun_out/c22_bench.json; tail -3 gpurun_out/c22_bench.err
