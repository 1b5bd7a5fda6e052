#!/bin/bash
# Round-2 GPU call 23: CTA-pair wgrad kernel (wide layers): conv parity tests (every layer's dW vs F.conv2d), A/B inside the pipelined step, bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl
timeout -s KILL 240 python -m pytest tests/test_gpu_conv.py -m gpu -q -x -k "tensor_core_path or forward_and_backward" > gpurun_out/c23_pytest_conv.log 2>&1
echo "pytest rc=$?"
SWEEP_TIMEOUT=100 SWEEP_ARGS="--micro 5" timeout -s KILL 400 python tools/r2_sweep.py pipe2 pipe2_nopairw pipe2_pairw_129 > gpurun_out/c23_sweep_b40.log 2>&1
SWEEP_TIMEOUT=100 SWEEP_ARGS="--batch 8 --J 14 --micro 5" timeout -s KILL 200 python tools/r2_sweep.py pipe2 pipe2_nopairw > gpurun_out/c23_sweep_b8.log 2>&1
timeout -s KILL 300 python bench.py --no_cpu_baseline --no_other_configs > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
tail -6 gpurun_out/c23_pytest_conv.log | cut -c1-900; cut -c1-250 gpurun_out/c23_sweep_b40.log; cut -c1-250 gpurun_out/c23_sweep_b8.log; cut -c1-300 gpurun_out/c23_bench.json; tail -3 gpurun_out/c23_bench.err
