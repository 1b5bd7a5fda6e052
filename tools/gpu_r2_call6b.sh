#!/bin/bash
# Round-2 GPU call 6b: one `ncu --set full` capture per kernel, summarised on the box (tools/ncu_kernels.sh).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 1800 bash tools/ncu_kernels.sh > gpurun_out/c6_ncu.log 2>&1
timeout -s KILL 300 python bench.py --no_cpu_baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
tail -60 gpurun_out/c6_ncu.log | cut -c1-330; cut -c1-300 gpurun_out/c6_bench.json
