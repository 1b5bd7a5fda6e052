#!/bin/bash
# Round-2 GPU call 10: programmatic dependent launch on / off at batch 40 and batch 8, tests, inference.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/c10_pytest.log 2>&1
timeout -s KILL 400 python tools/r2_sweep.py base no_pdl no_lanes no_lanes_no_pdl > gpurun_out/c10_sweep.log 2>&1
SWEEP_ARGS="--batch 8 --J 14" timeout -s KILL 300 python tools/r2_sweep.py base no_pdl > gpurun_out/c10_sweep_b8.log 2>&1
timeout -s KILL 200 python tools/bench_infer.py --check 2 > gpurun_out/c10_infer_sweep.json 2> gpurun_out/c10_infer_sweep.err
DENSEREG_PDL=0 timeout -s KILL 200 python tools/bench_infer.py --check 0 > gpurun_out/c10_infer_sweep_nopdl.json 2>> gpurun_out/c10_infer_sweep.err
timeout -s KILL 300 python bench.py --no_cpu_baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
tail -5 gpurun_out/c10_pytest.log; cut -c1-260 gpurun_out/c10_sweep.log; cut -c1-260 gpurun_out/c10_sweep_b8.log; cut -c1-700 gpurun_out/c10_infer_sweep.json; cut -c1-500 gpurun_out/c10_infer_sweep_nopdl.json; cut -c1-300 gpurun_out/c10_bench.json; tail -3 gpurun_out/c10_bench.err
