#!/bin/bash
# Round-2 GPU call 3: full test suite on the new defaults, inference bench with the two-level accumulation, ncu launch list of the step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/c3_pytest.log 2>&1
timeout -s KILL 300 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c3_bench_infer_chunk1.json 2> gpurun_out/c3_bench_infer.err
DENSEREG_TC_CHUNK_EVAL=0 timeout -s KILL 300 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c3_bench_infer_chunk0.json 2>> gpurun_out/c3_bench_infer.err
DENSEREG_TC_CHUNK_EVAL=4 timeout -s KILL 300 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c3_bench_infer_chunk4.json 2>> gpurun_out/c3_bench_infer.err
timeout -s KILL 300 python bench.py --no_cpu_baseline > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_launches.csv python tools/step_once.py --micro 2 > gpurun_out/c3_ncu.log 2>&1
tail -4 gpurun_out/c3_pytest.log; cat gpurun_out/c3_bench_infer_chunk*.json | cut -c1-300; cut -c1-300 gpurun_out/c3_bench.json
