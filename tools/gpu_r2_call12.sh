#!/bin/bash
# Round-2 GPU call 12: gradient aliasing (no copies of d(residual sum)), full tests, bench, launch list of bench.py itself.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/c12_pytest.log 2>&1
timeout -s KILL 400 python tools/r2_sweep.py base no_grad_alias no_pdl > gpurun_out/c12_sweep.log 2>&1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12_smoke.log 2>&1
timeout -s KILL 400 python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
timeout -s KILL 300 python bench.py --config msra_infer > gpurun_out/c12_bench_infer.json 2> gpurun_out/c12_bench_infer.err
DENSEREG_TC_CHUNK_EVAL=0 timeout -s KILL 300 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c12_bench_infer_fast.json 2>> gpurun_out/c12_bench_infer.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 5300 -c 1100 --csv --log-file gpurun_out/c12_launches_bench.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline > gpurun_out/c12_ncu_bench.log 2>&1
gzip -f gpurun_out/c12_launches_bench.csv
tail -4 gpurun_out/c12_pytest.log; cut -c1-260 gpurun_out/c12_sweep.log; tail -2 gpurun_out/c12_smoke.log; cut -c1-500 gpurun_out/c12_bench.json; cut -c1-300 gpurun_out/c12_bench_infer.json; cut -c1-200 gpurun_out/c12_bench_infer_fast.json; tail -2 gpurun_out/c12_ncu_bench.log
