#!/bin/bash
# Round-2 GPU call 19: final build -- full tests (+ the experimental-switch suite), smoke, bench of all four configs, ncu of the BRN kernels on big layers.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/c19_pytest.log 2>&1
DENSEREG_TEST_EXPERIMENTAL=1 timeout -s KILL 600 python -m pytest tests/test_gpu_experimental.py -m gpu -q > gpurun_out/c19_pytest_exp.log 2>&1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c19_smoke.log 2>&1
timeout -s KILL 400 python bench.py > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err
timeout -s KILL 300 python bench.py --config msra_infer > gpurun_out/c19_bench_infer.json 2> gpurun_out/c19_bench_infer.err
timeout -s KILL 300 python bench.py --config nyu64_dp --no_cpu_baseline > gpurun_out/c19_bench_nyu.json 2> gpurun_out/c19_bench_nyu.err
timeout -s KILL 300 python bench.py --config vote > gpurun_out/c19_bench_vote.json 2> gpurun_out/c19_bench_vote.err
timeout -s KILL 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/c19_bench_ref.json 2> gpurun_out/c19_bench_ref.err
bash tools/ncu_brn_big.sh > gpurun_out/c19_ncu_brn.log 2>&1
rm -f gpurun_out/qc_ref.pt
tail -3 gpurun_out/c19_pytest.log | cut -c1-300; tail -3 gpurun_out/c19_pytest_exp.log | cut -c1-300; tail -2 gpurun_out/c19_smoke.log; for f in bench bench_infer bench_nyu bench_vote bench_ref; do cut -c1-330 gpurun_out/c19_$f.json; done; cat gpurun_out/r2_kernels_brn.md
