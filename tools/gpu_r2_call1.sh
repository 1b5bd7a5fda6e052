#!/bin/bash
# Round-2 GPU call 1: switch sweep (parity vs the fp32 engine + step time per setting), full GPU test suite, per-layer wgrad timings,
# smoke, bench.   gpurun --timeout 2400 -- 'bash tools/gpu_r2_call1.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_gpu.txt 2>&1
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 1500 python tools/r2_sweep.py > gpurun_out/c1_sweep.log 2>&1
timeout -s KILL 120 python tools/time_wgrad.py > gpurun_out/c1_wgrad_base.jsonl 2>&1
DENSEREG_WGRAD_SWAP=1 timeout -s KILL 120 python tools/time_wgrad.py > gpurun_out/c1_wgrad_swap.jsonl 2>&1
DENSEREG_WGRAD_PERSIST=1 DENSEREG_WGRAD_SWAP=1 timeout -s KILL 120 python tools/time_wgrad.py > gpurun_out/c1_wgrad_persist_swap.jsonl 2>&1
DENSEREG_TC_CHUNK=0 timeout -s KILL 1200 python -m pytest tests -m gpu -q > gpurun_out/c1_pytest.log 2>&1
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.log 2>&1
timeout -s KILL 300 python bench.py --no_cpu_baseline > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_pytest.log; tail -2 gpurun_out/c1_smoke.log; cat gpurun_out/c1_sweep.log | cut -c1-220
