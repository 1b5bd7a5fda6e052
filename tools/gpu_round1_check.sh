#!/bin/bash
# One-call GPU check of the round's final state (GPU minutes are scarce: most valuable steps first, each with its own timeout/log).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout -s KILL 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s1_smoke.log 2>&1; el "smoke rc=$?"; tail -2 gpurun_out/s1_smoke.log
timeout -s KILL 90 python tools/pair_check.py conv > gpurun_out/s2_pair_conv.log 2>&1; PC=$?; el "pair conv rc=$PC"; tail -4 gpurun_out/s2_pair_conv.log
if [ $PC -eq 0 ]; then
  timeout -s KILL 120 python tools/pair_check.py step > gpurun_out/s3_pair_step.log 2>&1; PS=$?; el "pair step rc=$PS"; tail -3 gpurun_out/s3_pair_step.log
else
  PS=1
fi
timeout -s KILL 420 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest.log 2>&1; el "pytest rc=$?"; tail -4 gpurun_out/s4_pytest.log
timeout -s KILL 150 python bench.py --no_cpu_baseline > gpurun_out/s5_bench_default.log 2>&1; el "bench default rc=$?"; tail -1 gpurun_out/s5_bench_default.log | cut -c1-400
if [ $PS -eq 0 ]; then
  DENSEREG_TC_PAIR=1 timeout -s KILL 150 python bench.py --no_cpu_baseline > gpurun_out/s6_bench_pair.log 2>&1; el "bench pair rc=$?"; tail -1 gpurun_out/s6_bench_pair.log | cut -c1-400
  DENSEREG_TC_PAIR=1 timeout -s KILL 300 python -m pytest tests/test_gpu_net.py tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/s7_pytest_pair.log 2>&1; el "pytest pair rc=$?"; tail -3 gpurun_out/s7_pytest_pair.log
fi
el done
