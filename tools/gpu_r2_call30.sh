#!/bin/bash
# Round-2 GPU call 30: conv + network parity suites with DENSEREG_SPLIT_TRUNC=0 and with DENSEREG_PIPELINE=2 forced on every training engine.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DENSEREG_TEST_EXPERIMENTAL=1 timeout -s KILL 420 python -m pytest tests/test_gpu_experimental.py -m gpu -q -k "env10 or env11" > gpurun_out/c30_pytest_switches.log 2>&1
echo "switches rc=$?"
tail -30 gpurun_out/c30_pytest_switches.log | cut -c1-1500
