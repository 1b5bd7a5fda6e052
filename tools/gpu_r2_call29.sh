#!/bin/bash
# Round-2 GPU call 29: the conv + network parity suites with the micro-batch pipeline FORCED on for every training engine, and with the wgrad
# splitters in their round-to-nearest form; pipeline tests; smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
DENSEREG_TEST_EXPERIMENTAL=1 timeout -s KILL 420 python -m pytest tests/test_gpu_experimental.py -m gpu -q -k "SPLIT_TRUNC or PIPELINE" > gpurun_out/c29_pytest_switches.log 2>&1
echo "switches rc=$?"
timeout -s KILL 200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_net.py -m gpu -q -x > gpurun_out/c29_pytest.log 2>&1
echo "pytest rc=$?"
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c29_smoke.log 2>&1
echo "smoke rc=$?"
tail -25 gpurun_out/c29_pytest_switches.log | cut -c1-1200; tail -3 gpurun_out/c29_pytest.log | cut -c1-400; tail -2 gpurun_out/c29_smoke.log | cut -c1-300
