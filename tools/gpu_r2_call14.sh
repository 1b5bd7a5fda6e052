#!/bin/bash
# Round-2 GPU call 14: one-cluster BRN backward for small layers (A/B at batch 40 and batch 8), full tests, memcheck of smoke().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/c14_pytest.log 2>&1
timeout -s KILL 500 python tools/r2_sweep.py base brn_small_0 brn_small_48k brn_small_192k brn_small_384k > gpurun_out/c14_sweep_b40.log 2>&1
SWEEP_ARGS="--batch 8 --J 14" timeout -s KILL 400 python tools/r2_sweep.py base brn_small_0 brn_small_48k brn_small_192k brn_small_384k > gpurun_out/c14_sweep_b8.log 2>&1
timeout -s KILL 900 compute-sanitizer --tool memcheck --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c14_memcheck.log 2>&1
tail -4 gpurun_out/c14_pytest.log; cut -c1-230 gpurun_out/c14_sweep_b40.log; cut -c1-230 gpurun_out/c14_sweep_b8.log; tail -12 gpurun_out/c14_memcheck.log | cut -c1-300
