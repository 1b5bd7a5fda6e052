#!/bin/bash
# Round-2 GPU call 8: pair kernel on narrower layers.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl
timeout -s KILL 400 python tools/r2_sweep.py base pair_mincout_64 pair_mincout_80 > gpurun_out/c8_sweep.log 2>&1
DENSEREG_TC_PAIR_MINCOUT=64 timeout -s KILL 300 python -m pytest tests/test_gpu_conv.py tests/test_gpu_net.py -m gpu -q -x > gpurun_out/c8_pytest_pair64.log 2>&1
cut -c1-300 gpurun_out/c8_sweep.log; tail -3 gpurun_out/c8_pytest_pair64.log
python - <<'PY'
import json
for l in open("gpurun_out/r2_sweep.jsonl"):
    d=json.loads(l)
    print(d["tag"], d.get("trace_ms"))
PY
