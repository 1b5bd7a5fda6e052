#!/bin/bash
# Round-2 GPU call 31: smoke() with its pipelined second micro-batch.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c31_smoke.log 2>&1
echo "smoke rc=$?"; tail -4 gpurun_out/c31_smoke.log | cut -c1-500
