#!/bin/bash
# Round-2 GPU call 25: where the end-to-end step loses time against the resident one (variants in one process), pipeline on / off.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 200 python tools/e2e_probe.py --pipeline 2 > gpurun_out/c25_probe_p2.json 2> gpurun_out/c25_probe_p2.err
timeout -s KILL 200 python tools/e2e_probe.py --pipeline 1 > gpurun_out/c25_probe_p1.json 2> gpurun_out/c25_probe_p1.err
cat gpurun_out/c25_probe_p2.json gpurun_out/c25_probe_p1.json; tail -2 gpurun_out/c25_probe_p2.err
