#!/bin/bash
# Round-2 GPU call 26: CUDA_DEVICE_MAX_CONNECTIONS (hardware work queues; default 8) against the ~15 streams of the pipelined engine.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 16 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$n timeout -s KILL 200 python tools/e2e_probe.py --pipeline 2 > gpurun_out/c26_probe_p2_conn$n.json 2> gpurun_out/c26_probe_conn$n.err
  echo "conn $n: $(cat gpurun_out/c26_probe_p2_conn$n.json)"
done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout -s KILL 200 python tools/e2e_probe.py --pipeline 2 --batch 8 > gpurun_out/c26_probe_p2_conn32_b8.json 2>/dev/null
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout -s KILL 200 python tools/e2e_probe.py --pipeline 2 --batch 8 > gpurun_out/c26_probe_p2_conn8_b8.json 2>/dev/null
echo "b8 conn32: $(cat gpurun_out/c26_probe_p2_conn32_b8.json)"; echo "b8 conn8: $(cat gpurun_out/c26_probe_p2_conn8_b8.json)"
