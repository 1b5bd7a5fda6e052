#!/bin/bash
# Round-2 GPU call 9 (8 GPUs): strong scaling of BASELINE configs[2] (NYU J=14, global batch 64) at 1 / 2 / 4 / 8 GPUs, weak scaling point at 8.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout -s KILL 200 python bench.py --config nyu64_dp --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/c9_nyu_n$n.json 2> gpurun_out/c9_nyu_n$n.err
  else
    timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --config nyu64_dp --steps 5 --warmup 3 > gpurun_out/c9_nyu_n$n.json 2> gpurun_out/c9_nyu_n$n.err
  fi
done
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c9_icvl_n8.json 2> gpurun_out/c9_icvl_n8.err
for f in gpurun_out/c9_*.json; do echo $f; grep -v "^NCCL" $f | cut -c1-330; done; tail -n 3 gpurun_out/c9_nyu_n8.err
