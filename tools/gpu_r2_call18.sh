#!/bin/bash
# Round-2 GPU call 18 (8 GPUs): scaling points of the final build -- BASELINE configs[2] (NYU J=14, GLOBAL batch 64: strong scaling) at 4 and 8 GPUs,
# configs[1] (ICVL, 40 crops per GPU: weak scaling) at 8 GPUs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519"
timeout -s KILL 300 $TR8 bench.py --gpus 8 --config nyu64_dp --steps 8 --warmup 3 > gpurun_out/c18_bench_nyu_n8.json 2> gpurun_out/c18_bench_nyu_n8.err
timeout -s KILL 300 $TR8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/c18_bench_n8.json 2> gpurun_out/c18_bench_n8.err
timeout -s KILL 300 $TR4 bench.py --gpus 4 --config nyu64_dp --steps 8 --warmup 3 > gpurun_out/c18_bench_nyu_n4.json 2> gpurun_out/c18_bench_nyu_n4.err
for f in gpurun_out/c18_bench*.json; do echo $f; grep -v "^NCCL" $f | cut -c1-260; done
