#!/bin/bash
# Round-2 GPU call 13 (2 GPUs): data-parallel check and scaling points on the final build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout -s KILL 300 $TR tools/dp_check.py > gpurun_out/c13_dp_check.json 2> gpurun_out/c13_dp_check.err
timeout -s KILL 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c13_bench_n2.json 2> gpurun_out/c13_bench_n2.err
timeout -s KILL 300 $TR bench.py --gpus 2 --config nyu64_dp --steps 5 --warmup 3 > gpurun_out/c13_bench_nyu_n2.json 2> gpurun_out/c13_bench_nyu_n2.err
timeout -s KILL 300 python bench.py --config nyu64_dp --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/c13_bench_nyu_n1.json 2> gpurun_out/c13_bench_nyu_n1.err
timeout -s KILL 300 python bench.py --config nyu64_dp --batch_size 8 --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/c13_bench_nyu_b8.json 2> gpurun_out/c13_bench_nyu_b8.err
timeout -s KILL 300 python bench.py --config vote > gpurun_out/c13_bench_vote.json 2> gpurun_out/c13_bench_vote.err
cat gpurun_out/c13_dp_check.json; for f in gpurun_out/c13_bench*.json; do echo $f; grep -v "^NCCL" $f | cut -c1-260; done
