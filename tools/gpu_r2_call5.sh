#!/bin/bash
# Round-2 GPU call 5 (2 GPUs): in-library NCCL data-parallel check, weak scaling (configs[1]) and strong scaling (configs[2]) at N=2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout -s KILL 300 $TR tools/dp_check.py > gpurun_out/c5_dp_check.json 2> gpurun_out/c5_dp_check.err
timeout -s KILL 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c5_bench_n2.json 2> gpurun_out/c5_bench_n2.err
timeout -s KILL 300 $TR bench.py --gpus 2 --config nyu64_dp --steps 5 --warmup 3 > gpurun_out/c5_bench_nyu_n2.json 2> gpurun_out/c5_bench_nyu_n2.err
timeout -s KILL 300 python bench.py --config nyu64_dp --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/c5_bench_nyu_n1.json 2> gpurun_out/c5_bench_nyu_n1.err
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/c5_bench_n1.json 2> gpurun_out/c5_bench_n1.err
timeout -s KILL 300 $TR bench.py --gpus 2 --config msra_infer --no_cpu_baseline > gpurun_out/c5_bench_infer_n2.json 2> gpurun_out/c5_bench_infer_n2.err
cat gpurun_out/c5_dp_check.json; tail -3 gpurun_out/c5_dp_check.err; for f in gpurun_out/c5_bench*.json; do echo $f; cut -c1-400 $f; done; tail -2 gpurun_out/c5_bench*.err
