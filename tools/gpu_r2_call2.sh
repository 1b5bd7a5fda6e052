#!/bin/bash
# Round-2 GPU call 2: second sweep pass on the new defaults, Adam fix, e2e accuracy under the two-level accumulation, new bench.py configs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 1200 python tools/r2_sweep.py > gpurun_out/c2_sweep.log 2>&1
for c in 0 1 2 4; do
  DENSEREG_TC_CHUNK=$c timeout -s KILL 300 python -m pytest tests/test_gpu_net.py -m gpu -q -k "infer_end_to_end and tf32x3" > gpurun_out/c2_e2e_chunk$c.log 2>&1
  for f in infer_e2e_S1F64J16_tf32x3 infer_e2e_S2F128J14_tf32x3; do cp gpurun_out/$f.json gpurun_out/c2_${f}_chunk$c.json; done
done
DENSEREG_TC_CHUNK=0 timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.log 2>&1
timeout -s KILL 300 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
timeout -s KILL 300 python bench.py --config nyu64_dp --no_cpu_baseline > gpurun_out/c2_bench_nyu.json 2> gpurun_out/c2_bench_nyu.err
timeout -s KILL 300 python bench.py --config msra_infer > gpurun_out/c2_bench_infer.json 2> gpurun_out/c2_bench_infer.err
timeout -s KILL 300 python bench.py --config vote > gpurun_out/c2_bench_vote.json 2> gpurun_out/c2_bench_vote.err
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/c2_bench_ref.json 2> gpurun_out/c2_bench_ref.err
tail -3 gpurun_out/c2_pytest.log; cut -c1-400 gpurun_out/c2_sweep.log; tail -2 gpurun_out/c2_bench*.err; for c in 0 1 2 4; do tail -1 gpurun_out/c2_e2e_chunk$c.log; done
