#!/bin/bash
# Round-2 GPU call 4: lanes (multi-stream op scheduling) on/off, full tests with lanes, inference accuracy/throughput vs chunk policy.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 900 python tools/r2_sweep.py > gpurun_out/c4_sweep.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q > gpurun_out/c4_pytest.log 2>&1
for cfg in "0 0" "1 0" "1 16" "2 16" "4 16" "4 8" "2 8"; do
  set -- $cfg
  tag=c$1_m$2
  DENSEREG_TC_CHUNK_EVAL=$1 DENSEREG_TC_CHUNK_MINKB=$2 timeout -s KILL 200 python -m pytest tests/test_gpu_net.py -m gpu -q -k "infer_end_to_end and tf32x3" > gpurun_out/c4_e2e_$tag.log 2>&1
  for f in infer_e2e_S1F64J16_tf32x3 infer_e2e_S2F128J14_tf32x3; do cp gpurun_out/$f.json gpurun_out/c4_${f}_$tag.json; done
  DENSEREG_TC_CHUNK_EVAL=$1 DENSEREG_TC_CHUNK_MINKB=$2 timeout -s KILL 200 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c4_bench_infer_$tag.json 2>> gpurun_out/c4_bench_infer.err
done
DENSEREG_LANES=0 timeout -s KILL 200 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c4_bench_infer_nolanes.json 2>> gpurun_out/c4_bench_infer.err
timeout -s KILL 200 python tools/bench_infer.py --check 0 > gpurun_out/c4_infer_sweep.json 2> gpurun_out/c4_infer_sweep.err
DENSEREG_LANES=0 timeout -s KILL 200 python tools/bench_infer.py --check 0 > gpurun_out/c4_infer_sweep_nolanes.json 2>> gpurun_out/c4_infer_sweep.err
timeout -s KILL 300 python bench.py --no_cpu_baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
tail -4 gpurun_out/c4_pytest.log; cut -c1-330 gpurun_out/c4_sweep.log; for f in gpurun_out/c4_bench_infer_*.json; do echo $f; cut -c1-170 $f; done; tail -3 gpurun_out/c4_bench_infer.err; cut -c1-600 gpurun_out/c4_infer_sweep.json; cut -c1-600 gpurun_out/c4_infer_sweep_nolanes.json
