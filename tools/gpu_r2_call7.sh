#!/bin/bash
# Round-2 GPU call 7: warpgroup-aligned roles + setmaxnreg + register-resident running sums (conv_tc.cu, conv_tc_atmem.cu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > gpurun_out/c7_pytest.log 2>&1
timeout -s KILL 400 python tools/r2_sweep.py base split_groups_1 a_tmem_2 a_tmem_0 chunk_train_1 > gpurun_out/c7_sweep.log 2>&1
for cfg in "1 0" "0 0"; do
  set -- $cfg
  DENSEREG_TC_CHUNK_EVAL=$1 DENSEREG_TC_CHUNK_MINKB=$2 timeout -s KILL 200 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c7_bench_infer_c$1.json 2>> gpurun_out/c7_bench_infer.err
  DENSEREG_TC_SPLIT_GROUPS=1 DENSEREG_TC_CHUNK_EVAL=$1 timeout -s KILL 200 python bench.py --config msra_infer --no_cpu_baseline > gpurun_out/c7_bench_infer_c$1_sg1.json 2>> gpurun_out/c7_bench_infer.err
done
timeout -s KILL 200 python tools/bench_infer.py --check 2 > gpurun_out/c7_infer_sweep.json 2> gpurun_out/c7_infer_sweep.err
timeout -s KILL 300 python bench.py > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
tail -5 gpurun_out/c7_pytest.log; cut -c1-300 gpurun_out/c7_sweep.log; for f in gpurun_out/c7_bench_infer_c*.json; do echo $f; cut -c1-170 $f; done; tail -3 gpurun_out/c7_bench_infer.err; cut -c1-900 gpurun_out/c7_infer_sweep.json; cut -c1-300 gpurun_out/c7_bench.json
