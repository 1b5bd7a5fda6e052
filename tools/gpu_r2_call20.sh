#!/bin/bash
# Round-2 GPU call 20: BRN backward reduce with 8 load pairs in flight (A/B), default bench.py run incl. its other_configs section, training tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep.jsonl gpurun_out/qc_ref.pt
timeout -s KILL 400 python -m pytest tests/test_gpu_net.py -m gpu -q > gpurun_out/c20_pytest.log 2>&1
timeout -s KILL 400 python tools/r2_sweep.py base brn_reduce_unroll_4 > gpurun_out/c20_sweep_b40.log 2>&1
SWEEP_ARGS="--batch 8 --J 14" timeout -s KILL 300 python tools/r2_sweep.py base brn_reduce_unroll_4 > gpurun_out/c20_sweep_b8.log 2>&1
rm -f gpurun_out/qc_ref.pt
( time timeout -s KILL 600 python bench.py > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err ) 2> gpurun_out/c20_bench_time.txt
tail -3 gpurun_out/c20_pytest.log | cut -c1-300; cut -c1-200 gpurun_out/c20_sweep_b40.log; cut -c1-200 gpurun_out/c20_sweep_b8.log; cat gpurun_out/c20_bench_time.txt; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c20_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "mean_joint_err_mm")}, d["e2e"], d["cpu_baseline"])
print(json.dumps(d.get("other_configs"))[:1500])
PY
