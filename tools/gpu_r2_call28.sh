#!/bin/bash
# Round-2 GPU call 28: the driver's own command -- python bench.py with no flags (headline + cpu_baseline + other_configs), timed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout -s KILL 600 python bench.py > gpurun_out/c28_bench.json 2> gpurun_out/c28_bench.err ) 2> gpurun_out/c28_bench_time.txt
cat gpurun_out/c28_bench_time.txt; tail -3 gpurun_out/c28_bench.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c28_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "mean_joint_err_mm", "gpu_launches")}, d["e2e"], d["cpu_baseline"], d["clocks"])
r = d["roofline"]; print(r["achieved"], r["frac"], r["dominant_class"], r["traffic"], r["whole_step"], r["other_kernel"]["achieved"])
print(json.dumps(d.get("other_configs"))[:2500])
PY
