#!/bin/bash
# Round-2 GPU call 27 (2 GPUs): data-parallel correctness of the pipelined engine (in-library all-reduce, plain and overlapped, vs torch.distributed),
# bench.py on 2 GPUs for both training configs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 240 $TR --master-port 29511 tools/dp_check.py > gpurun_out/c27_dp_check.json 2> gpurun_out/c27_dp_check.err
timeout -s KILL 300 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c27_bench_n2.json 2> gpurun_out/c27_bench_n2.err
timeout -s KILL 300 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --config nyu64_dp > gpurun_out/c27_bench_nyu_n2.json 2> gpurun_out/c27_bench_nyu_n2.err
grep '^{' gpurun_out/c27_dp_check.json; tail -3 gpurun_out/c27_dp_check.err | cut -c1-300
for f in c27_bench_n2 c27_bench_nyu_n2; do grep '^{' gpurun_out/$f.json | cut -c1-260; tail -2 gpurun_out/$f.err | cut -c1-300; done
python - <<'PY'
import json
for f in ("c27_bench_n2", "c27_bench_nyu_n2"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"], d["config"]["collective"][:80])
    except Exception as e:
        print(f, "ERR", e)
PY
