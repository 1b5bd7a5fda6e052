#!/bin/bash
# Round-2 GPU call 24: full GPU suite on the pipelined build, smoke, bench line (copy-stream e2e), ncu --set full of the CTA-pair wgrad kernel,
# launch list of two micro-batches of bench.py's timed step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ncu_r2b
timeout -s KILL 500 python -m pytest tests -m gpu -q -x > gpurun_out/c24_pytest.log 2>&1
echo "pytest rc=$?"
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c24_smoke.log 2>&1
echo "smoke rc=$?"
timeout -s KILL 300 python bench.py --no_cpu_baseline --no_other_configs > gpurun_out/c24_bench.json 2> gpurun_out/c24_bench.err
NCU="ncu --set full --clock-control none --import-source on"
timeout -s KILL 240 $NCU -k regex:wgrad_tc_pair_kernel -s 4 -c 2 -f -o gpurun_out/ncu_r2b/wgrad_pair_um_comb_c2 python tools/profile_layer.py --layer s0/um_comb/c2 --what wgrad > gpurun_out/ncu_r2b/wgrad_pair_um_comb_c2.log 2>&1
timeout -s KILL 240 $NCU -k regex:wgrad_tc_pair_kernel -s 4 -c 2 -f -o gpurun_out/ncu_r2b/wgrad_pair_um_full1 python tools/profile_layer.py --layer s0/um_full1 --what wgrad > gpurun_out/ncu_r2b/wgrad_pair_um_full1.log 2>&1
python tools/ncu_summary.py gpurun_out/ncu_r2b > gpurun_out/r2_kernels_pair_wgrad.md 2> gpurun_out/ncu_r2b/summary.err
for f in gpurun_out/ncu_r2b/*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_r2b/$b.raw.csv.gz
done
ncu -i gpurun_out/ncu_r2b/wgrad_pair_um_comb_c2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_r2b/wgrad_pair_um_comb_c2.source.csv.gz
rm -f gpurun_out/ncu_r2b/wgrad_pair_um_full1.ncu-rep
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 1800 --csv --log-file gpurun_out/c24_launches_bench.csv python bench.py --steps 1 --warmup 1 --no_cpu_baseline --no_other_configs > gpurun_out/c24_ncu_bench.log 2>&1
gzip -f gpurun_out/c24_launches_bench.csv
tail -5 gpurun_out/c24_pytest.log | cut -c1-600; tail -2 gpurun_out/c24_smoke.log | cut -c1-300; cut -c1-200 gpurun_out/c24_bench.json; tail -2 gpurun_out/c24_bench.err; cat gpurun_out/r2_kernels_pair_wgrad.md; ls -la gpurun_out/c24_launches_bench.csv.gz
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/c24_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"], d["roofline"]["achieved"], d["roofline"]["dominant_class"], d["roofline"]["other_kernel"]["achieved"])
PY
