#!/usr/bin/env python
"""Round-2 switch sweep: one child process (tools/quick_check.py) per environment setting -> gpurun_out/r2_sweep.jsonl.
Each child checks the training micro-step against the fp32 FFMA engine (GPU vs GPU) and times it; ~15-25 s per setting.
  gpurun --timeout 1200 -- 'python tools/r2_sweep.py [name ...]'"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
CONFIGS = [   # round-2 fifth pass: setmaxnreg role layout, register-resident running sums, two splitter warpgroups in the A-TMEM conv kernel
    ("base", {}),
    ("split_groups_1", {"DENSEREG_TC_SPLIT_GROUPS": "1"}),
    ("a_tmem_2", {"DENSEREG_TC_A_TMEM": "2"}),
    ("a_tmem_0", {"DENSEREG_TC_A_TMEM": "0"}),
    ("pair_mincout_64", {"DENSEREG_TC_PAIR_MINCOUT": "64"}),
    ("pair_mincout_80", {"DENSEREG_TC_PAIR_MINCOUT": "80"}),
    ("pair_mincout_64_w", {"DENSEREG_TC_PAIR_MINCOUT": "64", "DENSEREG_TC_PAIR_MINWORK": "0"}),
    ("wgrad_streams_1", {"DENSEREG_WGRAD_STREAMS": "1"}),
    ("no_lanes", {"DENSEREG_LANES": "0"}),
]
want = set(sys.argv[1:])
path = os.path.join(OUT, "r2_sweep.jsonl")
for name, env in CONFIGS:
    if want and name not in want:
        continue
    e = dict(os.environ, **env)
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_check.py"), "--tag", name], env=e, capture_output=True, text=True,
                           timeout=int(os.environ.get("SWEEP_TIMEOUT", "150")))
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        rec = json.loads(line[-1]) if line else {"tag": name, "error": "no output", "stderr": r.stderr[-800:], "rc": r.returncode}
    except subprocess.TimeoutExpired:
        rec = {"tag": name, "error": "timeout (hang?)"}
    with open(path, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print("%-24s %s" % (name, {k: rec.get(k) for k in ("parity_ok", "ms_per_micro", "crops_per_s", "cpu_enqueue_one_micro_ms", "grad_worst_rel", "xyz_max_mm", "trace_ms", "error")}), flush=True)
