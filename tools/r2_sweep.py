#!/usr/bin/env python
"""Round-2 switch sweep: one child process (tools/quick_check.py) per environment setting -> gpurun_out/r2_sweep.jsonl.
Each child checks the training micro-step against the fp32 FFMA engine (GPU vs GPU) and times it; ~15-25 s per setting.
  gpurun --timeout 1200 -- 'python tools/r2_sweep.py [name ...]'"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
CONFIGS = [
    ("base", {}),
    ("prep_every_forward", {"DENSEREG_PREP_ONCE": "0"}),
    ("tmap_cache_off", {"DENSEREG_TMAP_CACHE": "0"}),
    ("brn_blocks_1184", {"DENSEREG_BRN_BLOCKS": "1184"}),
    ("brn_blocks_592", {"DENSEREG_BRN_BLOCKS": "592"}),
    ("brn_blocks_148", {"DENSEREG_BRN_BLOCKS": "148"}),
    ("pool_bwd_v4", {"DENSEREG_POOL_BWD_V4": "1"}),
    ("stats_per_cta", {"DENSEREG_TC_STATS_PER_CTA": "1"}),
    ("pair_tail", {"DENSEREG_TC_PAIR_TAIL": "1"}),
    ("wgrad_swap", {"DENSEREG_WGRAD_SWAP": "1"}),
    ("wgrad_swap_all", {"DENSEREG_WGRAD_SWAP": "2"}),
    ("wgrad_persist", {"DENSEREG_WGRAD_PERSIST": "1"}),
    ("wgrad_persist_swap", {"DENSEREG_WGRAD_PERSIST": "1", "DENSEREG_WGRAD_SWAP": "1"}),
    ("wgrad_persist_swap_w3", {"DENSEREG_WGRAD_PERSIST": "1", "DENSEREG_WGRAD_SWAP": "1", "DENSEREG_WGRAD_WAVES": "3"}),
    ("wgrad_persist_swap_w4", {"DENSEREG_WGRAD_PERSIST": "1", "DENSEREG_WGRAD_SWAP": "1", "DENSEREG_WGRAD_WAVES": "4"}),
    ("wgrad_w1", {"DENSEREG_WGRAD_WAVES": "1"}),
    ("wgrad_w3", {"DENSEREG_WGRAD_WAVES": "3"}),
    ("wgrad_a_tmem", {"DENSEREG_WGRAD_A_TMEM": "1"}),
    ("wgrad_a_tmem_swap", {"DENSEREG_WGRAD_A_TMEM": "1", "DENSEREG_WGRAD_SWAP": "1"}),
    ("a_tmem_1", {"DENSEREG_TC_A_TMEM": "1"}),
    ("a_tmem_2", {"DENSEREG_TC_A_TMEM": "2"}),
    ("chunk2", {"DENSEREG_TC_CHUNK": "2"}),
    ("chunk4", {"DENSEREG_TC_CHUNK": "4"}),
    ("chunk8", {"DENSEREG_TC_CHUNK": "8"}),
    ("no_side_stream", {"DENSEREG_SIDE_STREAM": "0"}),
    ("no_pair", {"DENSEREG_TC_PAIR": "0"}),
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
    print("%-24s %s" % (name, {k: rec.get(k) for k in ("parity_ok", "ms_per_micro", "crops_per_s", "grad_worst_rel", "um_rel", "xyz_max_mm", "error")}), flush=True)
