#!/usr/bin/env python
"""Round-2 switch sweep: one child process (tools/quick_check.py) per environment setting -> gpurun_out/r2_sweep.jsonl.
Each child checks the training micro-step against the fp32 FFMA engine (GPU vs GPU) and times it; ~15-25 s per setting.
  gpurun --timeout 1200 -- 'python tools/r2_sweep.py [name ...]'"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out"); os.makedirs(OUT, exist_ok=True)
CONFIGS = [   # round-2 sixth pass: programmatic dependent launch
    ("base", {}),
    ("pipe2", {"DENSEREG_PIPELINE": "2"}),                         # micro-batch pipeline: forward(i+1) next to backward(i)
    ("pipe2_prio", {"DENSEREG_PIPELINE": "2", "DENSEREG_CHAIN_PRIO": "1"}),
    ("pipe2_wgrad_a_tmem", {"DENSEREG_PIPELINE": "2", "DENSEREG_WGRAD_A_TMEM": "1"}),
    ("pipe2_no_lanes", {"DENSEREG_PIPELINE": "2", "DENSEREG_LANES": "0"}),
    ("pipe2_wgrad_streams_1", {"DENSEREG_PIPELINE": "2", "DENSEREG_WGRAD_STREAMS": "1"}),
    ("chain_prio", {"DENSEREG_CHAIN_PRIO": "1"}),
    ("pipe2_trunc", {"DENSEREG_PIPELINE": "2", "DENSEREG_SPLIT_TRUNC": "1"}),       # wgrad: landed fp32 tile = hi operand, splitters write lo only
    ("trunc", {"DENSEREG_SPLIT_TRUNC": "1"}),
    ("pipe2_trunc_waves2", {"DENSEREG_PIPELINE": "2", "DENSEREG_SPLIT_TRUNC": "1", "DENSEREG_WGRAD_WAVES": "2"}),
    ("pipe2_waves2", {"DENSEREG_PIPELINE": "2", "DENSEREG_WGRAD_WAVES": "2"}),
    ("pipe2_nopairw", {"DENSEREG_PIPELINE": "2", "DENSEREG_WGRAD_PAIR_MINM": "0"}),      # CTA-pair wgrad kernel off / also for 129..255-row layers
    ("pipe2_pairw_129", {"DENSEREG_PIPELINE": "2", "DENSEREG_WGRAD_PAIR_MINM": "129"}),
    ("no_pdl", {"DENSEREG_PDL": "0"}),
    ("no_grad_alias", {"DENSEREG_GRAD_ALIAS": "0"}),
    ("no_lanes", {"DENSEREG_LANES": "0"}),
    ("no_lanes_no_pdl", {"DENSEREG_LANES": "0", "DENSEREG_PDL": "0"}),
    ("a_tmem_2", {"DENSEREG_TC_A_TMEM": "2"}),
    ("wgrad_streams_1", {"DENSEREG_WGRAD_STREAMS": "1"}),
    ("ew_reverse_0", {"DENSEREG_EW_REVERSE": "0"}),              # BRN normalise / backward reduce walk front to back like the convs
    ("brn_small_0", {"DENSEREG_BRN_SMALL_ELEMS": "0"}),            # one-cluster BRN backward off / larger reach (default 96 k elements)
    ("brn_small_48k", {"DENSEREG_BRN_SMALL_ELEMS": "49152"}),
    ("brn_small_192k", {"DENSEREG_BRN_SMALL_ELEMS": "196608"}),
    ("brn_small_384k", {"DENSEREG_BRN_SMALL_ELEMS": "393216"}),
]
EXTRA = [a for a in os.environ.get("SWEEP_ARGS", "").split() if a]     # e.g. SWEEP_ARGS="--batch 8 --J 14"
want = set(sys.argv[1:])
path = os.path.join(OUT, "r2_sweep.jsonl")
for name, env in CONFIGS:
    if want and name not in want:
        continue
    e = dict(os.environ, **env)
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_check.py"), "--tag", name] + EXTRA, env=e, capture_output=True, text=True,
                           timeout=int(os.environ.get("SWEEP_TIMEOUT", "150")))
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        rec = json.loads(line[-1]) if line else {"tag": name, "error": "no output", "stderr": r.stderr[-800:], "rc": r.returncode}
    except subprocess.TimeoutExpired:
        rec = {"tag": name, "error": "timeout (hang?)"}
    with open(path, "a") as f:
        f.write(json.dumps(rec) + "\n")
    print("%-24s %s" % (name, {k: rec.get(k) for k in ("parity_ok", "ms_per_micro", "crops_per_s", "cpu_enqueue_one_micro_ms", "grad_worst_rel", "xyz_max_mm", "trace_ms", "error")}), flush=True)
