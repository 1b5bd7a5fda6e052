#!/usr/bin/env python
"""Per-layer time table of one training micro-batch: runs tools/step_once.py with DENSEREG_TRACE=1 (each conv / dgrad / wgrad launch
timed alone) and aggregates the TRACE lines by (kind, shape).  Usage: python tools/layer_times.py [--pair] > gpurun_out/layer_times.txt"""
import collections, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
env = dict(os.environ, DENSEREG_TRACE="1", DENSEREG_SIDE_STREAM="0")
if "--single" in sys.argv:
    env["DENSEREG_TC_PAIR"] = "0"
r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_once.py"), "--micro", "2"], env=env, capture_output=True, text=True)
rows = [l.split() for l in r.stderr.splitlines() if l.startswith("TRACE")]
half = len(rows) // 2
rows = rows[half:]                       # second micro-batch (warm)
agg = collections.OrderedDict()
for _, kind, B, H, cin, cout, k, kern, ms in rows:
    key = (kind, int(H), int(cin), int(cout), int(k), kern)
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += float(ms)
B = int(rows[0][2]) if rows else 0
tot = sum(v[1] for v in agg.values())
print("batch %d: %d conv-type launches, %.3f ms serialised" % (B, len(rows), tot))
print("%-6s %4s %5s %5s %2s %-5s %4s %9s %8s %8s %6s" % ("kind", "HW", "Cin", "Cout", "k", "kern", "n", "total ms", "ms each", "TFLOP/s", "share"))
for (kind, H, cin, cout, k, kern), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    fl = 2.0 * B * H * H * k * k * cin * cout
    print("%-6s %4d %5d %5d %2d %-5s %4d %9.3f %8.4f %8.1f %5.1f%%" % (kind, H, cin, cout, k, kern, n, ms, ms / n, fl / (ms / n) / 1e9, 100 * ms / tot))
by_kind = collections.defaultdict(float)
for (kind, *_), (n, ms) in agg.items():
    by_kind[kind] += ms
print({k: round(v, 3) for k, v in by_kind.items()})
if r.returncode != 0:
    print(r.stderr[-2000:])
