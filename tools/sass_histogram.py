#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass): which kernels contain the Blackwell tensor-core / TMA / TMEM
instructions.   python tools/sass_histogram.py > profiles/sass_r2.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "densereg_b200", "libdensereg_sm100.so")
KEY = ["USETMAXREG", "UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTCCP", "SYNCS", "HMMA", "FFMA",
       "LDG", "STG", "RED", "ATOM", "LDS", "STS", "SHFL", "BAR", "MEMBAR", "CCTL", "ERRBAR", "LDL", "STL"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and kern:
        op = m.group(1)
        hist[kern][op.split(".")[0]] += 1
        if op.startswith(("UTC", "UTMA", "LDTM", "STTM", "USETMAXREG")):
            hist[kern]["=" + op] += 1
dem = subprocess.run(["c++filt"], input="\n".join(hist.keys()), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode histogram per kernel of densereg_b200/libdensereg_sm100.so (sm_100a), `cuobjdump -sass`; columns = instruction counts.")
print("# tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS.")
tot = collections.Counter()
for (k, h), d in zip(hist.items(), dem):
    name = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "").replace("void ", ""))
    n = sum(v for kk, v in h.items() if not kk.startswith("="))
    cols = " ".join("%s=%d" % (kk, h[kk]) for kk in KEY if h.get(kk))
    det = " ".join("%s:%d" % (kk[1:], v) for kk, v in sorted(h.items()) if kk.startswith("="))
    print("%-44s insts=%-6d %s" % (name[:44], n, cols))
    if det:
        print("%-44s   variants: %s" % ("", det))
    for kk in KEY:
        tot[kk] += h.get(kk, 0)
print("TOTAL " + " ".join("%s=%d" % (kk, tot[kk]) for kk in KEY if tot[kk]))
