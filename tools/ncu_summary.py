#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/ncu/*.ncu-rep, written by tools/ncu_kernels.sh on the GPU box) into one markdown
table: duration, DRAM bytes / GB/s / % of peak, tensor-pipe activity, SM throughput, registers, grid, L2 hit rate.  Runs where ncu is
installed (no GPU needed):   python tools/ncu_summary.py gpurun_out/ncu > profiles/r2_kernels.md"""
import csv, glob, io, os, subprocess, sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}     # -> us / MB
COLS = [("gpu__time_duration.sum", "us", None), ("dram__bytes_read.sum", "MB rd", None), ("dram__bytes_write.sum", "MB wr", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % (active)", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % (elapsed)", 1.0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("launch__grid_size", "grid", 1.0), ("launch__block_size", "block", 1.0)]


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return None


def main(d):
    print("| capture | kernel | " + " | ".join(c[1] for c in COLS) + " | GB/s |")
    print("|---|---|" + "---:|" * (len(COLS) + 1))
    for path in sorted(glob.glob(os.path.join(d, "*.ncu-rep"))):
        r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        rows = [x for x in rows if len(x) > 10]
        if len(rows) < 3:
            print("| %s | (no launches captured) |" % os.path.basename(path)); continue
        hdr = rows[0]
        ki = hdr.index("Kernel Name")
        for row in rows[2:]:
            vals = []
            for name, _, scale in COLS:
                v = num(row[hdr.index(name)]) if name in hdr else None
                if v is not None and scale is None:
                    scale = UNIT.get(rows[1][hdr.index(name)], 1.0)
                vals.append(None if v is None else v * scale)
            us, rd, wr = vals[0], vals[1], vals[2]
            gbs = (rd + wr) * 1e6 / (us * 1e-6) / 1e9 if us and rd is not None and wr is not None else None       # MB and us -> GB/s
            kn = row[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:44]
            print("| %s | %s | " % (os.path.basename(path)[:-8], kn) +
                  " | ".join("-" if v is None else ("%.1f" % v if abs(v) < 1e5 else "%.3g" % v) for v in vals) +
                  " | %s |" % ("-" if gbs is None else "%.0f" % gbs))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ncu")
