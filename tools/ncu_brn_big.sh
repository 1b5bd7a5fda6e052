#!/bin/bash
# ncu --set full captures of the BRN kernels on their BIG layers (the heads of the last stack: 256 / 512 channels at 32x32, batch 40) and of the
# one-cluster kernel for small layers.  In the reverse walk the first BRN layers of a micro-batch are s1/um_comb (512 / 256 channels); in the
# forward walk they are the last ones.  Summaries -> gpurun_out/r2_kernels_brn.md, raw metric tables gzip'ed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ncu_brn
NCU="ncu --set full --clock-control none --import-source on"
cap() { local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout -s KILL 240 $NCU -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/ncu_brn/$name "$@" > gpurun_out/ncu_brn/$name.log 2>&1; }
STEP="python tools/step_once.py --micro 1"
cap big_brn_bwd_reduce   brn_bwd_reduce_v4    0 3 $STEP
cap big_brn_bwd_apply    brn_bwd_apply_v4     0 3 $STEP
cap big_brn_apply        brn_apply_v4         124 6 $STEP
cap small_brn_bwd_cluster brn_bwd_cluster     0 3 $STEP
python tools/ncu_summary.py gpurun_out/ncu_brn > gpurun_out/r2_kernels_brn.md 2> gpurun_out/ncu_brn/summary.err
for f in gpurun_out/ncu_brn/*.ncu-rep; do b=$(basename $f .ncu-rep); ncu -i $f --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/ncu_brn/$b.raw.csv.gz; done
rm -f gpurun_out/ncu_brn/*.ncu-rep
cat gpurun_out/r2_kernels_brn.md
