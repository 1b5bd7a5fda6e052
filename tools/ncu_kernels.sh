#!/bin/bash
# One `ncu --set full` capture per kernel of the training step and of inference (profiles/r2_kernels.md is built from these by
# tools/ncu_summary.py).  Each capture re-runs a short program and profiles 2 launches of ONE kernel name from the second micro-batch on.
#   gpurun --timeout 1800 -- 'bash tools/ncu_kernels.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/ncu
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # name regex skip count program...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout -s KILL 240 $NCU -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/ncu/$name "$@" > gpurun_out/ncu/$name.log 2>&1
}
STEP="python tools/step_once.py --micro 2"
# big-layer conv / wgrad through the per-layer debug entry points (isolated, B=40)
cap conv_pair_um_comb_c2   conv_tc_pair_kernel   4 2 python tools/profile_layer.py --layer s0/um_comb/c2 --what fwd
cap conv_atmem_um_res1_c2  conv_tc_atmem_kernel  4 2 python tools/profile_layer.py --layer s0/um_res1/c2 --what fwd
cap conv_chunk_um_comb_c2  "conv_tc_kernel"      4 2 python tools/profile_layer.py --layer s0/um_comb/c2 --what fwd_chunk
cap wgrad_um_comb_c2       wgrad_tc_kernel       4 2 python tools/profile_layer.py --layer s0/um_comb/c2 --what wgrad
cap wgrad_um_res2_c3       wgrad_tc_kernel       4 2 python tools/profile_layer.py --layer s0/um_res2/c3 --what wgrad
# kernels of the training step, inside the step (second micro-batch)
cap step_conv_pair         conv_tc_pair_kernel   130 3 $STEP
cap step_conv_atmem        conv_tc_atmem_kernel  280 3 $STEP
cap step_wgrad             wgrad_tc_kernel       150 3 $STEP
cap step_wgrad_simt        conv_wgrad_kernel     40 2 $STEP
cap step_brn_bwd_reduce    brn_bwd_reduce_v4     160 3 $STEP
cap step_brn_bwd_apply     brn_bwd_apply_v4      160 3 $STEP
cap step_brn_apply         brn_apply_v4          160 3 $STEP
cap step_maxpool_bwd       maxpool_bwd_v4        10 2 $STEP
cap step_maxpool           maxpool_kernel        10 2 $STEP
cap step_upadd             upadd_kernel          9 2 $STEP
cap step_copy_view         copy_view_v4          60 3 $STEP
cap step_bias_bwd          bias_bwd_kernel       13 3 $STEP
cap step_loss              loss_kernel           1 1 $STEP
cap step_wd                wd_kernel             1 1 $STEP
cap step_adam              adam_kernel           0 1 $STEP
cap step_prep_weights      prep_weights_kernel   0 1 $STEP
cap step_stem_conv         conv_fwd_kernel       1 1 $STEP
cap infer_vote             vote_kernel           1 1 python tools/step_once.py --infer --micro 2 --batch 64
cap vote_microbench        vote_kernel           2 1 python tools/bench_vote.py --batch 1024 --iters 2 --cpu_samples 0
# the reports are ~4 MB each and gpurun brings back at most 64 MiB: summarise them here, keep the raw metric tables (gzip) and three reports
python tools/ncu_summary.py gpurun_out/ncu > gpurun_out/r2_kernels.md 2> gpurun_out/ncu/summary.err
for f in gpurun_out/ncu/*.ncu-rep; do
  b=$(basename $f .ncu-rep)
  ncu -i $f --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/ncu/$b.raw.csv.gz
done
for b in wgrad_um_comb_c2 conv_atmem_um_res1_c2 step_wgrad; do
  ncu -i gpurun_out/ncu/$b.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/ncu/$b.source.csv.gz
done
mkdir -p gpurun_out/ncu_keep
mv gpurun_out/ncu/conv_pair_um_comb_c2.ncu-rep gpurun_out/ncu/conv_atmem_um_res1_c2.ncu-rep gpurun_out/ncu/wgrad_um_comb_c2.ncu-rep gpurun_out/ncu_keep/ 2>/dev/null
rm -f gpurun_out/ncu/*.ncu-rep
du -sh gpurun_out/ncu gpurun_out/ncu_keep; cat gpurun_out/r2_kernels.md
