#!/bin/bash
# Round-2 opener: verify and measure every opt-in switch written (unverified) at the end of round 1, one child process per setting
# (the switches are read once per process).  For each: the conv + network parity tests, then the training bench.
#   gpurun --timeout 1500 -- 'bash tools/r2_sweep.sh'      -> gpurun_out/r2_sweep.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_sweep.txt
: > $OUT
run() {
  local tag="$1"; shift
  echo "=== $tag: $*" | tee -a $OUT
  env "$@" timeout -s KILL 240 python -m pytest tests/test_gpu_conv.py tests/test_gpu_net.py -m gpu -x -q 2>&1 | tail -2 | tee -a $OUT
  env "$@" timeout -s KILL 120 python bench.py --no_cpu_baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('crops/s %.1f  ms/step %.2f  roofline kernel %.4f ms' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))" 2>&1 | tee -a $OUT
}
run base DENSEREG_NOP=1
run brn_blocks_1184 DENSEREG_BRN_BLOCKS=1184      # round-1 measured setting (reduce AND apply); default is now 296 for the reduce kernel only
run brn_blocks_592 DENSEREG_BRN_BLOCKS=592
run stats_per_cta DENSEREG_TC_STATS_PER_CTA=1
run wgrad_a_tmem DENSEREG_WGRAD_A_TMEM=1
run wgrad_a_tmem_swap DENSEREG_WGRAD_A_TMEM=1 DENSEREG_WGRAD_SWAP=1
run a_tmem_1 DENSEREG_TC_A_TMEM=1
run a_tmem_2 DENSEREG_TC_A_TMEM=2
run pool_bwd_v4 DENSEREG_POOL_BWD_V4=1
run pair_tail DENSEREG_TC_PAIR_TAIL=1
run wgrad_swap DENSEREG_WGRAD_SWAP=1
run wgrad_persist DENSEREG_WGRAD_PERSIST=1
run wgrad_persist_swap_w3 DENSEREG_WGRAD_PERSIST=1 DENSEREG_WGRAD_SWAP=1 DENSEREG_WGRAD_WAVES=3
run wgrad_persist_swap_w4 DENSEREG_WGRAD_PERSIST=1 DENSEREG_WGRAD_SWAP=1 DENSEREG_WGRAD_WAVES=4
timeout -s KILL 120 python tools/time_wgrad.py > gpurun_out/r2_wgrad_base.jsonl 2>&1
DENSEREG_WGRAD_PERSIST=1 DENSEREG_WGRAD_SWAP=1 timeout -s KILL 120 python tools/time_wgrad.py > gpurun_out/r2_wgrad_persist_swap.jsonl 2>&1
timeout -s KILL 90 python tools/layer_times.py > gpurun_out/r2_layer_times.txt 2>&1
DENSEREG_TC_A_TMEM=2 timeout -s KILL 60 python tools/time_layers.py > gpurun_out/r2_eval_layers_atmem.jsonl 2>&1
timeout -s KILL 60 python tools/time_layers.py > gpurun_out/r2_eval_layers_base.jsonl 2>&1
echo done | tee -a $OUT
