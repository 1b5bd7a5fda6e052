"""Opt-in kernels that have NOT been measured / verified on a B200 yet (written at the end of round 1 when the GPU budget was spent):
  DENSEREG_WGRAD_SWAP=2      wgrad with exchanged operand roles (M = cout, coalesced reductions)       wgrad_tc.cu
  DENSEREG_WGRAD_A_TMEM=1    persistent wgrad with the split A operand in tensor memory                    wgrad_tc.cu
  DENSEREG_WGRAD_PERSIST=1   persistent wgrad kernel with double-buffered TMEM accumulators            wgrad_tc.cu
  DENSEREG_TC_STATS_PER_CTA=1  fused BRN statistics accumulated per CTA (one fence + counter per CTA)   conv_tc_epilogue.cuh
  DENSEREG_TC_PAIR_TAIL=1    pair conv kernel: last wave's items sliced along N over all clusters         conv_tc_pair.cu
  DENSEREG_TC_A_TMEM=1|2     3xTF32 conv with the split A operand in tensor memory (2: also instead of pairs)  conv_tc_atmem.cu
  DENSEREG_POOL_BWD_V4=1     float4-over-channels max-pool backward                                          ew.cu
  DENSEREG_BRN_BLOCKS=1184   the round-1 grid cap of the BRN-backward kernels (default now 296 for the reduce)   ew.cu
Each case re-runs the existing conv / network parity tests in a child process with the switch set (the switches are read once per
process).  Skipped unless DENSEREG_TEST_EXPERIMENTAL=1, so that the default suite only covers what ships enabled."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("DENSEREG_TEST_EXPERIMENTAL") != "1",
                                                  reason="experimental kernels: set DENSEREG_TEST_EXPERIMENTAL=1")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"DENSEREG_WGRAD_SWAP": "2"}, {"DENSEREG_WGRAD_SWAP": "1"}, {"DENSEREG_WGRAD_PERSIST": "1"},
                                 {"DENSEREG_WGRAD_PERSIST": "1", "DENSEREG_WGRAD_SWAP": "1", "DENSEREG_WGRAD_WAVES": "4"},
                                 {"DENSEREG_TC_STATS_PER_CTA": "1"}, {"DENSEREG_BRN_BLOCKS": "1184"}, {"DENSEREG_TC_PAIR_TAIL": "1"}, {"DENSEREG_POOL_BWD_V4": "1"}, {"DENSEREG_TC_A_TMEM": "2"}, {"DENSEREG_TC_A_TMEM": "1"}, {"DENSEREG_WGRAD_A_TMEM": "1"},
                                 {"DENSEREG_WGRAD_A_TMEM": "1", "DENSEREG_WGRAD_SWAP": "1"}])
def test_parity_suite_with_switch(env):
    e = dict(os.environ, **env)
    e.pop("DENSEREG_TEST_EXPERIMENTAL", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_conv.py", "tests/test_gpu_net.py", "-m", "gpu", "-x", "-q"], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (env, r.stdout[-3000:])
