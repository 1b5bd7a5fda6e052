"""Non-default settings of the library (measured in round 2 and not adopted, or kept as options): each case re-runs the conv / network parity
tests in a child process with the switch set (the switches are read once per process).  Skipped unless DENSEREG_TEST_EXPERIMENTAL=1, so that
the default suite covers exactly what ships enabled; tools/r2_sweep.py additionally checks every setting against the fp32 engine and times it.
  DENSEREG_WGRAD_A_TMEM=1     persistent wgrad with the split A operand in tensor memory        wgrad_tc.cu
  DENSEREG_TC_A_TMEM=0 | 2    A-in-tensor-memory conv kernel nowhere / also instead of CTA pairs  conv_tc_atmem.cu
  DENSEREG_LANES=0            single stream instead of the lane plan                            engine.cu
  DENSEREG_WGRAD_STREAMS=1, DENSEREG_SIDE_STREAM=0   one / no filter-gradient side stream       engine.cu
  DENSEREG_PDL=0, DENSEREG_GRAD_ALIAS=0   no programmatic dependent launch / residual gradients copied instead of aliased   engine.cu
  DENSEREG_BRN_SMALL_ELEMS=0, DENSEREG_EW_REVERSE=0   no one-cluster BRN backward / BRN passes front to back               ew.cu
  DENSEREG_SPLIT_TRUNC=0      wgrad splitters rewrite hi = rn_tf32(v) as well (default: the landed fp32 tile is the hi operand)     wgrad_tc.cu
  DENSEREG_PIPELINE=2         micro-batch pipeline forced on for every training engine of the parity suites                     engine.cu
  DENSEREG_WGRAD_SWAP=0, DENSEREG_WGRAD_WAVES=2, DENSEREG_BRN_BLOCKS=1184, DENSEREG_TC_STATS_PER_CTA=0, DENSEREG_POOL_BWD_V4=0   the round-1 settings"""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("DENSEREG_TEST_EXPERIMENTAL") != "1",
                                                  reason="non-default settings: set DENSEREG_TEST_EXPERIMENTAL=1")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"DENSEREG_WGRAD_A_TMEM": "1"}, {"DENSEREG_TC_A_TMEM": "0"}, {"DENSEREG_TC_A_TMEM": "2"}, 
                                 {"DENSEREG_LANES": "0"}, {"DENSEREG_WGRAD_STREAMS": "1"}, {"DENSEREG_SIDE_STREAM": "0"},
                                 {"DENSEREG_WGRAD_SWAP": "0", "DENSEREG_WGRAD_WAVES": "2", "DENSEREG_BRN_BLOCKS": "1184", "DENSEREG_TC_STATS_PER_CTA": "0",
                                  "DENSEREG_POOL_BWD_V4": "0"},
                                 {"DENSEREG_PDL": "0"}, {"DENSEREG_GRAD_ALIAS": "0"}, {"DENSEREG_BRN_SMALL_ELEMS": "0", "DENSEREG_EW_REVERSE": "0"},
                                 {"DENSEREG_SPLIT_TRUNC": "0"}, {"DENSEREG_PIPELINE": "2"}])
def test_parity_suite_with_switch(env):
    e = dict(os.environ, **env)
    e.pop("DENSEREG_TEST_EXPERIMENTAL", None)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_conv.py", "tests/test_gpu_net.py", "-m", "gpu", "-x", "-q"], cwd=ROOT, env=e,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (env, r.stdout[-3000:])
