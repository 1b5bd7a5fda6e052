"""CPU oracle for the denseReg hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (NumPy fp32 / PyTorch-CPU fp32), the arithmetic of
melonwan/denseReg's hot path (network/um_v1.py, network/slim/ops.py,
model/hourglass_um_crop_tiny.py, data/preprocess.py, data/util.py) so that the sm_100a
kernels in densereg_b200/csrc can be checked against it.

PARITY UNPINNED: the reference is Python 2.7 + TensorFlow 1.3 graph code.  TensorFlow is
neither vendored in /root/reference nor installable here, and the reference ships no
tests, golden vectors, weights or input data for this path.  The
oracle therefore follows the reference *source* line by line (every function cites the
file:line it restates) plus the documented TF 1.x op semantics,
and is pinned only by its own property tests (tests/test_oracle_*.py), not by outputs of
the reference itself.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  Nothing under densereg_b200/ imports it; the product path fails
loudly when the CUDA library is missing instead of falling back to this code.
"""
