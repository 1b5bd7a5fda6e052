"""densereg_b200 -- B200-native (sm_100a) engine for the denseReg hand-pose hot path.

Host side of the C-ABI in include/densereg.h: PyTorch tensors are used only as device buffers /
streams; all arithmetic happens in libdensereg_sm100.so (hand-written CUDA).  There is no CPU or
PyTorch fallback: importing the engine without the built library raises.
"""
__version__ = "0.1.0"

import os as _os

# The engine spreads one training step over ~15 CUDA streams (lanes, filter-gradient side streams, the two arenas of the micro-batch
# pipeline, the communication stream) next to the caller's own copy streams.  The driver maps streams onto CUDA_DEVICE_MAX_CONNECTIONS
# hardware work queues (default 8); streams that share a queue serialise falsely -- measured: host-to-device input copies issued while
# the pipelined step runs cost 2.7 ms per optimiser step at batch 40 (10 % at batch 8) with 8 queues and nothing with 16 or 32
# (profiles/r2_final.md section 8).  The variable is read when the CUDA context is created, so it is set here, at import time, unless the
# user has chosen a value.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
