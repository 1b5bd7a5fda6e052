"""densereg_b200 -- B200-native (sm_100a) engine for the denseReg hand-pose hot path.

Host side of the C-ABI in include/densereg.h: PyTorch tensors are used only as device buffers /
streams; all arithmetic happens in libdensereg_sm100.so (hand-written CUDA).  There is no CPU or
PyTorch fallback: importing the engine without the built library raises.
"""
__version__ = "0.1.0"
