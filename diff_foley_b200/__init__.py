"""diff_foley_b200 -- B200-native (sm_100a) implementation of Diff-Foley's DDIM sampling hot path.

Everything that computes lives in csrc/libdfb.so (hand-written CUDA, C ABI in include/dfb.h);
this package is the host-side mirror of the reference's plugin interface (UNetModel / DDIMSampler).
"""
__version__ = "0.1.0"
