"""dv_matcher_b200: B200-native (sm_100a) dense correspondence + deformation hot path of DV-Matcher.

Host code is Python/PyTorch and mirrors the reference's call surface (models/loss.py,
lib/deformation_graph_point.py); the arithmetic runs in hand-written CUDA kernels behind the C ABI of
include/dvm_b200.h (libdvm_b200.so, loaded with ctypes).  No Triton, no dispatch, no CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["ops", "synthetic"]
__version__ = "0.1.0"
