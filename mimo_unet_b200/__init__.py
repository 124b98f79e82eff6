"""mimo_unet_b200 -- B200-native (sm_100a) kernels and executor for the MIMO U-Net hot path.

Layout:
  csrc/            hand-written CUDA (tcgen05/TMA convolutions, fused memory-bound kernels, C++ executor)
  libmimo_b200.so  the built C-ABI library (include/mimo_b200.h)
  _lib.py          ctypes binding
  engine.py        UNetPlan: workspace + whole-network forward/backward
  functional.py    tensor-level wrappers of the loss / aggregation / loss-buffer kernels
The reference-compatible python surface (``mimo.models...``) lives in the top-level ``mimo`` package.
"""
from ._lib import MimoError, build, lib  # noqa: F401
