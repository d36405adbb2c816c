"""pointnav-vo_b200: B200-native (sm_100a) implementation of PointNav-VO's data-parallel hot path.

Layout:
  csrc/      hand-written CUDA kernels + the C-ABI (include/pnvo.h) built into csrc/libpnvo.so
  lib.py     ctypes binding of the C-ABI (fails loudly when the library is missing)
  vo/, rl/, model_utils/, utils/   host-side mirror of the reference's module surface
"""
__version__ = "0.1.0"
