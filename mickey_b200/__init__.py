"""mickey_b200 — B200-native (sm_100a) implementation of the MicKey inference hot path.

Package contents (only what the path needs):
  csrc/      hand-written CUDA kernels + the C-ABI (include/mickey_b200.h)
  _lib.py    ctypes binding of libmickey_b200.so (fails loudly when the library is missing)
  engine.py  device-side pipeline: workspaces, packed weights, kernel sequencing
  model.py   host mirror of the reference's MickeyRelativePose / build_model surface
  weights.py synthetic (seeded) state dicts with the reference's tensor names + weight packing
  config.py  yacs-compatible config tree
  dist.py    pair sharding across ranks + the single all-gather of poses
"""
__version__ = "0.1.0"
