"""bifrost3d_b200 - B200-native path tracer behind Bifrost3D's OptiXRenderer API.

The product is `libbpt.so` (hand-written sm_100a CUDA behind the C ABI in include/bpt_c_api.h).
This package is the thin Python host side used by the tests and bench.py: a ctypes binding
(`capi`) and procedural scene builders (`scenes`). It never falls back to a CPU implementation.
"""
from .capi import Bpt, BptError, load_library, library_path  # noqa: F401
