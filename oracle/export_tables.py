#!/usr/bin/env python3
"""Export the reference's precomputed rho / alpha tables as a data file.

The tables are DATA the renderer consumes (the reference uploads them as textures,
Renderer.cpp:400-466; a host integrating libbpt.so passes its own
Bifrost::Assets::Shading arrays to bpt_set_tables). For the Python-driven tests and bench
we keep a binary copy: 3 x 32 x 32 float32 = [GGX_with_fresnel | GGX | estimate_alpha],
read out of the staged reference through oracle/_ref/libbifrost_ref.so. The dielectric GGX rho tables
(2 x 16 x 16 x 16 float2, Fittings.h:36-46) go to dielectric_tables.bin the same way.
"""
import ctypes
from pathlib import Path
import numpy as np

REPO = Path(__file__).resolve().parent.parent
lib = ctypes.CDLL(str(REPO / "oracle/_ref/libbifrost_ref.so"))
a = np.zeros(1024, np.float32); b = np.zeros(1024, np.float32); c = np.zeros(1024, np.float32)
dims = (ctypes.c_int * 6)()
fp = ctypes.POINTER(ctypes.c_float)
lib.ref_get_tables(a.ctypes.data_as(fp), b.ctypes.data_as(fp), c.ctypes.data_as(fp), dims)
assert list(dims) == [32] * 6, list(dims)
out = REPO / "bifrost3d_b200/data/shading_tables.bin"
np.concatenate([a, b, c]).astype("<f4").tofile(out)
print("wrote", out, out.stat().st_size, "bytes")

light = np.zeros(8192, np.float32); dense = np.zeros(8192, np.float32)
dims3 = (ctypes.c_int * 3)()
lib.ref_get_dielectric_tables(light.ctypes.data_as(fp), dense.ctypes.data_as(fp), dims3)
assert list(dims3) == [16] * 3, list(dims3)
out = REPO / "bifrost3d_b200/data/dielectric_tables.bin"
np.concatenate([light, dense]).astype("<f4").tofile(out)
print("wrote", out, out.stat().st_size, "bytes")
