"""ctypes binding of oracle/_ref/libbifrost_ref.so (the REFERENCE's own headers behind a C API).

Test infrastructure only: used as the checker, never as the thing measured or shipped.
"""
import ctypes as C
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
LIB_PATH = REPO / "oracle" / "_ref" / "libbifrost_ref.so"

_lib = None


def available():
    return LIB_PATH.exists()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Ref:
    def __init__(self, lib):
        self.lib = lib
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        lib.ref_sizeof.argtypes = [C.c_char_p]
        lib.ref_get_tables.argtypes = [vp, vp, vp, vp]; lib.ref_get_tables.restype = None
        lib.ref_pcg2d.argtypes = [i64, vp, vp, vp]; lib.ref_pcg2d.restype = None
        lib.ref_sobol_sample4.argtypes = [i64, vp, vp, vp, vp, vp]; lib.ref_sobol_sample4.restype = None
        lib.ref_reverse_halton4.argtypes = [i32, vp]; lib.ref_reverse_halton4.restype = None
        lib.ref_sample02.argtypes = [i64, vp]; lib.ref_sample02.restype = None
        lib.ref_octahedral_encode_precise.argtypes = [i64, vp, vp]; lib.ref_octahedral_encode_precise.restype = None
        lib.ref_octahedral_decode.argtypes = [i64, vp, vp]; lib.ref_octahedral_decode.restype = None
        lib.ref_bsdf_eval_sample_pdf.argtypes = [i32, i64] + [vp] * 11 + [i32]; lib.ref_bsdf_eval_sample_pdf.restype = None
        lib.ref_default_shading_regularized.argtypes = [i64] + [vp] * 11; lib.ref_default_shading_regularized.restype = None
        lib.ref_light_sample_pdf_evaluate.argtypes = [i64, vp, i32, vp, vp, vp, vp, vp, vp]; lib.ref_light_sample_pdf_evaluate.restype = None
        lib.ref_balance_heuristic.argtypes = [i64, vp, vp, vp]; lib.ref_balance_heuristic.restype = None
        lib.ref_offset_ray_origin.argtypes = [i64, vp, vp, vp, vp]; lib.ref_offset_ray_origin.restype = None

    def sizeof(self, name):
        return self.lib.ref_sizeof(name.encode())

    def max_threads(self):
        return self.lib.ref_max_threads()

    def tables(self):
        a, b, c = (np.zeros(1024, np.float32) for _ in range(3))
        dims = np.zeros(6, np.int32)
        self.lib.ref_get_tables(_p(a), _p(b), _p(c), _p(dims))
        return a, b, c, dims

    def sobol_sample4(self, accumulation, pixel_hash, dimension):
        a, h, d = (np.ascontiguousarray(x, dtype=np.uint32).reshape(-1) for x in (accumulation, pixel_hash, dimension))
        n = a.shape[0]
        ui = np.empty((n, 4), np.uint32); f = np.empty((n, 4), np.float32)
        self.lib.ref_sobol_sample4(n, _p(a), _p(h), _p(d), _p(ui), _p(f))
        return ui, f

    def pcg2d(self, x, y):
        x, y = (np.ascontiguousarray(v, dtype=np.uint32).reshape(-1) for v in (x, y))
        out = np.empty((x.shape[0], 2), np.uint32)
        self.lib.ref_pcg2d(x.shape[0], _p(x), _p(y), _p(out))
        return out

    def reverse_halton4(self, n):
        out = np.empty((n, 4), np.float32)
        self.lib.ref_reverse_halton4(n, _p(out))
        return out

    def sample02(self, n):
        out = np.empty((n, 2), np.float32)
        self.lib.ref_sample02(n, _p(out))
        return out

    def bsdf_eval_sample_pdf(self, kind, wo, wi, tint, rms, u, coat=None, threads=0):
        wo, wi, tint, rms, u = (_f32(a).reshape(-1, 3) for a in (wo, wi, tint, rms, u))
        n = wo.shape[0]
        coat = None if coat is None else _f32(coat).reshape(-1, 2)
        ef, sf, sd = (np.empty((n, 3), np.float32) for _ in range(3))
        ep, sp = (np.empty(n, np.float32) for _ in range(2))
        self.lib.ref_bsdf_eval_sample_pdf(kind, n, _p(wo), _p(wi), _p(tint), _p(rms), _p(coat), _p(u), _p(ef), _p(ep), _p(sf), _p(sp), _p(sd), threads)
        return {"eval_f": ef, "eval_pdf": ep, "sample_f": sf, "sample_pdf": sp, "sample_dir": sd}

    def default_shading_regularized(self, materials, max_pdf_hint, wo, wi, u, tint_roughness_scale=None):
        m = np.ascontiguousarray(materials)
        n = m.shape[0]
        hint = _f32(max_pdf_hint).reshape(n); wo, wi, u = (_f32(a).reshape(n, 3) for a in (wo, wi, u))
        sc = None if tint_roughness_scale is None else _f32(tint_roughness_scale).reshape(n, 4)
        ef, sf, sd = (np.empty((n, 3), np.float32) for _ in range(3))
        ep, sp = (np.empty(n, np.float32) for _ in range(2))
        self.lib.ref_default_shading_regularized(n, _p(m), _p(sc), _p(hint), _p(wo), _p(wi), _p(u), _p(ef), _p(ep), _p(sf), _p(sp), _p(sd))
        return {"eval_f": ef, "eval_pdf": ep, "sample_f": sf, "sample_pdf": sp, "sample_dir": sd}

    def light_sample_pdf_evaluate(self, lights, position, u2, query_direction):
        l = np.ascontiguousarray(lights).reshape(-1)
        position, query_direction = _f32(position).reshape(-1, 3), _f32(query_direction).reshape(-1, 3)
        u2 = _f32(u2).reshape(-1, 2)
        n = position.shape[0]
        stride = 0 if l.shape[0] == 1 and n != 1 else 1
        samples = np.empty((n, 8), np.float32); pdf = np.empty(n, np.float32); rad = np.empty((n, 3), np.float32)
        self.lib.ref_light_sample_pdf_evaluate(n, _p(l), stride, _p(position), _p(u2), _p(query_direction), _p(samples), _p(pdf), _p(rad))
        return samples, pdf, rad

    def offset_ray_origin(self, origin, direction, normal):
        o, d, nrm = (_f32(a).reshape(-1, 3) for a in (origin, direction, normal))
        out = np.empty_like(o)
        self.lib.ref_offset_ray_origin(o.shape[0], _p(o), _p(d), _p(nrm), _p(out))
        return out


def load():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle` where /root/reference is available")
        _lib = Ref(C.CDLL(str(LIB_PATH)))
    return _lib
