"""ctypes binding of oracle/_ref/libbifrost_ref.so (the REFERENCE's own headers behind a C API).

Test infrastructure only: used as the checker, never as the thing measured or shipped.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ["BPT_ORACLE_LIB"]) if os.environ.get("BPT_ORACLE_LIB") else REPO / "oracle" / "_ref" / "libbifrost_ref.so"  # the override: a sanitizer build

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
        lib.ref_get_dielectric_tables.argtypes = [vp, vp, vp]; lib.ref_get_dielectric_tables.restype = None
        lib.ref_sample_dielectric_rho.argtypes = [i64, vp, vp, vp, vp]; lib.ref_sample_dielectric_rho.restype = None
        lib.ref_tonemap.argtypes = [i32, C.c_float, vp, i64, vp, vp]; lib.ref_tonemap.restype = None
        lib.ref_linear_to_srgb.argtypes = [i64, vp, vp]; lib.ref_linear_to_srgb.restype = None
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

    def dielectric_tables(self):
        a, b = np.zeros(8192, np.float32), np.zeros(8192, np.float32)
        dims = np.zeros(3, np.int32)
        self.lib.ref_get_dielectric_tables(_p(a), _p(b), _p(dims))
        return a, b, dims

    def tonemap(self, mode, rgb, exposure=1.0, filmic=(0.0, 0.53, 0.91, 0.23, 0.035)):
        rgb = _f32(rgb).reshape(-1, 3); out = np.empty_like(rgb); s = _f32(filmic)
        self.lib.ref_tonemap(int(mode), float(exposure), _p(s), rgb.shape[0], _p(rgb), _p(out))
        return out

    def linear_to_srgb(self, values):
        v = _f32(values).reshape(-1); out = np.empty_like(v)
        self.lib.ref_linear_to_srgb(v.shape[0], _p(v), _p(out))
        return out

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


class OracleScene:
    """CPU oracle scene (oracle/pt_oracle.cpp): flattening, brute force / BVH intersection and the integrator restatement."""

    def __init__(self, scene):
        self.ref = load()
        lib = self.ref.lib
        vp, i64, i32, u32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint32
        lib.pto_scene_create.restype = vp
        lib.pto_scene_destroy.argtypes = [vp]; lib.pto_scene_destroy.restype = None
        lib.pto_scene_add_mesh.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32]; lib.pto_scene_add_mesh.restype = None
        for name in ("pto_scene_set_instances", "pto_scene_set_materials", "pto_scene_set_lights"):
            getattr(lib, name).argtypes = [vp, vp, i32]; getattr(lib, name).restype = None
        lib.pto_scene_set_environment.argtypes = [vp, vp, vp, i32, i32, vp, i32, i32, vp, i32]; lib.pto_scene_set_environment.restype = None
        lib.pto_scene_set_environment_cdfs.argtypes = [vp, vp, vp, i32, i32, i32]; lib.pto_scene_set_environment_cdfs.restype = None
        lib.pto_scene_set_mesh_emission.argtypes = [vp, i32, vp, i32]; lib.pto_scene_set_mesh_emission.restype = None
        lib.pto_scene_set_mesh_texcoords.argtypes = [vp, i32, vp, i32]; lib.pto_scene_set_mesh_texcoords.restype = None
        lib.pto_scene_add_texture.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]; lib.pto_scene_add_texture.restype = None
        lib.pto_texture_sample.argtypes = [vp, i32, i64, vp, vp]; lib.pto_texture_sample.restype = None
        lib.pto_scene_build.argtypes = [vp]; lib.pto_scene_build.restype = None
        lib.pto_triangle_count.argtypes = [vp]; lib.pto_triangle_count.restype = i64
        lib.pto_world_vertices.argtypes = [vp, vp]; lib.pto_world_vertices.restype = None
        lib.pto_intersect.argtypes = [vp, i64, vp, vp, vp, vp, i32, vp, vp, vp, vp]; lib.pto_intersect.restype = None
        lib.pto_render.argtypes = [vp, vp, vp, i32, i32, u32, u32, i32, i32, vp, vp, i32]; lib.pto_render.restype = None
        self.lib = lib
        self.h = C.c_void_p(lib.pto_scene_create())
        for mesh_id, m in scene["meshes"].items():
            idx = np.ascontiguousarray(m["indices"], np.uint32); pos = _f32(m["positions"])
            nrm = None if m.get("normals") is None else _f32(m["normals"])
            tints = None if m.get("tints") is None else np.ascontiguousarray(m["tints"], np.uint8)
            lib.pto_scene_add_mesh(self.h, int(mesh_id), _p(idx), idx.shape[0], _p(pos), _p(nrm), _p(tints), pos.shape[0])
            if m.get("emission") is not None:
                em = _f32(m["emission"])
                lib.pto_scene_set_mesh_emission(self.h, int(mesh_id), _p(em), em.shape[0])
            if m.get("texcoords") is not None:
                uv = _f32(m["texcoords"])
                lib.pto_scene_set_mesh_texcoords(self.h, int(mesh_id), _p(uv), uv.shape[0])
        from bifrost3d_b200 import capi
        for texture_id, tex in scene.get("textures", {}).items():
            fmt, pixels = capi.pixel_format_of(tex["pixels"])
            lib.pto_scene_add_texture(self.h, int(texture_id), pixels.shape[1], pixels.shape[0], fmt, int(tex.get("srgb", False)),
                                      int(tex.get("wrap_u", capi.WRAP_REPEAT)), int(tex.get("wrap_v", capi.WRAP_REPEAT)), int(tex.get("linear", True)), _p(pixels))
        mats = np.ascontiguousarray(scene["materials"]); inst = np.ascontiguousarray(scene["instances"]); lights = np.ascontiguousarray(scene["lights"])
        lib.pto_scene_set_materials(self.h, _p(mats), mats.shape[0])
        lib.pto_scene_set_instances(self.h, _p(inst), inst.shape[0])
        lib.pto_scene_set_lights(self.h, _p(lights) if lights.size else None, lights.shape[0])
        env = scene.get("environment", {"tint": (0, 0, 0)})
        tint = _f32(env["tint"])
        if env.get("texels") is None:
            lib.pto_scene_set_environment(self.h, _p(tint), None, 0, 0, None, 0, 0, None, 0)
        else:
            tex = _f32(env["texels"]); pdf = _f32(env["per_pixel_pdf"]); s = np.ascontiguousarray(env["samples"])
            lib.pto_scene_set_environment(self.h, _p(tint), _p(tex), tex.shape[1], tex.shape[0], _p(pdf), pdf.shape[1], pdf.shape[0], _p(s), s.shape[0])
            if env.get("marginal_cdf") is not None and env["conditional_cdf"].shape[1] - 1 == pdf.shape[1]:
                m = _f32(env["marginal_cdf"]); c = _f32(env["conditional_cdf"])
                lib.pto_scene_set_environment_cdfs(self.h, _p(m), _p(c), pdf.shape[1], pdf.shape[0], int(env.get("nee", "presampled") == "cdf"))
        lib.pto_scene_build(self.h)

    def close(self):
        if self.h:
            self.lib.pto_scene_destroy(self.h); self.h = None

    def texture_sample(self, texture_id, uv):
        uv = _f32(uv).reshape(-1, 2)
        out = np.empty((uv.shape[0], 4), np.float32)
        self.lib.pto_texture_sample(self.h, int(texture_id), uv.shape[0], _p(uv), _p(out))
        return out

    def triangle_count(self):
        return self.lib.pto_triangle_count(self.h)

    def world_vertices(self):
        out = np.empty((self.triangle_count(), 3, 3), np.float32)
        self.lib.pto_world_vertices(self.h, _p(out))
        return out

    def intersect(self, origins, directions, tmin=None, tmax=None, brute=False):
        o, d = _f32(origins).reshape(-1, 3), _f32(directions).reshape(-1, 3)
        n = o.shape[0]
        tmin = np.zeros(n, np.float32) if tmin is None else _f32(np.broadcast_to(tmin, (n,)))
        tmax = np.full(n, 1e30, np.float32) if tmax is None else _f32(np.broadcast_to(tmax, (n,)))
        prim = np.empty(n, np.int32); t = np.empty(n, np.float32); uv = np.empty((n, 2), np.float32); occ = np.empty(n, np.uint8)
        self.lib.pto_intersect(self.h, n, _p(o), _p(d), _p(tmin), _p(tmax), int(brute), _p(prim), _p(t), _p(uv), _p(occ))
        return prim, t, uv, occ

    def render(self, camera, width, height, first_sample, sample_count, max_bounces=4, nee_samples=3, pdf_scale=0.5, rows=None, threads=0, accum=None,
               russian_roulette_start=0):
        from bifrost3d_b200 import capi
        cam = capi.make_camera(*camera)
        s = capi.Settings(max_bounces, nee_samples, pdf_scale, russian_roulette_start)
        if accum is None:
            accum = np.zeros((height, width, 4), np.float64)
        counters = np.zeros(2, np.uint64)
        r0, r1 = (0, height) if rows is None else rows
        self.lib.pto_render(self.h, C.byref(cam), C.byref(s), width, height, first_sample, sample_count, r0, r1, _p(accum), _p(counters), threads)
        return accum, counters


def reference_environment(texels, sample_count=8192):
    """The reference's own environment build (InfiniteAreaLight + PresampledEnvironmentMap) through oracle/ref_api.cpp."""
    lib = load().lib
    lib.ref_environment_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_environment_build.restype = C.c_int
    tex = _f32(texels); h, w = tex.shape[:2]
    pw, ph = C.c_int(), C.c_int(); integral = C.c_float()
    pdf = np.zeros(w * max(h, 128), np.float32)
    n = max(2, 1 << int(np.ceil(np.log2(max(sample_count, 1)))))
    samples = np.zeros((n, 8), np.float32)
    produced = lib.ref_environment_build(_p(tex), w, h, sample_count, C.byref(pw), C.byref(ph), _p(pdf), _p(samples), C.byref(integral))
    return {"per_pixel_pdf": pdf[: pw.value * ph.value].reshape(ph.value, pw.value), "samples": samples[:produced], "integral": integral.value}


def reference_environment_sample(texels, points):
    """InfiniteAreaLight::sample at `points` (n x 2 in [0, 1)) and InfiniteAreaLight::PDF of the sampled directions, plus the
    reference's own marginal / conditional CDFs of the map (oracle/ref_api.cpp: ref_environment_sample)."""
    lib = load().lib
    lib.ref_environment_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 5
    lib.ref_environment_sample.restype = C.c_int
    tex = _f32(texels); h, w = tex.shape[:2]
    assert h >= 128, "lower maps are resampled by the reference (InfiniteAreaLight.cpp:45-52)"
    pts = _f32(points).reshape(-1, 2); n = pts.shape[0]
    samples = np.zeros((n, 8), np.float32); pdf = np.zeros(n, np.float32)
    marginal = np.zeros(h + 1, np.float32); conditional = np.zeros((h, w + 1), np.float32)
    lib.ref_environment_sample(_p(tex), w, h, n, _p(pts), _p(samples), _p(pdf), _p(marginal), _p(conditional))
    return {"samples": samples, "pdf_of_direction": pdf, "marginal_cdf": marginal, "conditional_cdf": conditional}


def reference_compare_images(reference, target, mssim_support, diff_images=False):
    """ImageOperations::Compare::{rms, ssim, mssim} of the reference (oracle/ref_api.cpp: ref_compare_images) on (H, W, 4) floats."""
    lib = load().lib
    lib.ref_compare_images.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    lib.ref_compare_images.restype = C.c_int
    a, b = _f32(reference), _f32(target)
    h, w = a.shape[:2]
    rms, ssim, mssim = C.c_float(), C.c_float(), C.c_float()
    rms_diff = np.zeros_like(a) if diff_images else None
    mssim_diff = np.zeros_like(a) if diff_images else None
    lib.ref_compare_images(w, h, _p(a), _p(b), int(mssim_support), C.byref(rms), C.byref(ssim), C.byref(mssim), _p(rms_diff), _p(mssim_diff))
    out = {"rms": rms.value, "ssim": ssim.value, "mssim": mssim.value}
    if diff_images:
        out["rms_diff"] = rms_diff; out["mssim_diff"] = mssim_diff
    return out
