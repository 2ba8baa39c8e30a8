"""ctypes binding of include/bpt_c_api.h. Arrays are numpy; nothing here computes anything."""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB = None


class BptError(RuntimeError):
    pass


def library_path() -> Path:
    # BPT_LIB selects an alternative build of the same library (kernel tuning experiments); the default is the product.
    return Path(os.environ["BPT_LIB"]) if os.environ.get("BPT_LIB") else _PKG / "libbpt.so"


class Material(C.Structure):  # Types.h:353-416
    _fields_ = [("flags", C.c_uint16), ("shading_model", C.c_uint16), ("tint", C.c_float * 3), ("roughness", C.c_float),
                ("tint_roughness_texture_id", C.c_int32), ("roughness_texture_id", C.c_int32), ("specularity", C.c_float),
                ("metallic", C.c_float), ("metallic_texture_id", C.c_int32), ("coverage", C.c_float),
                ("coverage_texture_id", C.c_int32), ("emission", C.c_float * 3), ("coat", C.c_uint16), ("coat_roughness", C.c_uint16)]


class Light(C.Structure):  # Types.h:290-312
    _fields_ = [("data", C.c_float * 11), ("flags", C.c_uint32)]


class LightSample(C.Structure):  # Types.h:210-222
    _fields_ = [("radiance", C.c_float * 3), ("pdf", C.c_float), ("direction_to_light", C.c_float * 3), ("distance", C.c_float)]


class Instance(C.Structure):
    _fields_ = [("mesh_id", C.c_int32), ("material_id", C.c_int32), ("to_world", C.c_float * 12)]


class Camera(C.Structure):
    _fields_ = [("view_to_world_rotation", C.c_float * 9), ("inverse_projection", C.c_float * 16), ("inverse_view_projection", C.c_float * 16)]


class Settings(C.Structure):
    _fields_ = [("max_bounce_count", C.c_uint32), ("next_event_sample_count", C.c_int32), ("path_regularization_pdf_scale", C.c_float), ("russian_roulette_start_bounce", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("extend_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("samples", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("extend_ms", C.c_float), ("shadow_ms", C.c_float), ("shade_ms", C.c_float), ("other_ms", C.c_float),
                ("extend_node_visits", C.c_uint64), ("extend_triangle_tests", C.c_uint64), ("nonfinite_samples", C.c_uint64), ("traversal_stack_overflows", C.c_uint64), ("iterations", C.c_uint64)]


assert C.sizeof(Material) == 64 and C.sizeof(Light) == 48 and C.sizeof(LightSample) == 32

MATERIAL_DTYPE = np.dtype([("flags", "<u2"), ("shading_model", "<u2"), ("tint", "<f4", 3), ("roughness", "<f4"),
                           ("tint_roughness_texture_id", "<i4"), ("roughness_texture_id", "<i4"), ("specularity", "<f4"),
                           ("metallic", "<f4"), ("metallic_texture_id", "<i4"), ("coverage", "<f4"), ("coverage_texture_id", "<i4"),
                           ("emission", "<f4", 3), ("coat", "<u2"), ("coat_roughness", "<u2")])
LIGHT_DTYPE = np.dtype([("data", "<f4", 11), ("flags", "<u4")])
LIGHT_SAMPLE_DTYPE = np.dtype([("radiance", "<f4", 3), ("pdf", "<f4"), ("direction_to_light", "<f4", 3), ("distance", "<f4")])
INSTANCE_DTYPE = np.dtype([("mesh_id", "<i4"), ("material_id", "<i4"), ("to_world", "<f4", 12)])
assert MATERIAL_DTYPE.itemsize == 64 and LIGHT_DTYPE.itemsize == 48 and LIGHT_SAMPLE_DTYPE.itemsize == 32 and INSTANCE_DTYPE.itemsize == 56

LIGHT_SPHERE, LIGHT_DIRECTIONAL, LIGHT_ENVIRONMENT, LIGHT_PRESAMPLED_ENVIRONMENT, LIGHT_SPOT = 1, 2, 3, 4, 5
ENVIRONMENT_NEE_PRESAMPLED, ENVIRONMENT_NEE_CDF = 0, 1

EXPORTS = ["bpt_create", "bpt_destroy", "bpt_last_error", "bpt_stream", "bpt_set_tables", "bpt_set_dielectric_tables", "bpt_upload_texture", "bpt_destroy_texture", "bpt_texture_sample", "bpt_upload_mesh", "bpt_set_mesh_emission", "bpt_remove_mesh", "bpt_set_instances",
           "bpt_set_materials", "bpt_set_lights", "bpt_set_environment", "bpt_set_environment_cdfs", "bpt_set_environment_sampling", "bpt_set_hit_sorting", "bpt_build_accel", "bpt_accel_info", "bpt_accel_hierarchy", "bpt_read_accumulation", "bpt_write_accumulation", "bpt_render", "bpt_render_aov",
           "bpt_accumulation_device_ptr", "bpt_select_accumulation", "bpt_release_accumulation", "bpt_comm_unique_id", "bpt_comm_init", "bpt_comm_destroy", "bpt_comm_check", "bpt_reduce_accumulation", "bpt_resolve_half4", "bpt_resolve_half4_async", "bpt_wait_frame", "bpt_resolve_float4", "bpt_resolve_tonemapped", "bpt_tonemap_colors", "bpt_synchronize", "bpt_set_profiling", "bpt_get_counters",
           "bpt_bsdf_eval_sample_pdf", "bpt_default_shading_regularized", "bpt_light_sample_pdf_evaluate", "bpt_rng_sample4",
           "bpt_intersect", "bpt_sort_pairs", "bpt_exclusive_scan", "bpt_compare_images"]


class TonemapSettings(C.Structure):
    _fields_ = [("mode", C.c_int32), ("exposure", C.c_float), ("black_clip", C.c_float), ("toe", C.c_float), ("slope", C.c_float),
                ("shoulder", C.c_float), ("white_clip", C.c_float), ("reserved", C.c_int32)]


TONEMAP = {"linear": 0, "filmic": 1, "agx": 2, "khronos_neutral": 3}
FILMIC_ACES = (0.0, 0.53, 0.91, 0.23, 0.035)  # TonemappingSettings::ACES(): black_clip, toe, slope, shoulder, white_clip


def tonemap_settings(mode="filmic", exposure=1.0, filmic=FILMIC_ACES):
    return TonemapSettings(TONEMAP[mode] if isinstance(mode, str) else int(mode), exposure, *filmic, 0)


class TextureDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("pixel_format", C.c_int32), ("is_srgb", C.c_int32),
                ("wrap_u", C.c_int32), ("wrap_v", C.c_int32), ("linear_filter", C.c_int32), ("reserved", C.c_int32)]


PIXEL_ALPHA8, PIXEL_RGB24, PIXEL_RGBA32, PIXEL_RGB_FLOAT, PIXEL_RGBA_FLOAT = 1, 3, 4, 6, 7
WRAP_CLAMP, WRAP_REPEAT = 0, 1


def pixel_format_of(pixels):
    """(format, contiguous array) for an image array of shape (H, W) / (H, W, C), uint8 or float32."""
    p = np.asarray(pixels)
    if p.ndim == 2:
        p = p[..., None]
    channels = p.shape[2]
    if p.dtype == np.uint8:
        fmt = {1: PIXEL_ALPHA8, 3: PIXEL_RGB24, 4: PIXEL_RGBA32}[channels]
    else:
        p = p.astype(np.float32)
        fmt = {3: PIXEL_RGB_FLOAT, 4: PIXEL_RGBA_FLOAT}[channels]
    return fmt, np.ascontiguousarray(p)


def load_library():
    """Loads libbpt.so. Fails loudly if the CUDA extension has not been built: there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not path.exists():
        raise BptError(f"{path} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(path))
    vp, i32, i64, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32
    lib.bpt_create.argtypes = [i32, C.POINTER(vp)]
    lib.bpt_destroy.argtypes = [vp]; lib.bpt_destroy.restype = None
    lib.bpt_last_error.argtypes = [vp]; lib.bpt_last_error.restype = C.c_char_p
    lib.bpt_stream.argtypes = [vp]; lib.bpt_stream.restype = vp
    lib.bpt_set_tables.argtypes = [vp, vp, vp, vp]
    lib.bpt_set_dielectric_tables.argtypes = [vp, vp, vp]
    lib.bpt_upload_texture.argtypes = [vp, i32, C.POINTER(TextureDesc), vp]
    lib.bpt_destroy_texture.argtypes = [vp, i32]
    lib.bpt_texture_sample.argtypes = [vp, i32, i64, vp, vp]
    lib.bpt_upload_mesh.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, i32]
    lib.bpt_set_mesh_emission.argtypes = [vp, i32, vp, i32]
    lib.bpt_remove_mesh.argtypes = [vp, i32]
    lib.bpt_set_instances.argtypes = [vp, vp, i32]
    lib.bpt_set_materials.argtypes = [vp, vp, i32]
    lib.bpt_set_lights.argtypes = [vp, vp, i32]
    lib.bpt_set_environment.argtypes = [vp, vp, vp, i32, i32, vp, i32, i32, vp, i32]
    lib.bpt_set_environment_cdfs.argtypes = [vp, vp, vp, i32, i32]
    lib.bpt_set_environment_sampling.argtypes = [vp, i32]
    lib.bpt_set_hit_sorting.argtypes = [vp, i32]
    lib.bpt_build_accel.argtypes = [vp]
    lib.bpt_accel_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_float)]
    lib.bpt_accel_hierarchy.argtypes = [vp, C.POINTER(i32), C.POINTER(i64), C.POINTER(i32)]
    lib.bpt_read_accumulation.argtypes = [vp, vp, C.POINTER(i32), C.POINTER(i32)]
    lib.bpt_write_accumulation.argtypes = [vp, i32, i32, vp]
    lib.bpt_render.argtypes = [vp, C.POINTER(Camera), C.POINTER(Settings), i32, i32, u32, u32, i32]
    lib.bpt_render_aov.argtypes = [vp, C.POINTER(Camera), i32, i32, i32, u32, u32, i32]
    lib.bpt_accumulation_device_ptr.argtypes = [vp]; lib.bpt_accumulation_device_ptr.restype = vp
    lib.bpt_comm_unique_id.argtypes = [vp]
    lib.bpt_comm_init.argtypes = [vp, vp, i32, i32]
    lib.bpt_comm_destroy.argtypes = [vp]
    lib.bpt_comm_check.argtypes = [vp]
    lib.bpt_reduce_accumulation.argtypes = [vp, i32]
    lib.bpt_compare_images.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.bpt_sort_pairs.argtypes = [vp, i64, vp, vp, i32, i32]
    lib.bpt_exclusive_scan.argtypes = [vp, i64, vp, vp, vp]
    lib.bpt_select_accumulation.argtypes = [vp, i32]
    lib.bpt_release_accumulation.argtypes = [vp, i32]
    lib.bpt_resolve_half4.argtypes = [vp, vp, i32]
    lib.bpt_resolve_half4_async.argtypes = [vp, vp, i32]
    lib.bpt_wait_frame.argtypes = [vp, i32]
    lib.bpt_resolve_float4.argtypes = [vp, vp]
    lib.bpt_resolve_tonemapped.argtypes = [vp, C.POINTER(TonemapSettings), vp, i32]
    lib.bpt_tonemap_colors.argtypes = [vp, C.POINTER(TonemapSettings), i64, vp, vp]
    lib.bpt_synchronize.argtypes = [vp]
    lib.bpt_set_profiling.argtypes = [vp, i32]
    lib.bpt_get_counters.argtypes = [vp, C.POINTER(Counters), i32]
    lib.bpt_bsdf_eval_sample_pdf.argtypes = [vp, i32, i64] + [vp] * 11 + [i32]
    lib.bpt_default_shading_regularized.argtypes = [vp, i64] + [vp] * 11
    lib.bpt_light_sample_pdf_evaluate.argtypes = [vp, i64, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.bpt_rng_sample4.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.bpt_intersect.argtypes = [vp, i64] + [vp] * 8
    for name in EXPORTS:
        if getattr(lib, name).restype is C.c_int:
            pass
    _LIB = lib
    return lib


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_tables():
    t = np.fromfile(_PKG / "data" / "shading_tables.bin", dtype="<f4")
    assert t.size == 3 * 1024
    return t[:1024].copy(), t[1024:2048].copy(), t[2048:].copy()


def default_dielectric_tables():
    """[into_light_medium | into_dense_medium], each 16^3 {total_rho, reflected_rho} pairs (oracle/export_tables.py)."""
    t = np.fromfile(_PKG / "data" / "dielectric_tables.bin", dtype="<f4")
    assert t.size == 2 * 8192
    return t[:8192].copy(), t[8192:].copy()


class Bpt:
    """One path tracer context on one CUDA device (mirrors OptiXRenderer::Renderer::initialize)."""

    def __init__(self, device: int = 0, tables=True):
        self.lib = load_library()
        handle = C.c_void_p()
        status = self.lib.bpt_create(device, C.byref(handle))
        if status != 0:
            raise BptError(f"bpt_create(device={device}) failed with status {status}: no usable CUDA device. There is no CPU fallback.")
        self.h = handle
        if tables:
            self.set_tables(*default_tables())
            self.set_dielectric_tables(*default_dielectric_tables())

    def close(self):
        if getattr(self, "h", None):
            self.lib.bpt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status):
        if status != 0:
            raise BptError(f"status {status}: {self.lib.bpt_last_error(self.h).decode()}")

    @property
    def stream(self):
        return self.lib.bpt_stream(self.h)

    def synchronize(self):
        self._check(self.lib.bpt_synchronize(self.h))

    def set_profiling(self, enabled):
        self._check(self.lib.bpt_set_profiling(self.h, int(enabled)))

    def set_tables(self, ggx_with_fresnel, ggx, alpha):
        a, b, c = _f32(ggx_with_fresnel), _f32(ggx), _f32(alpha)
        assert a.size == b.size == c.size == 1024
        self._check(self.lib.bpt_set_tables(self.h, _ptr(a), _ptr(b), _ptr(c)))

    def set_dielectric_tables(self, into_light_medium, into_dense_medium):
        a, b = _f32(into_light_medium), _f32(into_dense_medium)
        assert a.size == b.size == 8192
        self._check(self.lib.bpt_set_dielectric_tables(self.h, _ptr(a), _ptr(b)))

    # ---- scene ----
    def upload_mesh(self, mesh_id, indices, positions, normals=None, texcoords=None, tint_roughness=None):
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        pos = _f32(positions).reshape(-1, 3)
        nrm = None if normals is None else _f32(normals).reshape(-1, 3)
        uv = None if texcoords is None else _f32(texcoords).reshape(-1, 2)
        tr = None if tint_roughness is None else np.ascontiguousarray(tint_roughness, dtype=np.uint8).reshape(-1, 4)
        for a in (nrm, uv, tr):
            assert a is None or a.shape[0] == pos.shape[0]
        self._check(self.lib.bpt_upload_mesh(self.h, mesh_id, _ptr(idx), idx.shape[0], _ptr(pos), _ptr(nrm), _ptr(uv), _ptr(tr), pos.shape[0]))

    def upload_texture(self, texture_id, pixels, srgb=False, wrap_u=WRAP_REPEAT, wrap_v=WRAP_REPEAT, linear=True):
        fmt, p = pixel_format_of(pixels)
        desc = TextureDesc(p.shape[1], p.shape[0], fmt, int(srgb), int(wrap_u), int(wrap_v), int(linear), 0)
        self._check(self.lib.bpt_upload_texture(self.h, int(texture_id), C.byref(desc), _ptr(p)))

    def destroy_texture(self, texture_id):
        self._check(self.lib.bpt_destroy_texture(self.h, int(texture_id)))

    def texture_sample(self, texture_id, uv):
        uv = _f32(uv).reshape(-1, 2)
        out = np.empty((uv.shape[0], 4), np.float32)
        self._check(self.lib.bpt_texture_sample(self.h, int(texture_id), uv.shape[0], _ptr(uv), _ptr(out)))
        return out

    def set_mesh_emission(self, mesh_id, emission):
        e = None if emission is None else _f32(emission).reshape(-1, 3)
        self._check(self.lib.bpt_set_mesh_emission(self.h, int(mesh_id), _ptr(e), 0 if e is None else e.shape[0]))

    def remove_mesh(self, mesh_id):
        self._check(self.lib.bpt_remove_mesh(self.h, int(mesh_id)))

    def set_instances(self, instances):
        inst = np.ascontiguousarray(instances, dtype=INSTANCE_DTYPE)
        self._check(self.lib.bpt_set_instances(self.h, _ptr(inst), inst.shape[0]))

    def set_materials(self, materials):
        m = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        self._check(self.lib.bpt_set_materials(self.h, _ptr(m), m.shape[0]))

    def set_lights(self, lights):
        l = np.ascontiguousarray(lights, dtype=LIGHT_DTYPE)
        self._check(self.lib.bpt_set_lights(self.h, _ptr(l) if l.size else None, l.shape[0]))

    def set_environment(self, tint, texels=None, per_pixel_pdf=None, samples=None):
        t = _f32(tint, (3,))
        if texels is None:
            self._check(self.lib.bpt_set_environment(self.h, _ptr(t), None, 0, 0, None, 0, 0, None, 0))
            return
        tex = _f32(texels); assert tex.ndim == 3 and tex.shape[2] == 4
        pdf = _f32(per_pixel_pdf); assert pdf.ndim == 2
        s = np.ascontiguousarray(samples, dtype=LIGHT_SAMPLE_DTYPE)
        self._check(self.lib.bpt_set_environment(self.h, _ptr(t), _ptr(tex), tex.shape[1], tex.shape[0], _ptr(pdf), pdf.shape[1], pdf.shape[0], _ptr(s), s.shape[0]))

    def set_environment_cdfs(self, marginal_cdf=None, conditional_cdf=None):
        """Distribution2D CDFs of the environment map: marginal (H + 1), conditional (H, W + 1); None, None removes them."""
        if marginal_cdf is None:
            self._check(self.lib.bpt_set_environment_cdfs(self.h, None, None, 0, 0))
            return
        m = _f32(marginal_cdf); c = _f32(conditional_cdf)
        assert m.ndim == 1 and c.ndim == 2 and m.shape[0] == c.shape[0] + 1
        self._check(self.lib.bpt_set_environment_cdfs(self.h, _ptr(m), _ptr(c), c.shape[1] - 1, c.shape[0]))

    def set_environment_sampling(self, mode):
        """mode: "presampled" (the reference renderer's behaviour) or "cdf" (CDF inversion on the device)."""
        self._check(self.lib.bpt_set_environment_sampling(self.h, {"presampled": ENVIRONMENT_NEE_PRESAMPLED, "cdf": ENVIRONMENT_NEE_CDF}[mode]))

    def set_hit_sorting(self, from_iteration):
        """Sort the surface hits by (shading class, hit cell) before shading from this iteration of a sample on; < 0: never."""
        self._check(self.lib.bpt_set_hit_sorting(self.h, int(from_iteration)))

    def build_accel(self):
        self._check(self.lib.bpt_build_accel(self.h))

    def accel_info(self):
        t, n, ms = C.c_int64(), C.c_int64(), C.c_float()
        self._check(self.lib.bpt_accel_info(self.h, C.byref(t), C.byref(n), C.byref(ms)))
        kind, wide_nodes, levels = C.c_int32(), C.c_int64(), C.c_int32()
        self._check(self.lib.bpt_accel_hierarchy(self.h, C.byref(kind), C.byref(wide_nodes), C.byref(levels)))
        return {"triangles": t.value, "nodes": n.value, "build_ms": ms.value,
                "node_width": kind.value, "traversed_nodes": wide_nodes.value, "levels": levels.value}

    # ---- checkpoint / resume ----
    def read_accumulation(self):
        """The selected target's state: float64 [height, width, 4] (radiance sums, sample count)."""
        w, h = C.c_int32(), C.c_int32()
        self._check(self.lib.bpt_read_accumulation(self.h, None, C.byref(w), C.byref(h)))
        sums = np.empty((h.value, w.value, 4), np.float64)
        self._check(self.lib.bpt_read_accumulation(self.h, _ptr(sums), C.byref(w), C.byref(h)))
        return sums

    def write_accumulation(self, sums):
        sums = np.ascontiguousarray(sums, np.float64)
        assert sums.ndim == 3 and sums.shape[2] == 4
        self._check(self.lib.bpt_write_accumulation(self.h, sums.shape[1], sums.shape[0], _ptr(sums)))

    # ---- rendering ----
    def render(self, camera, width, height, first_sample, sample_count, max_bounces=4, nee_samples=3, pdf_scale=0.5, reset=False,
               russian_roulette_start=0):
        cam = camera if isinstance(camera, Camera) else make_camera(*camera)
        s = Settings(max_bounces, nee_samples, pdf_scale, russian_roulette_start)
        self._check(self.lib.bpt_render(self.h, C.byref(cam), C.byref(s), width, height, first_sample, sample_count, int(reset)))
        self._size = (width, height)

    AOV = {"depth": 3, "albedo": 4, "tint": 5, "roughness": 6, "shading_normal": 7, "primitive_id": 8}

    def render_aov(self, camera, kind, width, height, first_sample=0, sample_count=1, reset=True):
        cam = camera if isinstance(camera, Camera) else make_camera(*camera)
        self._check(self.lib.bpt_render_aov(self.h, C.byref(cam), self.AOV[kind], width, height, first_sample, sample_count, int(reset)))
        self._size = (width, height)

    def accumulation_device_ptr(self):
        return self.lib.bpt_accumulation_device_ptr(self.h)

    # ---- multi-GPU (NCCL bound at run time inside libbpt.so) ----
    def comm_unique_id(self):
        """ncclGetUniqueId as 128 plain bytes (rank 0 creates it, the host distributes it)."""
        buf = C.create_string_buffer(128)
        status = self.lib.bpt_comm_unique_id(buf)
        if status != 0:
            raise BptError(f"bpt_comm_unique_id failed with status {status}: libnccl.so.2 not loadable?")
        return buf.raw

    def comm_init(self, unique_id, rank_count, rank):
        assert len(unique_id) == 128
        self._check(self.lib.bpt_comm_init(self.h, C.create_string_buffer(unique_id, 128), int(rank_count), int(rank)))

    def comm_destroy(self):
        self._check(self.lib.bpt_comm_destroy(self.h))

    def comm_check(self):
        """Raises if NCCL has recorded an asynchronous error on the communicator."""
        self._check(self.lib.bpt_comm_check(self.h))

    def reduce_accumulation(self, root=0):
        """Sum of all ranks' selected accumulation targets into `root` (ncclReduce on the render stream); root < 0: all-reduce."""
        self._check(self.lib.bpt_reduce_accumulation(self.h, int(root)))

    def select_accumulation(self, slot):
        """One accumulation target per camera (Renderer.cpp:199-222): render / resolve act on the selected slot."""
        sizes = self.__dict__.setdefault("_target_sizes", {})
        sizes[getattr(self, "_slot", 0)] = getattr(self, "_size", None)
        self._check(self.lib.bpt_select_accumulation(self.h, int(slot)))
        self._slot = int(slot)
        self._size = sizes.get(self._slot)

    def release_accumulation(self, slot):
        self._check(self.lib.bpt_release_accumulation(self.h, int(slot)))
        if int(slot) == getattr(self, "_slot", 0):
            self._size = None
        else:
            self.__dict__.setdefault("_target_sizes", {}).pop(int(slot), None)

    def resolve_float4(self):
        w, h = self._size
        out = np.empty((h, w, 4), np.float32)
        self._check(self.lib.bpt_resolve_float4(self.h, _ptr(out)))
        return out

    def resolve_half4_async(self, out_host, slot):
        """out_host: a (H, W, 4) uint16 array (pinned for a truly asynchronous copy); complete after wait_frame(slot)."""
        self._check(self.lib.bpt_resolve_half4_async(self.h, out_host.ctypes.data, int(slot)))

    def wait_frame(self, slot):
        self._check(self.lib.bpt_wait_frame(self.h, int(slot)))

    def resolve_tonemapped(self, mode="filmic", exposure=1.0, filmic=FILMIC_ACES, rgba8=False):
        """Mean radiance -> exposure -> tonemapping operator: linear float4, or sRGB-encoded RGBA8 when rgba8 is set."""
        w, h = self._size
        s = tonemap_settings(mode, exposure, filmic)
        out = np.empty((h, w, 4), np.uint8 if rgba8 else np.float32)
        self._check(self.lib.bpt_resolve_tonemapped(self.h, C.byref(s), _ptr(out), 1 if rgba8 else 0))
        return out

    def tonemap_colors(self, rgb, mode="filmic", exposure=1.0, filmic=FILMIC_ACES):
        rgb = _f32(rgb).reshape(-1, 3)
        out = np.empty_like(rgb)
        s = tonemap_settings(mode, exposure, filmic)
        self._check(self.lib.bpt_tonemap_colors(self.h, C.byref(s), rgb.shape[0], _ptr(rgb), _ptr(out)))
        return out

    def resolve_half4(self):
        w, h = self._size
        out = np.empty((h, w, 4), np.uint16)
        self._check(self.lib.bpt_resolve_half4(self.h, _ptr(out), 0))
        return out.view(np.float16)

    def counters(self, reset=False):
        c = Counters()
        self._check(self.lib.bpt_get_counters(self.h, C.byref(c), int(reset)))
        return {k: getattr(c, k) for k, _ in Counters._fields_}

    # ---- batched unit entry points (host arrays) ----
    def bsdf_eval_sample_pdf(self, kind, wo, wi, tint, rms, u, coat=None):
        wo, wi, tint, rms, u = (_f32(a).reshape(-1, 3) for a in (wo, wi, tint, rms, u))
        n = wo.shape[0]
        coat = None if coat is None else _f32(coat).reshape(-1, 2)
        ef, sf, sd = (np.empty((n, 3), np.float32) for _ in range(3))
        ep, sp = (np.empty(n, np.float32) for _ in range(2))
        self._check(self.lib.bpt_bsdf_eval_sample_pdf(self.h, kind, n, _ptr(wo), _ptr(wi), _ptr(tint), _ptr(rms), _ptr(coat), _ptr(u),
                                                      _ptr(ef), _ptr(ep), _ptr(sf), _ptr(sp), _ptr(sd), 0))
        return {"eval_f": ef, "eval_pdf": ep, "sample_f": sf, "sample_pdf": sp, "sample_dir": sd}

    def bsdf_eval_sample_pdf_device(self, kind, n, ptrs):
        """ptrs: dict of raw device pointers (ints) with the argument names of bpt_bsdf_eval_sample_pdf."""
        order = ["wo", "wi", "tint", "rms", "coat", "u", "eval_f", "eval_pdf", "sample_f", "sample_pdf", "sample_dir"]
        args = [C.c_void_p(ptrs.get(k) or None) for k in order]
        self._check(self.lib.bpt_bsdf_eval_sample_pdf(self.h, kind, n, *args, 1))

    def default_shading_regularized(self, materials, max_pdf_hint, wo, wi, u, tint_roughness_scale=None):
        m = np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE)
        n = m.shape[0]
        hint = _f32(max_pdf_hint, (n,)); wo, wi, u = (_f32(a).reshape(n, 3) for a in (wo, wi, u))
        sc = None if tint_roughness_scale is None else _f32(tint_roughness_scale).reshape(n, 4)
        ef, sf, sd = (np.empty((n, 3), np.float32) for _ in range(3))
        ep, sp = (np.empty(n, np.float32) for _ in range(2))
        self._check(self.lib.bpt_default_shading_regularized(self.h, n, _ptr(m), _ptr(sc), _ptr(hint), _ptr(wo), _ptr(wi), _ptr(u),
                                                             _ptr(ef), _ptr(ep), _ptr(sf), _ptr(sp), _ptr(sd)))
        return {"eval_f": ef, "eval_pdf": ep, "sample_f": sf, "sample_pdf": sp, "sample_dir": sd}

    def light_sample_pdf_evaluate(self, lights, position, u2, query_direction):
        l = np.ascontiguousarray(lights, dtype=LIGHT_DTYPE).reshape(-1)
        position, query_direction = _f32(position).reshape(-1, 3), _f32(query_direction).reshape(-1, 3)
        u2 = _f32(u2).reshape(-1, 2)
        n = position.shape[0]
        stride = 0 if l.shape[0] == 1 and n != 1 else 1
        assert stride == 0 or l.shape[0] == n
        samples = np.empty(n, LIGHT_SAMPLE_DTYPE); pdf = np.empty(n, np.float32); rad = np.empty((n, 3), np.float32)
        self._check(self.lib.bpt_light_sample_pdf_evaluate(self.h, n, _ptr(l), stride, _ptr(position), _ptr(u2), _ptr(query_direction),
                                                           _ptr(samples), _ptr(pdf), _ptr(rad)))
        return samples, pdf, rad

    def compare_images(self, reference, target, mssim_support=0, diff_images=False):
        """ImageOperations::Compare rms / ssim / mssim (mssim only with a support > 0) of two (H, W, 3 or 4) float images."""
        def rgba(image):
            image = np.asarray(image, np.float32)
            if image.shape[-1] == 3:
                image = np.concatenate([image, np.ones(image.shape[:-1] + (1,), np.float32)], axis=-1)
            return np.ascontiguousarray(image)
        a, b = rgba(reference), rgba(target)
        assert a.shape == b.shape and a.ndim == 3
        h, w = a.shape[:2]
        rms, ssim, mssim = C.c_float(), C.c_float(), C.c_float()
        rms_diff = np.empty_like(a) if diff_images else None
        mssim_diff = np.empty_like(a) if diff_images and mssim_support > 0 else None
        self._check(self.lib.bpt_compare_images(self.h, w, h, _ptr(a), _ptr(b), int(mssim_support), C.byref(rms), C.byref(ssim),
                                                C.byref(mssim) if mssim_support > 0 else None, _ptr(rms_diff), _ptr(mssim_diff)))
        out = {"rms": rms.value, "ssim": ssim.value}
        if mssim_support > 0:
            out["mssim"] = mssim.value
        if diff_images:
            out["rms_diff"] = rms_diff; out["mssim_diff"] = mssim_diff
        return out

    def sort_pairs(self, keys, values, begin_bit=0, end_bit=64):
        """Stable radix sort of (uint64 key, uint32 value) pairs by key bits [begin_bit, end_bit): the sort of the BVH build."""
        k = np.array(keys, dtype=np.uint64).reshape(-1); v = np.array(values, dtype=np.uint32).reshape(-1)
        assert k.shape == v.shape
        self._check(self.lib.bpt_sort_pairs(self.h, k.shape[0], _ptr(k), _ptr(v), int(begin_bit), int(end_bit)))
        return k, v

    def exclusive_scan(self, values):
        v = np.ascontiguousarray(values, dtype=np.uint32).reshape(-1)
        out = np.empty_like(v); total = np.zeros(1, np.uint32)
        self._check(self.lib.bpt_exclusive_scan(self.h, v.shape[0], _ptr(v), _ptr(out), _ptr(total)))
        return out, int(total[0])

    def rng_sample4(self, accumulation, pixel_hash, dimension):
        a, h, d = (np.ascontiguousarray(x, dtype=np.uint32).reshape(-1) for x in (accumulation, pixel_hash, dimension))
        n = a.shape[0]
        ui = np.empty((n, 4), np.uint32); f = np.empty((n, 4), np.float32)
        self._check(self.lib.bpt_rng_sample4(self.h, n, _ptr(a), _ptr(h), _ptr(d), _ptr(ui), _ptr(f)))
        return ui, f

    def intersect(self, origins, directions, tmin=None, tmax=None, want_occluded=True):
        o, d = _f32(origins).reshape(-1, 3), _f32(directions).reshape(-1, 3)
        n = o.shape[0]
        tmin = np.zeros(n, np.float32) if tmin is None else _f32(np.broadcast_to(tmin, (n,)))
        tmax = np.full(n, 1e30, np.float32) if tmax is None else _f32(np.broadcast_to(tmax, (n,)))
        prim = np.empty(n, np.int32); t = np.empty(n, np.float32); uv = np.empty((n, 2), np.float32)
        occ = np.empty(n, np.uint8) if want_occluded else None
        self._check(self.lib.bpt_intersect(self.h, n, _ptr(o), _ptr(d), _ptr(tmin), _ptr(tmax), _ptr(prim), _ptr(t), _ptr(uv), _ptr(occ)))
        return prim, t, uv, occ


def make_camera(view_to_world_rotation, inverse_projection, inverse_view_projection):
    cam = Camera()
    cam.view_to_world_rotation[:] = [float(x) for x in np.asarray(view_to_world_rotation, np.float32).reshape(9)]
    cam.inverse_projection[:] = [float(x) for x in np.asarray(inverse_projection, np.float32).reshape(16)]
    cam.inverse_view_projection[:] = [float(x) for x in np.asarray(inverse_view_projection, np.float32).reshape(16)]
    return cam
