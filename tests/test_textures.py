"""Textured materials (Types.h:388-414, Renderer.cpp:650-751): the CUDA texture objects of the product against the oracle's
restatement of the texture unit's documented arithmetic, texel by texel and through the integrator."""
import numpy as np
import pytest

from tests import oracle_lib
from tests.test_render_parity import REL_MSE_BOUND, rel_mse, render_both
from bifrost3d_b200 import capi, scenes

needs_oracle = pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")


def checker(size, cells, a, b, dtype=np.uint8):
    y, x = np.mgrid[0:size, 0:size]
    on = ((x * cells // size) + (y * cells // size)) % 2 == 0
    img = np.where(on[..., None], np.array(a)[None, None, :], np.array(b)[None, None, :])
    return img.astype(dtype)


def test_cpu_sampler_known_answers():
    """The oracle's sampler on a 2x2 texture: texel centres return the texel, the middle returns the mean, repeat wraps."""
    if not oracle_lib.available():
        pytest.skip("oracle/_ref not built")
    tex = np.array([[[0, 0, 0, 255], [255, 0, 0, 255]], [[0, 255, 0, 255], [255, 255, 255, 255]]], np.uint8)
    scene = {"meshes": {}, "materials": np.array([scenes.material((0, 0, 0), 0)], capi.MATERIAL_DTYPE),
             "instances": np.zeros(0, capi.INSTANCE_DTYPE), "lights": np.zeros(0, capi.LIGHT_DTYPE),
             "textures": {1: {"pixels": tex, "linear": True, "wrap_u": capi.WRAP_REPEAT, "wrap_v": capi.WRAP_CLAMP},
                          2: {"pixels": tex, "linear": False, "srgb": True}}}
    sc = oracle_lib.OracleScene(scene)
    uv = np.array([[0.25, 0.25], [0.75, 0.25], [0.25, 0.75], [0.5, 0.5], [1.25, 0.25], [0.0, 0.25], [0.25, -3.0]], np.float32)
    got = sc.texture_sample(1, uv)
    assert np.allclose(got[0], [0, 0, 0, 1]) and np.allclose(got[1], [1, 0, 0, 1]) and np.allclose(got[2], [0, 1, 0, 1])
    assert np.allclose(got[3], [0.5, 0.5, 0.25, 1])
    assert np.allclose(got[4], got[0])                      # repeat in u
    assert np.allclose(got[5], [0.5, 0, 0, 1])              # u = 0 blends the last and the first column
    assert np.allclose(got[6], got[0])                      # clamp in v
    near = sc.texture_sample(2, np.array([[0.74, 0.2], [0.76, 0.7]], np.float32))
    assert np.allclose(near, [[1, 0, 0, 1], [1, 1, 1, 1]])
    sc.close()


@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("name", ["rgba8_linear_repeat", "rgba8_srgb_linear_clamp", "alpha8_nearest", "alpha8_linear", "rgb24_linear", "float4_linear", "float3_nearest"])
def test_texture_objects_match_the_sampler_restatement(bpt, name):
    rng = np.random.default_rng(hash(name) % 1000)
    w, h = 37, 23
    if name.startswith("rgba8"):
        pixels = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    elif name.startswith("alpha8"):
        pixels = rng.integers(0, 256, (h, w, 1), dtype=np.uint8)
    elif name.startswith("rgb24"):
        pixels = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif name.startswith("float4"):
        pixels = rng.random((h, w, 4), dtype=np.float32) * 4
    else:
        pixels = rng.random((h, w, 3), dtype=np.float32) * 4
    tex = {"pixels": pixels, "srgb": "srgb" in name, "linear": "linear" in name,
           "wrap_u": capi.WRAP_CLAMP if "clamp" in name else capi.WRAP_REPEAT, "wrap_v": capi.WRAP_CLAMP if "clamp" in name else capi.WRAP_REPEAT}
    scene = {"meshes": {}, "materials": np.array([scenes.material((0, 0, 0), 0)], capi.MATERIAL_DTYPE),
             "instances": np.zeros(0, capi.INSTANCE_DTYPE), "lights": np.zeros(0, capi.LIGHT_DTYPE), "textures": {3: tex}}
    sc = oracle_lib.OracleScene(scene)
    bpt.upload_texture(3, tex["pixels"], tex["srgb"], tex["wrap_u"], tex["wrap_v"], tex["linear"])
    n = 1 << 16
    uv = rng.uniform(-1.5, 2.5, (n, 2)).astype(np.float32)
    uv[: n // 2] = rng.random((n // 2, 2), dtype=np.float32)
    got, want = bpt.texture_sample(3, uv), sc.texture_sample(3, uv)
    sc.close()
    channels = [0] if pixels.shape[2] == 1 else [0, 1, 2, 3]
    scale = 4.0 if pixels.dtype == np.float32 else 1.0
    err = np.abs(got[:, channels] - want[:, channels]).max(axis=1) / scale
    print(f"{name}: max abs err {err.max():.3e}, mean {err.mean():.3e}, above 1/256: {np.mean(err > 1 / 256):.2e}")
    if tex["linear"]:
        # the hardware holds the interpolation weights in 1.8 fixed point: at most ~2/256 of the texel range per fetch
        # (and decodes sRGB texels through a table of limited precision)
        assert err.max() <= 3.0 / 256 and err.mean() < 2.5e-3
    else:
        assert np.mean(err > 1e-6) < 1e-3  # a texcoord on a texel boundary may round to either side


def textured_cornell():
    scene = scenes.cornell_box(sphere_quads=(24, 12))
    rng = np.random.default_rng(9)
    tint = checker(64, 8, (230, 60, 40, 255), (250, 250, 250, 120))                   # tint rgb (sRGB) + roughness in alpha
    roughness = rng.integers(30, 256, (16, 16, 1), dtype=np.uint8)
    metallic = checker(32, 4, (255,), (0,))
    coverage = checker(32, 6, (255,), (40,))
    scene["textures"] = {1: {"pixels": tint, "srgb": True, "linear": True},
                         2: {"pixels": roughness, "linear": True, "wrap_u": capi.WRAP_CLAMP, "wrap_v": capi.WRAP_CLAMP},
                         3: {"pixels": metallic, "linear": False},
                         4: {"pixels": coverage, "linear": True}}
    mats = scene["materials"].copy()
    mats[1]["tint_roughness_texture_id"] = 1                                           # floor, roof, back wall
    mats[4]["roughness_texture_id"] = 2; mats[4]["metallic_texture_id"] = 3            # left sphere
    mats[3]["coverage_texture_id"] = 4; mats[3]["coverage"] = 0.9                      # right wall: partial coverage
    mats[2]["coverage_texture_id"] = 4; mats[2]["coverage"] = 0.5; mats[2]["flags"] = 3  # left wall: cutout at threshold 0.5
    scene["materials"] = mats
    scene["environment"] = {"tint": (0.3, 0.35, 0.5)}                                  # light leaks in through the cut-outs
    return scene


@pytest.mark.gpu
@needs_oracle
def test_textured_materials_match_oracle(tracer):
    """Tint/roughness (sRGB RGBA8), roughness + metallic (Alpha8), partial coverage and cutout textures: closest hit,
    stochastic coverage and the shadow any-hit all read them (MonteCarlo.cu:152-164, 278-285)."""
    bpt = tracer
    scene = textured_cornell()
    gpu, cpu, counters, oc = render_both(bpt, scene, 96, 96, 6)
    assert np.isfinite(gpu).all() and cpu.mean() > 0.01
    e = rel_mse(gpu, cpu)
    diff = np.abs(gpu[..., :3] - cpu).max(axis=-1)
    close = diff <= 2e-3 * (1 + np.abs(cpu).max(axis=-1))
    print(f"relMSE {e:.3e}; pixels within 2e-3: {close.mean():.4f}; rays gpu {counters['extend_rays']}+{counters['shadow_rays']} cpu {oc[0]}+{oc[1]}")
    assert e <= REL_MSE_BOUND
    assert close.mean() > 0.95
    assert abs(int(counters["extend_rays"]) - int(oc[0])) <= 0.01 * int(oc[0])
    # and the textures matter: the same scene without them renders a different image
    plain = scenes.cornell_box(sphere_quads=(24, 12)); plain["environment"] = scene["environment"]
    scenes.upload(bpt, plain)
    bpt.render(plain["camera"], 96, 96, 0, 6, reset=True)
    assert rel_mse(bpt.resolve_float4(), cpu) > 1e-2


@pytest.mark.gpu
def test_texture_aovs_and_errors(bpt):
    scene = textured_cornell()
    scenes.upload(bpt, scene)
    bpt.render_aov(scene["camera"], "tint", 64, 64)
    tint = bpt.resolve_float4()[..., :3]
    floor = tint[56:, 8:56]  # the textured floor shows both checker colours
    assert floor[..., 1].min() < 0.2 and floor[..., 1].max() > 0.8
    bpt.render_aov(scene["camera"], "roughness", 64, 64)
    rough = bpt.resolve_float4()[..., 0]
    assert np.isfinite(rough).all() and rough.max() <= 1.0 + 1e-6
    with pytest.raises(capi.BptError, match="still references"):
        bpt.destroy_texture(1)
    bad = scene["materials"].copy(); bad[1]["tint_roughness_texture_id"] = 77
    with pytest.raises(capi.BptError, match="not uploaded"):
        bpt.set_materials(bad)
    bad = scene["materials"].copy(); bad[1]["tint_roughness_texture_id"] = 2  # one channel where four are needed
    with pytest.raises(capi.BptError, match="four channels"):
        bpt.set_materials(bad)
    with pytest.raises(capi.BptError, match="unsupported pixel format"):
        desc = capi.TextureDesc(2, 2, 2, 0, 0, 0, 1, 0)  # Intensity8: the reference does not upload it either
        import ctypes
        bpt._check(bpt.lib.bpt_upload_texture(bpt.h, 9, ctypes.byref(desc), np.zeros(4, np.uint8).ctypes.data_as(ctypes.c_void_p)))


@pytest.mark.gpu
def test_first_textured_material_asks_for_a_rebuild(bpt):
    """An acceleration structure built for untextured materials carries no per-primitive texcoords: the first textured
    material invalidates it (bpt_set_materials), and after the rebuild the scene renders like one uploaded from scratch."""
    scene = textured_cornell()
    plain = dict(scene); plain["materials"] = scenes.cornell_box(sphere_quads=(24, 12))["materials"]
    scenes.upload(bpt, plain)                      # textures are uploaded, but no material references them yet
    bpt.render(scene["camera"], 48, 48, 0, 2, reset=True)
    bpt.set_materials(scene["materials"])
    with pytest.raises(capi.BptError, match="bpt_build_accel"):
        bpt.render(scene["camera"], 48, 48, 0, 1)
    bpt.build_accel()
    bpt.render(scene["camera"], 48, 48, 0, 2, reset=True)
    edited = bpt.resolve_float4()
    fresh = capi.Bpt(0)
    scenes.upload(fresh, scene)
    fresh.render(scene["camera"], 48, 48, 0, 2, reset=True)
    assert np.array_equal(edited, fresh.resolve_float4())
    fresh.close()


@pytest.mark.gpu
def test_material_edits_keep_an_acceleration_structure_built_without_texcoords(bpt):
    """Textured materials over meshes that carry no texcoords: the build has nothing to add for them, so a later material
    edit must not invalidate the acceleration structure (it did before round 2: every edit forced a full rebuild)."""
    scene = textured_cornell()
    scene["meshes"] = {k: {name: value for name, value in m.items() if name != "texcoords"} for k, m in scene["meshes"].items()}
    scenes.upload(bpt, scene)
    bpt.render(scene["camera"], 32, 32, 0, 1, reset=True)
    mats = scene["materials"].copy()
    mats["roughness"] = np.clip(mats["roughness"] * 0.5, 0, 1)
    bpt.set_materials(mats)
    bpt.render(scene["camera"], 32, 32, 0, 1, reset=True)  # no bpt_build_accel in between
    assert np.isfinite(bpt.resolve_float4()).all()
