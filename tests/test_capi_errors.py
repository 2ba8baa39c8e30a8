"""Error behaviour of the C ABI: status codes + bpt_last_error instead of exceptions or crashes (the reference's
Renderer never throws either: Renderer.cpp:1365-1378,1255-1256)."""
import numpy as np
import pytest

import bifrost3d_b200 as b
from bifrost3d_b200 import scenes, capi

pytestmark = pytest.mark.gpu


def test_render_before_build_is_refused():
    ctx = b.Bpt(0)
    with pytest.raises(b.BptError, match="bpt_build_accel"):
        ctx.render(scenes.cornell_box()["camera"], 8, 8, 0, 1)
    with pytest.raises(b.BptError, match="bpt_build_accel"):
        ctx.intersect(np.zeros((1, 3), np.float32), np.array([[0, 0, 1]], np.float32))
    ctx.close()


def test_bad_inputs_are_rejected_loudly():
    ctx = b.Bpt(0)
    tri = {"indices": np.array([[0, 1, 5]], np.uint32), "positions": np.zeros((3, 3), np.float32)}
    with pytest.raises(b.BptError, match="out of range"):
        ctx.upload_mesh(0, tri["indices"], tri["positions"])
    with pytest.raises(b.BptError, match="unknown mesh"):
        ctx.set_instances(np.array([scenes._instance(7, 0, scenes.affine())], capi.INSTANCE_DTYPE))
    unknown = scenes.material((1, 1, 1), 0.1); unknown["shading_model"] = 3
    with pytest.raises(b.BptError, match="unknown shading model"):
        ctx.set_materials(np.array([unknown], capi.MATERIAL_DTYPE))
    textured = scenes.material((1, 1, 1), 0.1); textured["coverage_texture_id"] = 3
    with pytest.raises(b.BptError, match="not uploaded"):
        ctx.set_materials(np.array([textured], capi.MATERIAL_DTYPE))
    with pytest.raises(b.BptError):
        ctx.set_lights(np.zeros(1, capi.LIGHT_DTYPE))  # type 0 = None
    ctx.close()


def test_tables_are_required_for_default_shading():
    ctx = b.Bpt(0, tables=False)
    z = np.zeros((4, 3), np.float32)
    with pytest.raises(b.BptError, match="bpt_set_tables"):
        ctx.bsdf_eval_sample_pdf(0, z, z, z, z, z)
    out = ctx.bsdf_eval_sample_pdf(3, z + np.float32([0, 0, 1]), z + np.float32([0, 0, 1]), z + 0.5, z + 0.5, z + 0.5)  # Burley needs no tables
    assert np.isfinite(out["eval_f"]).all()
    ctx.close()


def test_dielectric_tables_are_required_for_transmissive_materials():
    ctx = b.Bpt(0, tables=False)
    ctx.set_tables(*capi.default_tables())
    z = np.zeros((4, 3), np.float32)
    with pytest.raises(b.BptError, match="bpt_set_dielectric_tables"):
        ctx.bsdf_eval_sample_pdf(4, z, z, z, z, z)
    scene = scenes.cornell_box(sphere_quads=(8, 4))
    scene["materials"][4]["shading_model"] = 2
    scenes.upload(ctx, scene)
    with pytest.raises(b.BptError, match="bpt_set_dielectric_tables"):
        ctx.render(scene["camera"], 8, 8, 0, 1)
    ctx.set_dielectric_tables(*capi.default_dielectric_tables())
    ctx.render(scene["camera"], 8, 8, 0, 1)
    assert np.isfinite(ctx.resolve_float4()).all()
    ctx.close()


def test_two_contexts_are_independent():
    a, c = b.Bpt(0), b.Bpt(0)
    s1 = scenes.cornell_box(sphere_quads=(8, 4))
    scenes.upload(a, s1)
    s2 = scenes.cornell_box(sphere_quads=(8, 4)); s2["lights"] = np.zeros(0, capi.LIGHT_DTYPE); s2["environment"] = {"tint": (1, 1, 1)}
    scenes.upload(c, s2)
    a.render(s1["camera"], 32, 32, 0, 2, reset=True); c.render(s2["camera"], 32, 32, 0, 2, reset=True)
    ia, ic = a.resolve_float4(), c.resolve_float4()
    assert not np.array_equal(ia, ic) and np.isfinite(ia).all() and np.isfinite(ic).all()
    a.close(); c.close()
