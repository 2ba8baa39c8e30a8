"""Scene edits through the C ABI (Renderer::handle_updates, Renderer.cpp:578-1205): meshes stay on the device, a material edit
needs no rebuild, a transform edit rebuilds from the resident meshes; every edited scene renders bit for bit like the same
scene uploaded from scratch."""
import numpy as np
import pytest

import bifrost3d_b200 as b
from bifrost3d_b200 import capi, scenes

pytestmark = pytest.mark.gpu


def fresh_render(scene, size=(48, 40), spp=3):
    ctx = b.Bpt(0)
    scenes.upload(ctx, scene)
    ctx.render(scene["camera"], size[0], size[1], 0, spp, reset=True)
    img = ctx.resolve_float4()
    ctx.close()
    return img


def test_material_edit_keeps_the_acceleration_structure(tracer):
    bpt = tracer
    scene = scenes.cornell_box(sphere_quads=(16, 8))
    scenes.upload(bpt, scene)
    bpt.render(scene["camera"], 48, 40, 0, 3, reset=True)
    before = bpt.resolve_float4()
    build_ms = bpt.accel_info()["build_ms"]
    mats = scene["materials"].copy()
    mats[2]["tint"] = (0.1, 0.2, 0.9); mats[4]["roughness"] = 0.9; mats[5]["coverage"] = 0.5
    bpt.set_materials(mats)                       # no bpt_build_accel
    assert bpt.accel_info()["build_ms"] == build_ms  # still the same structure
    bpt.render(scene["camera"], 48, 40, 0, 3, reset=True)
    after = bpt.resolve_float4()
    assert not np.array_equal(before, after)
    edited = dict(scene); edited["materials"] = mats
    assert np.array_equal(after, fresh_render(edited))


def test_transform_edit_rebuilds_from_resident_meshes(tracer):
    bpt = tracer
    scene = scenes.cornell_box(sphere_quads=(16, 8))
    scenes.upload(bpt, scene)
    inst = scene["instances"].copy()
    inst[5]["to_world"] = scenes.affine((-0.1, -0.2, 0.0), scenes.quat_from_angle_axis(0.5, (0, 1, 0)), 0.4)
    inst = inst[:-1]                               # and one model disappears
    bpt.set_instances(inst)
    with pytest.raises(capi.BptError, match="bpt_build_accel"):
        bpt.render(scene["camera"], 48, 40, 0, 1)
    bpt.build_accel()                              # no mesh upload: the meshes are resident
    bpt.render(scene["camera"], 48, 40, 0, 3, reset=True)
    edited = dict(scene); edited["instances"] = inst
    assert np.array_equal(bpt.resolve_float4(), fresh_render(edited))


def test_mesh_lifetime(bpt):
    scene = scenes.cornell_box(sphere_quads=(8, 4))
    scenes.upload(bpt, scene)
    with pytest.raises(capi.BptError, match="still references"):
        bpt.remove_mesh(1)
    inst = scene["instances"][:5].copy()           # walls only: mesh 1 (the sphere) is unreferenced now
    bpt.set_instances(inst); bpt.build_accel()
    bpt.remove_mesh(1)
    with pytest.raises(capi.BptError, match="unknown mesh"):
        bpt.remove_mesh(1)
    bpt.render(scene["camera"], 32, 32, 0, 2, reset=True)
    assert np.isfinite(bpt.resolve_float4()).all()
    mats = scene["materials"][:2].copy()           # the walls' red/green materials go away while still referenced
    bpt.set_materials(mats)
    with pytest.raises(capi.BptError, match="bpt_build_accel"):
        bpt.render(scene["camera"], 32, 32, 0, 1)
    with pytest.raises(capi.BptError, match="material that was not uploaded"):
        bpt.build_accel()
