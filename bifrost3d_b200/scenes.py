"""Procedural scenes of the shapes BASELINE.json names, built with numpy.

Scene recipes mirror the reference's own content where it has any:
  * plane / revolved_sphere tessellation: core/Bifrost/Bifrost/Assets/MeshCreation.cpp:30-69,336-386
  * Cornell box layout, materials and light: apps/SimpleViewer/Scenes/CornellBox.h:23-122 (walls, light),
    apps/SmallPT/smallpt.h:47-57 (the two 16.5-radius spheres)
  * perspective projection: core/Bifrost/Bifrost/Scene/Camera.cpp:237-267
A scene is a plain dict of numpy arrays that both the product (capi.Bpt) and the oracle consume.
"""
import numpy as np

from . import capi

IRON_TINT = (0.560, 0.570, 0.580)    # core/Bifrost/Bifrost/Assets/Material.h metal tints
COPPER_TINT = (0.955, 0.637, 0.538)


# ---- meshes -------------------------------------------------------------------------------------

def plane(quads_per_edge=1):
    size = quads_per_edge + 1
    tc = np.arange(size, dtype=np.float32) * np.float32(1.0 / quads_per_edge)
    x, z = np.meshgrid(tc, tc)  # z-major
    positions = np.stack([x - np.float32(0.5), np.zeros_like(x), z - np.float32(0.5)], axis=-1).reshape(-1, 3).astype(np.float32)
    normals = np.tile(np.array([0, 1, 0], np.float32), (size * size, 1))
    zz, xx = np.meshgrid(np.arange(quads_per_edge), np.arange(quads_per_edge), indexing="ij")
    base = (xx + zz * size).reshape(-1)
    tris = np.stack([np.stack([base, base + size, base + 1], axis=1), np.stack([base + 1, base + size, base + size + 1], axis=1)], axis=1)
    texcoords = np.stack([x, z], axis=-1).reshape(-1, 2).astype(np.float32)  # MeshCreation::plane: texcoord = (x, z) in [0, 1]
    return {"indices": tris.reshape(-1, 3).astype(np.uint32), "positions": positions, "normals": normals, "texcoords": texcoords}


def revolved_sphere(longitude_quads=100, latitude_quads=50):
    """Radius 0.5 latitude/longitude sphere with 2 * (lat * lon - lon) triangles."""
    lat_size, lon_size = latitude_quads + 1, longitude_quads + 1
    ty = (np.arange(lat_size, dtype=np.float32) * np.float32(1.0 / latitude_quads))
    tx = (np.arange(lon_size, dtype=np.float32) * np.float32(1.0 / longitude_quads))
    theta = (ty * np.float32(np.pi))[:, None]
    phi = (tx * np.float32(2.0) * np.float32(np.pi))[None, :]
    sin_theta = np.sin(theta).astype(np.float32)
    d = np.stack([-sin_theta * np.sin(phi), np.cos(theta) * np.ones_like(phi), sin_theta * np.cos(phi)], axis=-1).astype(np.float32)
    positions = (d * np.float32(0.5)).reshape(-1, 3)
    normals = positions / np.linalg.norm(positions, axis=1, keepdims=True).astype(np.float32)
    positions[:lon_size] = (0, 0.5, 0)
    positions[(lat_size - 1) * lon_size:] = (0, -0.5, 0)
    tris = []
    for y in range(latitude_quads):
        base = np.arange(longitude_quads) + y * lon_size
        if y != 0:
            tris.append(np.stack([base, base + 1, base + lon_size], axis=1))
        if y != latitude_quads - 1:
            tris.append(np.stack([base + 1, base + lon_size + 1, base + lon_size], axis=1))
    # the reference interleaves the two triangles of each quad; the order only permutes primitive ids
    texcoords = np.stack(np.broadcast_arrays(tx[None, :], ty[:, None]), axis=-1).reshape(-1, 2).astype(np.float32)  # (longitude, latitude) in [0, 1]
    return {"indices": np.concatenate(tris).astype(np.uint32), "positions": positions.astype(np.float32), "normals": normals.astype(np.float32),
            "texcoords": texcoords}


def displaced_grid(quads_per_edge, seed, amplitude=0.08, octaves=4):
    """plane(N) displaced along y by seeded value noise; 2 * N^2 triangles."""
    m = plane(quads_per_edge)
    p = m["positions"]
    h = value_noise(p[:, 0] + 0.5, p[:, 2] + 0.5, seed, octaves) * np.float32(amplitude)
    p[:, 1] = h
    size = quads_per_edge + 1
    hh = h.reshape(size, size)
    dz, dx = np.gradient(hh, np.float32(1.0 / quads_per_edge))
    n = np.stack([-dx, np.ones_like(dx), -dz], axis=-1).reshape(-1, 3)
    m["normals"] = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    return m


def value_noise(x, y, seed, octaves=4):
    rng = np.random.default_rng(seed)
    out = np.zeros_like(x, dtype=np.float32)
    amp, freq = 1.0, 4.0
    for _ in range(octaves):
        n = int(freq) + 2
        lattice = rng.random((n, n), dtype=np.float32)
        fx, fy = x * freq, y * freq
        ix, iy = np.floor(fx).astype(int), np.floor(fy).astype(int)
        tx, ty = (fx - ix).astype(np.float32), (fy - iy).astype(np.float32)
        tx, ty = tx * tx * (3 - 2 * tx), ty * ty * (3 - 2 * ty)
        ix, iy = np.clip(ix, 0, n - 2), np.clip(iy, 0, n - 2)
        a = lattice[iy, ix] * (1 - tx) + lattice[iy, ix + 1] * tx
        b = lattice[iy + 1, ix] * (1 - tx) + lattice[iy + 1, ix + 1] * tx
        out += np.float32(amp) * (a * (1 - ty) + b * ty - np.float32(0.5))
        amp *= 0.5; freq *= 2.0
    return out.astype(np.float32)


# ---- transforms / camera ------------------------------------------------------------------------

def quat_from_angle_axis(angle, axis):
    axis = np.asarray(axis, np.float64); axis = axis / np.linalg.norm(axis)
    s = np.sin(angle * 0.5)
    return np.array([axis[0] * s, axis[1] * s, axis[2] * s, np.cos(angle * 0.5)])


def rotation_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def affine(translation=(0, 0, 0), rotation=None, scale=1.0):
    """Row-major 3x4 object->world matrix of a Bifrost Transform {rotation, translation, uniform scale}."""
    r = np.eye(3) if rotation is None else rotation_matrix(rotation)
    m = np.zeros((3, 4))
    m[:, :3] = r * scale
    m[:, 3] = translation
    return m.astype(np.float32).reshape(12)


def perspective_camera(position, rotation=None, fov=np.pi / 4, aspect=1.0, near=0.1, far=100.0):
    """(view_to_world_rotation 3x3, inverse_projection 4x4, inverse_view_projection 4x4), row-major float32.
    +Z is forward (Camera.cpp:237-267)."""
    f = 1.0 / np.tan(fov * 0.5)
    a = (far + near) / (near - far)
    b = (2.0 * far * near) / (near - far)
    proj = np.zeros((4, 4))
    proj[0, 0] = f / aspect; proj[1, 1] = f; proj[2, 2] = -a; proj[2, 3] = b; proj[3, 2] = 1.0
    inv_proj = np.zeros((4, 4))
    inv_proj[0, 0] = 1.0 / proj[0, 0]; inv_proj[1, 1] = 1.0 / proj[1, 1]; inv_proj[2, 3] = 1.0
    inv_proj[3, 2] = 1.0 / proj[2, 3]; inv_proj[3, 3] = -proj[2, 2] / proj[2, 3]
    r = np.eye(3) if rotation is None else rotation_matrix(rotation)
    inv_view = np.eye(4); inv_view[:3, :3] = r; inv_view[:3, 3] = position
    inv_view_proj = inv_view @ inv_proj
    return r.astype(np.float32), inv_proj.astype(np.float32), inv_view_proj.astype(np.float32)


# ---- materials / lights -------------------------------------------------------------------------

def material(tint, roughness, specularity=0.04, metallic=0.0, coat=0.0, coat_roughness=0.0, thin_walled=False, emission=(0, 0, 0), coverage=1.0):
    m = np.zeros((), capi.MATERIAL_DTYPE)
    m["flags"] = 1 if thin_walled else 0
    m["tint"] = tint; m["roughness"] = roughness; m["specularity"] = specularity; m["metallic"] = metallic
    m["coverage"] = coverage; m["emission"] = emission
    m["coat"] = int(min(max(coat, 0.0), 1.0) * 65535.0 + 0.5); m["coat_roughness"] = int(min(max(coat_roughness, 0.0), 1.0) * 65535.0 + 0.5)
    return m


def sphere_light(power, position, radius):
    l = np.zeros((), capi.LIGHT_DTYPE)
    l["data"][0:3] = power; l["data"][3:6] = position; l["data"][6] = radius
    l["flags"] = capi.LIGHT_SPHERE
    return l


def spot_light(power, position, radius, direction, cos_angle):
    l = np.zeros((), capi.LIGHT_DTYPE)
    d = np.asarray(direction, np.float64); d = d / np.linalg.norm(d)
    l["data"][0:3] = power; l["data"][3:6] = position
    l["data"][6] = np.float32(np.float16(radius))  # the core stores the radius as half (Scene/LightSource.cpp:105)
    l["data"][7:10] = d
    l["data"][10] = np.float32(np.floor(cos_angle * 65535 + 0.5) / 65535)  # ... and the cosine as unorm16 (:106)
    l["flags"] = capi.LIGHT_SPOT
    return l


def directional_light(radiance, direction):
    l = np.zeros((), capi.LIGHT_DTYPE)
    d = np.asarray(direction, np.float64); d = d / np.linalg.norm(d)
    l["data"][0:3] = radiance; l["data"][3:6] = d
    l["flags"] = capi.LIGHT_DIRECTIONAL
    return l


def _instance(mesh_id, material_id, to_world):
    i = np.zeros((), capi.INSTANCE_DTYPE)
    i["mesh_id"] = mesh_id; i["material_id"] = material_id; i["to_world"] = to_world
    return i


# ---- scenes -------------------------------------------------------------------------------------

def cornell_box(sphere_quads=(100, 50), width=1024, height=1024):
    """BASELINE.json configs[1]: SmallPT-style Cornell box, tessellated spheres (~20k triangles), one sphere light."""
    meshes = {0: plane(1), 1: revolved_sphere(*sphere_quads)}
    materials = np.array([
        material((0, 0, 0), 0.0),                                              # 0: invalid material (Renderer.cpp:821)
        material((0.98, 0.98, 0.98), 1.0, 0.02, thin_walled=True),             # 1: white
        material((0.98, 0.02, 0.02), 1.0, 0.02, thin_walled=True),             # 2: red
        material((0.02, 0.98, 0.02), 1.0, 0.02, thin_walled=True),             # 3: green
        material(IRON_TINT, 0.4, 0.5, metallic=1.0),                           # 4: iron
        material(COPPER_TINT, 0.02, 0.5, metallic=1.0),                        # 5: copper (near mirror)
    ], capi.MATERIAL_DTYPE)
    half_pi = np.pi * 0.5
    fwd, right = (0, 0, 1), (1, 0, 0)
    instances = np.array([
        _instance(0, 1, affine((0, -0.5, 0))),                                                    # floor
        _instance(0, 1, affine((0, 0.5, 0), quat_from_angle_axis(np.pi, fwd))),                   # roof
        _instance(0, 1, affine((0, 0, 0.5), quat_from_angle_axis(-half_pi, right))),              # back
        _instance(0, 2, affine((-0.5, 0, 0), quat_from_angle_axis(-half_pi, fwd))),               # left
        _instance(0, 3, affine((0.5, 0, 0), quat_from_angle_axis(half_pi, fwd))),                 # right
        _instance(1, 4, affine((-0.23, -0.335, 0.12), None, 0.33)),                               # rough iron sphere, r = 0.165
        _instance(1, 5, affine((0.23, -0.335, -0.12), None, 0.33)),                               # smooth copper sphere
    ], capi.INSTANCE_DTYPE)
    lights = np.array([sphere_light((2.0, 2.0, 2.0), (0.0, 0.45, 0.0), 0.05)], capi.LIGHT_DTYPE)
    camera = perspective_camera((0.0, 0.0, -1.5), None, np.pi / 4, width / height)
    return {"name": "cornell_box", "meshes": meshes, "materials": materials, "instances": instances, "lights": lights,
            "environment": {"tint": (0.0, 0.0, 0.0)}, "camera": camera, "width": width, "height": height}


def material_grid(width=1920, height=1080, grid=10, sphere_quads=(100, 50), env_size=(2048, 1024), env_samples=8192):
    """BASELINE.json configs[2]: grid x grid DefaultMaterial spheres sweeping roughness (x) and metallic (z) on a ground plane,
    lit by a procedural HDR environment with importance sampling + MIS (~1M triangles at the default tessellation)."""
    from . import environment
    meshes = {0: plane(1), 1: revolved_sphere(*sphere_quads)}
    mats = [material((0, 0, 0), 0.0), material((0.5, 0.5, 0.5), 0.9, 0.02)]
    instances = [_instance(0, 1, affine((0, 0, 0), None, 3.0 * grid))]
    spacing = 1.25
    for j in range(grid):
        for i in range(grid):
            mats.append(material((0.8, 0.5, 0.3), i / max(grid - 1, 1), 0.04, metallic=j / max(grid - 1, 1)))
            pos = ((i - (grid - 1) / 2) * spacing, 0.5, (j - (grid - 1) / 2) * spacing)
            instances.append(_instance(1, len(mats) - 1, affine(pos)))
    env = environment.build_environment(environment.procedural_sky(*env_size), sample_count=env_samples)
    rot = quat_from_angle_axis(np.radians(32.0), (1, 0, 0))
    extent = grid * spacing
    camera = perspective_camera((0.0, 0.62 * extent + 1.0, -0.95 * extent - 1.0), rot, np.radians(45.0), width / height)
    return {"name": "material_grid", "meshes": meshes, "materials": np.array(mats, capi.MATERIAL_DTYPE),
            "instances": np.array(instances, capi.INSTANCE_DTYPE), "lights": np.zeros(0, capi.LIGHT_DTYPE),
            "environment": env, "camera": camera, "width": width, "height": height}


def instanced_terrain(width=3840, height=2160, instances_per_side=(25, 20), quads_per_edge=224, distinct_meshes=8):
    """BASELINE.json configs[3]: displaced meshes (2 * quads^2 triangles each) instanced on a jittered grid with random rigid
    transforms, 16 materials, sphere + spot + directional lights. Defaults: 500 instances x 100 352 triangles = 50.2M."""
    rng = np.random.default_rng(13)
    meshes = {k: displaced_grid(quads_per_edge, seed=11 + k) for k in range(distinct_meshes)}
    mats = [material((0, 0, 0), 0.0)]
    for k in range(16):
        mats.append(material(tuple(0.25 + 0.6 * rng.random(3)), 0.1 + 0.9 * (k % 4) / 3.0, 0.04, metallic=float(k // 4) / 3.0))
    nx, nz = instances_per_side
    inst = []
    for iz in range(nz):
        for ix in range(nx):
            jitter = rng.uniform(-0.08, 0.08, 2)
            pos = ((ix - (nx - 1) / 2) * 0.98 + jitter[0], rng.uniform(-0.02, 0.02), (iz - (nz - 1) / 2) * 0.98 + jitter[1])
            rot = quat_from_angle_axis(rng.uniform(0, 2 * np.pi), (rng.normal(scale=0.06), 1.0, rng.normal(scale=0.06)))
            inst.append(_instance(int(rng.integers(0, distinct_meshes)), 1 + int(rng.integers(0, 16)), affine(pos, rot, 1.05)))
    lights = np.array([sphere_light((400.0, 380.0, 350.0), (0.0, 6.0, 0.0), 0.5),
                       spot_light((300.0, 300.0, 360.0), (-6.0, 5.0, -4.0), 0.2, (0.7, -0.6, 0.4), 0.8),
                       directional_light((1.2, 1.1, 1.0), (0.3, -1.0, 0.2))], capi.LIGHT_DTYPE)
    rot = quat_from_angle_axis(np.radians(38.0), (1, 0, 0))
    camera = perspective_camera((0.0, 9.5, -13.5), rot, np.radians(45.0), width / height)
    return {"name": "instanced_terrain", "meshes": meshes, "materials": np.array(mats, capi.MATERIAL_DTYPE),
            "instances": np.array(inst, capi.INSTANCE_DTYPE), "lights": lights, "environment": {"tint": (0.02, 0.03, 0.05)},
            "camera": camera, "width": width, "height": height}


def random_triangles(n, seed, extent=1.0, size=0.2):
    """Triangle soup for traversal parity tests."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-extent, extent, (n, 1, 3))
    p = (c + rng.uniform(-size, size, (n, 3, 3))).astype(np.float32).reshape(-1, 3)
    idx = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    return {"indices": idx, "positions": p}


def triangle_count(scene):
    return int(sum(scene["meshes"][int(i["mesh_id"])]["indices"].shape[0] for i in scene["instances"]))


def upload(ctx, scene):
    """Feeds a scene dict through the C ABI (mirrors Renderer::handle_updates, Renderer.cpp:578-1205)."""
    for texture_id, tex in scene.get("textures", {}).items():
        ctx.upload_texture(texture_id, tex["pixels"], tex.get("srgb", False), tex.get("wrap_u", capi.WRAP_REPEAT), tex.get("wrap_v", capi.WRAP_REPEAT),
                           tex.get("linear", True))
    for mesh_id, m in scene["meshes"].items():
        ctx.upload_mesh(mesh_id, m["indices"], m["positions"], m.get("normals"), m.get("texcoords"), m.get("tints"))
        if m.get("emission") is not None:
            ctx.set_mesh_emission(mesh_id, m["emission"])
    ctx.set_materials(scene["materials"])
    ctx.set_instances(scene["instances"])
    ctx.set_lights(scene["lights"])
    env = scene.get("environment", {"tint": (0, 0, 0)})
    ctx.set_environment(env["tint"], env.get("texels"), env.get("per_pixel_pdf"), env.get("samples"))
    if env.get("marginal_cdf") is not None and env["per_pixel_pdf"].shape == env["conditional_cdf"][:, 1:].shape:
        ctx.set_environment_cdfs(env["marginal_cdf"], env["conditional_cdf"])
    ctx.set_environment_sampling(env.get("nee", "presampled"))
    ctx.build_accel()
