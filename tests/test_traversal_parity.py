"""BVH build + traversal parity. Closest-hit primitive indices, t and barycentrics must be bit exact against the
oracle's brute force intersector (the reference delegates traversal to closed source OptiX: parity unpinned by the
reference, the brute force CPU loop is the oracle)."""
import numpy as np
import pytest

from tests import oracle_lib
from bifrost3d_b200 import scenes, capi

needs_oracle = pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")


def soup_scene(n, seed, **kw):
    mesh = scenes.random_triangles(n, seed, **kw)
    mats = np.array([scenes.material((0, 0, 0), 0.0), scenes.material((0.5, 0.5, 0.5), 0.5)], capi.MATERIAL_DTYPE)
    inst = np.array([scenes._instance(0, 1, scenes.affine())], capi.INSTANCE_DTYPE)
    return {"meshes": {0: mesh}, "materials": mats, "instances": inst, "lights": np.zeros(0, capi.LIGHT_DTYPE), "environment": {"tint": (0, 0, 0)}}


def random_rays(n, seed, extent=1.5):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    # some axis aligned and some rays that start on geometry
    d[: n // 50] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 50)] * rng.choice([-1.0, 1.0], (n // 50, 1)).astype(np.float32)
    return o, d


# ---- CPU only: the oracle's BVH equals its brute force loop -------------------------------------------

@needs_oracle
def test_oracle_bvh_equals_brute_force():
    sc = oracle_lib.OracleScene(soup_scene(3000, 1))
    o, d = random_rays(20000, 2)
    pb, tb, uvb, ob = sc.intersect(o, d, brute=True)
    pv, tv, uvv, ov = sc.intersect(o, d, brute=False)
    assert np.array_equal(pb, pv) and np.array_equal(tb, tv) and np.array_equal(uvb, uvv) and np.array_equal(ob, ov)
    assert (pb >= 0).mean() > 0.3
    sc.close()


@needs_oracle
def test_oracle_flatten_matches_numpy():
    scene = scenes.cornell_box(sphere_quads=(12, 6))
    sc = oracle_lib.OracleScene(scene)
    assert sc.triangle_count() == scenes.triangle_count(scene)
    wv = sc.world_vertices()
    inst = scene["instances"][5]; mesh = scene["meshes"][1]
    m = inst["to_world"].reshape(3, 4)
    p = mesh["positions"][mesh["indices"][7]]
    expect = ((m[:, 0] * p[:, 0:1] + m[:, 1] * p[:, 1:2]) + m[:, 2] * p[:, 2:3]) + m[:, 3]
    first = 5 * 2  # five walls of two triangles precede the first sphere
    assert np.array_equal(wv[first + 7], expect.astype(np.float32))
    sc.close()


# ---- GPU ----------------------------------------------------------------------------------------------

@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("n_tris,seed", [(1, 3), (2, 4), (5, 5), (4000, 6), (50000, 7)])
def test_closest_hit_bit_exact_vs_brute_force(tracer, n_tris, seed):
    bpt = tracer
    scene = soup_scene(n_tris, seed)
    scenes.upload(bpt, scene)
    info = bpt.accel_info()
    assert info["triangles"] == n_tris
    sc = oracle_lib.OracleScene(scene)
    n_rays = 200000 if n_tris <= 4000 else 40000
    o, d = random_rays(n_rays, seed + 100)
    if n_tris < 10:
        # aim at the triangles
        wv = sc.world_vertices().reshape(-1, 3)
        rng = np.random.default_rng(seed)
        target = wv[rng.integers(0, wv.shape[0], n_rays)] + rng.normal(scale=0.05, size=(n_rays, 3)).astype(np.float32)
        d = (target - o); d /= np.linalg.norm(d, axis=1, keepdims=True); d = d.astype(np.float32)
    gp, gt, guv, gocc = bpt.intersect(o, d)
    rp, rt, ruv, rocc = sc.intersect(o, d, brute=True)
    mismatch = gp != rp
    # north_star: "bit-exact against a brute-force CPU intersector except for exact-distance ties" - ties are resolved
    # identically here (lower primitive index), so no exemption is needed.
    assert not mismatch.any(), f"{mismatch.sum()} of {n_rays} primitive ids differ, e.g. {np.flatnonzero(mismatch)[:5]}"
    hit = rp >= 0
    assert np.array_equal(gt[hit], rt[hit])
    assert np.array_equal(guv[hit], ruv[hit])
    assert np.array_equal(gocc, rocc)
    assert hit.mean() > 0.05
    sc.close()


@pytest.mark.gpu
@needs_oracle
def test_interval_and_occlusion_semantics(tracer):
    bpt = tracer
    scene = soup_scene(2000, 11)
    scenes.upload(bpt, scene)
    sc = oracle_lib.OracleScene(scene)
    o, d = random_rays(50000, 12)
    _, t_first, _, _ = sc.intersect(o, d, brute=True)
    rng = np.random.default_rng(0)
    # intervals that end exactly at / just before / just after the first hit, and that start exactly at it
    tmax = np.where(np.isfinite(t_first), t_first, 1.0).astype(np.float32)
    choice = rng.integers(0, 4, o.shape[0])
    tmax = np.where(choice == 1, np.nextafter(tmax, np.float32(0)), np.where(choice == 2, np.nextafter(tmax, np.float32(np.inf)), tmax)).astype(np.float32)
    tmin = np.where(choice == 3, np.where(np.isfinite(t_first), t_first, 0), 0).astype(np.float32)
    tmax = np.where(choice == 3, np.float32(1e30), tmax)
    gp, gt, guv, gocc = bpt.intersect(o, d, tmin, tmax)
    rp, rt, ruv, rocc = sc.intersect(o, d, tmin, tmax, brute=True)
    assert np.array_equal(gp, rp)
    assert np.array_equal(gocc, rocc)
    sc.close()


@pytest.mark.gpu
@needs_oracle
def test_cornell_box_instances_and_flattening(tracer):
    bpt = tracer
    scene = scenes.cornell_box(sphere_quads=(40, 20))
    scenes.upload(bpt, scene)
    sc = oracle_lib.OracleScene(scene)
    assert bpt.accel_info()["triangles"] == sc.triangle_count()
    rng = np.random.default_rng(5)
    n = 100000
    o = rng.uniform(-0.45, 0.45, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    gp, gt, guv, _ = bpt.intersect(o, d, want_occluded=False)
    rp, rt, ruv, _ = sc.intersect(o, d, brute=False)
    assert np.array_equal(gp, rp)
    assert np.array_equal(gt, rt)
    assert np.array_equal(guv, ruv)
    assert (rp >= 0).mean() > 0.8  # five walls, open towards the camera
    sc.close()


@pytest.mark.gpu
def test_empty_scene_misses_everything(bpt):
    scene = soup_scene(1, 1)
    scene["instances"] = np.zeros(0, capi.INSTANCE_DTYPE)
    scenes.upload(bpt, scene)
    o, d = random_rays(1000, 1)
    p, t, _, occ = bpt.intersect(o, d)
    assert (p == -1).all() and np.isinf(t).all() and (occ == 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("environment,node_width", [({"BPT_CW": "1"}, 8), ({"BPT_CW": "0", "BPT_WIDE": "0"}, 2), ({"BPT_BVH": "lbvh"}, 4),
                                                    ({"BPT_BVH": "lbvh", "BPT_CW": "1"}, 8)])
def test_every_hierarchy_returns_the_same_hits(bpt, environment, node_width, monkeypatch):
    """The four-wide PLOC hierarchy (what a scene of this size gets), the compressed eight-wide nodes (BPT_CW=1: what scenes
    from 131 072 triangles on get), the binary nodes (the fallback for trees too deep for either stack) and the plain Morton
    hierarchy (BPT_BVH=lbvh: the fallback when PLOC gives up) must return identical hits: the result is defined by
    min (t, primitive id), not by the traversal order."""
    scene = soup_scene(20000, 11, size=0.08)
    o, d = random_rays(60000, 12)
    scenes.upload(bpt, scene)
    assert bpt.accel_info()["node_width"] == 4
    want = bpt.intersect(o, d)
    for variable, value in environment.items():
        monkeypatch.setenv(variable, value)
    other = capi.Bpt(0)
    scenes.upload(other, scene)
    assert other.accel_info()["node_width"] == node_width
    got = other.intersect(o, d)
    other.close()
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert (want[0] >= 0).mean() > 0.2


@pytest.mark.gpu
@needs_oracle
def test_degenerate_geometry_ties_and_large_coordinates(tracer):
    """Coincident triangles (exact ties in t resolve to the lower primitive id), zero-area triangles (never hit, must not
    break the build), identical centroids (equal Morton codes) and a scene 10 km from the origin, bit exact vs brute force."""
    rng = np.random.default_rng(31)
    base = scenes.random_triangles(600, 32, extent=1.0, size=0.3)
    pos = base["positions"].reshape(-1, 3, 3)
    dup = np.concatenate([pos, pos[:200], pos[:50]])                        # duplicates: ties in t
    zero = np.repeat(rng.uniform(-1, 1, (100, 1, 3)).astype(np.float32), 3, axis=1)          # three coincident vertices
    sliver = pos[:100].copy(); sliver[:, 2] = sliver[:, 0] + (sliver[:, 1] - sliver[:, 0]) * 0.5  # collinear vertices
    same_centroid = np.tile(pos[:1], (64, 1, 1))                              # 64 identical triangles
    tris = (np.concatenate([dup, zero, sliver, same_centroid]) + np.float32([10000.0, -2500.0, 400.0])).astype(np.float32)
    mesh = {"indices": np.arange(3 * tris.shape[0], dtype=np.uint32).reshape(-1, 3), "positions": tris.reshape(-1, 3)}
    mats = np.array([scenes.material((0, 0, 0), 0.0), scenes.material((0.5, 0.5, 0.5), 0.5)], capi.MATERIAL_DTYPE)
    scene = {"meshes": {0: mesh}, "materials": mats, "instances": np.array([scenes._instance(0, 1, scenes.affine())], capi.INSTANCE_DTYPE),
             "lights": np.zeros(0, capi.LIGHT_DTYPE), "environment": {"tint": (0, 0, 0)}}
    bpt = tracer
    scenes.upload(bpt, scene)
    sc = oracle_lib.OracleScene(scene)
    o, d = random_rays(60000, 33)
    o = (o + np.float32([10000.0, -2500.0, 400.0])).astype(np.float32)
    gp, gt, guv, gocc = bpt.intersect(o, d)
    rp, rt, ruv, rocc = sc.intersect(o, d, brute=True)
    assert np.array_equal(gp, rp)
    hit = rp >= 0
    assert np.array_equal(gt[hit], rt[hit]) and np.array_equal(guv[hit], ruv[hit]) and np.array_equal(gocc, rocc)
    assert hit.mean() > 0.2
    n_dup = pos.shape[0]
    assert not np.isin(rp, np.arange(n_dup, n_dup + 250)).any()            # a duplicate never wins against its lower-id twin
    sc.close()


@pytest.mark.gpu
def test_odd_frame_sizes_and_dark_scene(bpt):
    """Frames that are not a multiple of the warp size, a 1 x 1 frame, and a scene without any light (every path dies black)."""
    scene = scenes.cornell_box(sphere_quads=(8, 4))
    scenes.upload(bpt, scene)
    for w, h in ((33, 17), (1, 1), (7, 129)):
        bpt.render(scene["camera"], w, h, 0, 3, reset=True)
        img = bpt.resolve_float4()
        assert img.shape == (h, w, 4) and np.isfinite(img).all()
    dark = dict(scene); dark["lights"] = np.zeros(0, capi.LIGHT_DTYPE)
    scenes.upload(bpt, dark)
    bpt.render(dark["camera"], 33, 17, 0, 2, reset=True)
    assert np.all(bpt.resolve_float4()[..., :3] == 0.0)


# ---- the sizes that are benchmarked ("parity at the sizes you benchmark") ---------------------------------

def camera_and_bounce_rays(scene, n, seed):
    """Half camera-like rays towards the scene, half incoherent rays that start near the geometry."""
    rng = np.random.default_rng(seed)
    wv = None
    o = np.empty((n, 3), np.float32); d = np.empty((n, 3), np.float32)
    half = n // 2
    lo, hi = scene["bounds"]
    centre, extent = (lo + hi) / 2, (hi - lo)
    o[:half] = (centre + np.float32([0, 1.5, -2.0]) * np.linalg.norm(extent) * 0.5).astype(np.float32)
    target = rng.uniform(lo, hi, (half, 3))
    d[:half] = (target - o[:half]).astype(np.float32)
    o[half:] = rng.uniform(lo - 0.05 * extent, hi + 0.05 * extent, (n - half, 3)).astype(np.float32)
    d[half:] = rng.normal(size=(n - half, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return o, d.astype(np.float32)


def scene_bounds(sc):
    wv = sc.world_vertices().reshape(-1, 3)
    return wv.min(axis=0).astype(np.float64), wv.max(axis=0).astype(np.float64)


@pytest.mark.gpu
@needs_oracle
def test_material_grid_980k_triangles_bit_exact(bpt):
    """configs[2]'s geometry at full size (980 002 triangles): 2^20 rays against the oracle's BVH, which
    test_oracle_bvh_equals_brute_force proves equal to the brute-force loop."""
    scene = scenes.material_grid()
    scenes.upload(bpt, scene)
    sc = oracle_lib.OracleScene(scene)
    assert bpt.accel_info()["triangles"] == sc.triangle_count() == 980002 and bpt.accel_info()["node_width"] == 8
    scene["bounds"] = scene_bounds(sc)
    o, d = camera_and_bounce_rays(scene, 1 << 20, 21)
    gp, gt, guv, gocc = bpt.intersect(o, d)
    rp, rt, ruv, rocc = sc.intersect(o, d, brute=False)
    assert np.array_equal(gp, rp), f"{(gp != rp).sum()} primitive ids differ"
    hit = rp >= 0
    assert np.array_equal(gt[hit], rt[hit]) and np.array_equal(guv[hit], ruv[hit]) and np.array_equal(gocc, rocc)
    assert hit.mean() > 0.2
    assert bpt.counters()["traversal_stack_overflows"] == 0
    sc.close()


@pytest.mark.gpu
@needs_oracle
def test_terrain_above_8m_triangles_bit_exact(bpt):
    """A subset of configs[3] with more than 8 000 000 triangles, so that the large-scene traversal parameters
    (traversal_budget_for / traversal_min_active_for in bpt_trace.cuh) are the ones that run: 2^19 rays, bit exact."""
    scene = scenes.instanced_terrain(480, 270, (12, 7), 224, 8)  # 84 instances x 100 352 triangles = 8.43 M
    scenes.upload(bpt, scene)
    sc = oracle_lib.OracleScene(scene)
    n_tris = sc.triangle_count()
    assert bpt.accel_info()["triangles"] == n_tris and n_tris > 8_000_000
    scene["bounds"] = scene_bounds(sc)
    o, d = camera_and_bounce_rays(scene, 1 << 19, 22)
    gp, gt, guv, gocc = bpt.intersect(o, d)
    rp, rt, ruv, rocc = sc.intersect(o, d, brute=False)
    assert np.array_equal(gp, rp), f"{(gp != rp).sum()} primitive ids differ"
    hit = rp >= 0
    assert np.array_equal(gt[hit], rt[hit]) and np.array_equal(guv[hit], ruv[hit]) and np.array_equal(gocc, rocc)
    assert hit.mean() > 0.2
    assert bpt.counters()["traversal_stack_overflows"] == 0
    sc.close()


def nested_slivers(count=80, growth=1.35):
    """Triangles of geometrically growing size stacked a hair apart around one axis: every agglomerative build turns them
    into a chain (each merge joins one more triangle to the cluster of all smaller ones), i.e. a hierarchy about as deep as
    it has leaves. A ray down the axis meets every box, and going near to far it defers up to three siblings per level."""
    positions, size = [], 0.01
    for k in range(count):
        z = 1e-4 * k
        positions += [(-size, -size, z), (size, -size, z), (0.0, 1.5 * size, z)]
        size *= growth
    p = np.array(positions, np.float32)
    return {"indices": np.arange(3 * count, dtype=np.uint32).reshape(count, 3), "positions": p}


def deep_hierarchy_rays(n, seed):
    rng = np.random.default_rng(seed)
    o = np.concatenate([rng.normal(scale=0.004, size=(n, 2)), np.full((n, 1), -1.0)], axis=1).astype(np.float32)
    o[n // 2:, 2] = 1.0  # half of the rays come from the far side: the nearest triangle is the largest one
    d = np.tile(np.float32([0, 0, 1]), (n, 1)); d[n // 2:, 2] = -1.0
    d[:, :2] += rng.normal(scale=1e-3, size=(n, 2)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    return o, d.astype(np.float32)


@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("environment", [{"BPT_CW": "1"}, {"BPT_CW": "0"}])
def test_deep_hierarchy_uses_the_spill_stack(environment, monkeypatch):
    """Forces traversal stacks deeper than the 32 shared-memory entries of the four-wide traversal (the local-memory spill
    part; BPT_CW=0) and checks the result against brute force, for closest-hit and for any-hit rays through partially
    covering surfaces (which never terminate early and therefore walk the whole chain). The eight-wide nodes swallow seven
    links of such a chain per node: the same scene through them is the other case."""
    for variable, value in environment.items():
        monkeypatch.setenv(variable, value)
    bpt = capi.Bpt(0)
    mesh = nested_slivers()
    mats = np.array([scenes.material((0, 0, 0), 0.0), scenes.material((0.5, 0.5, 0.5), 0.5)], capi.MATERIAL_DTYPE)
    mats[1]["coverage"] = 0.01
    scene = {"meshes": {0: mesh}, "materials": mats, "instances": np.array([scenes._instance(0, 1, scenes.affine())], capi.INSTANCE_DTYPE),
             "lights": np.zeros(0, capi.LIGHT_DTYPE), "environment": {"tint": (0, 0, 0)}}
    scenes.upload(bpt, scene)
    assert bpt.accel_info()["node_width"] == (8 if environment["BPT_CW"] == "1" else 4)
    sc = oracle_lib.OracleScene(scene)
    o, d = deep_hierarchy_rays(20000, 31)
    bpt.counters(reset=True)
    gp, gt, guv, gocc = bpt.intersect(o, d)
    rp, rt, ruv, rocc = sc.intersect(o, d, brute=True)
    assert np.array_equal(gp, rp) and np.array_equal(gocc, rocc)
    hit = rp >= 0
    assert hit.mean() > 0.9 and np.array_equal(gt[hit], rt[hit]) and np.array_equal(guv[hit], ruv[hit])
    assert bpt.counters()["traversal_stack_overflows"] == 0
    sc.close(); bpt.close()


def nested_clusters(count=150, growth=1.15, per_cluster=5):
    """Like nested_slivers, but every link of the chain is a small fan of triangles, i.e. an inner node: an eight-wide node
    over such a chain holds up to seven inner children plus the rest of the chain, a ray down the axis hits them all and
    leaves the unvisited ones on the stack at every level."""
    positions, size = [], 0.01
    for k in range(count):
        z = 1e-4 * k
        for f in range(per_cluster):
            a0, a1 = 2 * np.pi * f / per_cluster, 2 * np.pi * (f + 1) / per_cluster
            positions += [(0.0, 0.0, z), (size * np.cos(a0), size * np.sin(a0), z), (size * np.cos(a1), size * np.sin(a1), z)]
        size *= growth
    p = np.array(positions, np.float32)
    return {"indices": np.arange(p.shape[0], dtype=np.uint32).reshape(-1, 3), "positions": p}


@pytest.mark.gpu
@needs_oracle
def test_deep_hierarchy_of_inner_nodes(bpt8):
    """A chain whose links are inner nodes (see nested_clusters): whichever node format the build settles on for its depth,
    closest hits and transmissions equal brute force and no push is dropped."""
    bpt = bpt8
    mesh = nested_clusters()
    mats = np.array([scenes.material((0, 0, 0), 0.0), scenes.material((0.5, 0.5, 0.5), 0.5)], capi.MATERIAL_DTYPE)
    mats[1]["coverage"] = 0.01
    scene = {"meshes": {0: mesh}, "materials": mats, "instances": np.array([scenes._instance(0, 1, scenes.affine())], capi.INSTANCE_DTYPE),
             "lights": np.zeros(0, capi.LIGHT_DTYPE), "environment": {"tint": (0, 0, 0)}}
    scenes.upload(bpt, scene)
    info = bpt.accel_info()
    print("nested clusters:", info)
    sc = oracle_lib.OracleScene(scene)
    o, d = deep_hierarchy_rays(20000, 37)
    bpt.counters(reset=True)
    gp, gt, guv, gocc = bpt.intersect(o, d)
    rp, rt, ruv, rocc = sc.intersect(o, d, brute=True)
    assert np.array_equal(gp, rp) and np.array_equal(gocc, rocc)
    hit = rp >= 0
    assert hit.mean() > 0.9 and np.array_equal(gt[hit], rt[hit]) and np.array_equal(guv[hit], ruv[hit])
    assert bpt.counters()["traversal_stack_overflows"] == 0
    sc.close()
