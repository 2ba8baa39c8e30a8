"""CPU tests that pin the oracle to the reference's own golden vectors (no GPU needed)."""
import subprocess

import numpy as np
import pytest

from tests import oracle_lib

pytestmark = pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")

GTESTS = oracle_lib.REPO / "oracle" / "_ref" / "ref_gtests"


def test_reference_gtest_suite_passes():
    """The reference's own host-compiled OptiXRendererTests cases (85) pass with the oracle toolchain."""
    out = subprocess.run([str(GTESTS)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "[  PASSED  ] 85 tests." in out.stdout


def test_pod_sizes(ref):
    # sizes measured from the reference's host build.
    expected = {"Material": 64, "Light": 48, "LightSample": 32, "BSDFSample": 32, "BSDFResponse": 16,
                "MonteCarloPayload": 160, "VertexGeometry": 16, "DefaultShading": 44}
    for name, size in expected.items():
        assert ref.sizeof(name) == size, name


def test_default_shading_regression_vectors(ref):
    """tests/OptiXRendererTests/ShadingModels/DefaultShadingTest.h:410-447 through the oracle's C API."""
    golden = np.array([
        [497358.250000, 380976.437500, 167112.35938, 497357.968750], [124339.296875, 95243.906250, 41778.00000, 124339.195313],
        [994714.562500, 762453.062500, 335647.75000, 703369.687500], [249080.015625, 190921.531250, 84049.10156, 175985.171875],
        [4957685248.0, 4900781568.0, 4796215808.0, 49668972.0], [1455754624.0, 1439689728.0, 1410168448.0, 13442245.0],
        [0.011624, 0.076557, 0.09214, 0.010905], [0.012486, 0.092185, 0.11131, 0.230607],
        [0.012840, 0.122771, 0.14915, 0.034218], [0.011330, 0.121562, 0.14802, 0.254778],
        [0.051809, 0.085369, 0.09342, 0.286622], [0.013969, 0.145090, 0.17656, 0.218950],
        [0.019217, 0.081176, 0.09605, 0.0164565], [0.019548, 0.0975228, 0.116237, 0.228887],
        [0.017939, 0.128357, 0.15486, 0.0377722], [0.014534, 0.125507, 0.15214, 0.239682],
        [0.088401, 0.115091, 0.12150, 0.317704], [0.018240, 0.147322, 0.17830, 0.192018]], np.float64)
    tuples = regression_inputs(ref)
    out = ref.bsdf_eval_sample_pdf(0, **tuples)
    got = np.concatenate([out["sample_f"], np.abs(out["sample_pdf"])[:, None]], axis=1).astype(np.float64)
    assert np.all(np.abs(got - golden) <= golden * 1e-4), np.abs(got - golden) / golden


def regression_inputs(ref):
    """gold / plastic / coated plastic x 3 wo x 2 sample02 points (ShadingModelTestUtils.h:23-47)."""
    mats = [((1.0, 0.766, 0.336), 0.02, 1.0, 1.0, 0.0, 0.0), ((0.02, 0.27, 0.33), 0.7, 0.0, 0.02, 0.0, 0.0),
            ((0.02, 0.27, 0.33), 0.7, 0.0, 0.02, 1.0, 0.7)]
    def normalize(v):
        v = np.array(v, np.float32); return v * np.float32(1.0 / np.sqrt(np.float32(np.dot(v, v))))
    wos = [np.array([0, 0, 1], np.float32), normalize([1, 0, 1]), normalize([1, 0, 0.01])]
    s02 = ref.sample02(2)
    rows = []
    for tint, rough, metal, spec, coat, coat_r in mats:
        for wo in wos:
            for s in range(2):
                rows.append((wo, tint, (rough, metal, spec), (s02[s, 0], s02[s, 1], (s + 0.5) / 2), (coat, coat_r)))
    n = len(rows)
    return {"wo": np.array([r[0] for r in rows], np.float32), "wi": np.tile(np.array([0, 0, 1], np.float32), (n, 1)),
            "tint": np.array([r[1] for r in rows], np.float32), "rms": np.array([r[2] for r in rows], np.float32),
            "u": np.array([r[3] for r in rows], np.float32), "coat": np.array([r[4] for r in rows], np.float32)}


# TransmissiveShadingTest.h:203-238: frosted glass, 4 cos_theta_o x 2 sample02 points -> {reflectance, |PDF|}
TRANSMISSIVE_REGRESSION = np.array([
    [102.196815, 102.196815, 102.196815, 70.955925], [30.308733, 30.308733, 30.308733, 19.304911],
    [4075.826172, 4075.826172, 4075.826172, 445.397308], [4660.390625, 4660.390625, 4660.390625, 235.904617],
    [610.321655, 623.170593, 610.321655, 504.575867], [149.033539, 152.171082, 149.033539, 125.492897],
    [1633.225708, 1667.609497, 1633.225708, 1715.760010], [408.740875, 417.345978, 408.740875, 429.358185]], np.float32)


def transmissive_regression_inputs(ref):
    """frosted_glass_parameters() (TransmissiveShadingTest.h:25-38) at cos_theta_o in {-0.7, -0.1, 0.4, 1}."""
    s02 = ref.sample02(2)
    rows = []
    for cos_theta_o in (np.float32(-0.7), np.float32(-0.1), np.float32(0.4), np.float32(1.0)):
        wo = np.array([np.sqrt(np.float32(1) - cos_theta_o * cos_theta_o), 0, abs(cos_theta_o)], np.float32)
        for s in range(2):
            rows.append((wo, (0.95, 0.97, 0.95), (0.2, cos_theta_o, 0.04), (s02[s, 0], s02[s, 1], (s + 0.5) / 2)))
    n = len(rows)
    return {"wo": np.array([r[0] for r in rows], np.float32), "wi": np.tile(np.array([0, 0, 1], np.float32), (n, 1)),
            "tint": np.array([r[1] for r in rows], np.float32), "rms": np.array([r[2] for r in rows], np.float32),
            "u": np.array([r[3] for r in rows], np.float32)}


def test_transmissive_shading_regression_vectors(ref):
    """The oracle reproduces the golden vectors of the reference's own TransmissiveShadingModel.regression_test."""
    got = ref.bsdf_eval_sample_pdf(4, **transmissive_regression_inputs(ref))
    values = np.concatenate([got["sample_f"], np.abs(got["sample_pdf"])[:, None]], axis=1).astype(np.float64)
    golden = TRANSMISSIVE_REGRESSION.astype(np.float64)
    assert np.all(np.abs(values - golden) <= golden * 1e-4), np.abs(values - golden) / golden


def test_dielectric_tables_match_committed_data(ref):
    from bifrost3d_b200.capi import default_dielectric_tables
    light, dense, dims = ref.dielectric_tables()
    assert list(dims) == [16] * 3
    tl, td = default_dielectric_tables()
    assert np.array_equal(light, tl) and np.array_equal(dense, td)
    assert np.all(light[0::2] >= light[1::2]) and np.all(dense[0::2] >= dense[1::2])  # total rho >= reflected rho


def test_tables_match_committed_data(ref):
    from bifrost3d_b200.capi import default_tables
    a, b, c, dims = ref.tables()
    assert list(dims) == [32] * 6
    ta, tb, tc = default_tables()
    assert np.array_equal(a, ta) and np.array_equal(b, tb) and np.array_equal(c, tc)


def test_golden_fixture_matches_oracle(ref):
    """tests/golden/bsdf_c1_small.npz was generated from this oracle (tests/golden/make_golden.py)."""
    from tests.golden import make_golden
    make_golden.check(ref)


# ---- the integrator restatement on its own (CPU only): properties that do not need the GPU -------------------------

def _cornell(**kw):
    from bifrost3d_b200 import scenes
    return scenes.cornell_box(sphere_quads=(8, 4), **kw)


def test_oracle_russian_roulette_keeps_the_expectation():
    """The opt-in roulette restated in the oracle is unbiased: same mean image (within Monte Carlo noise), fewer rays."""
    scene = _cornell()
    sc = oracle_lib.OracleScene(scene)
    plain, rays_plain = sc.render(scene["camera"], 24, 24, 0, 256, max_bounces=8)
    rr, rays_rr = sc.render(scene["camera"], 24, 24, 0, 256, max_bounces=8, russian_roulette_start=2)
    sc.close()
    mean_plain = (plain[..., :3] / plain[..., 3:4]).mean(); mean_rr = (rr[..., :3] / rr[..., 3:4]).mean()
    assert abs(mean_rr / mean_plain - 1.0) < 0.02
    assert rays_rr[0] < 0.85 * rays_plain[0]


def test_oracle_per_vertex_emission_scales_the_material_emission():
    """TriangleAttributes.cu:78-83: a constant per-vertex emission scale s multiplies the emitted radiance by s."""
    from bifrost3d_b200 import capi
    scene = _cornell()
    scene["lights"] = np.zeros(0, capi.LIGHT_DTYPE)
    mats = scene["materials"].copy(); mats[4]["emission"] = (0.5, 0.25, 0.125); scene["materials"] = mats
    images = []
    for scale in (1.0, 3.0):
        scene["meshes"][1]["emission"] = np.full((scene["meshes"][1]["positions"].shape[0], 3), scale, np.float32)
        sc = oracle_lib.OracleScene(scene)
        accum, _ = sc.render(scene["camera"], 24, 24, 0, 8, max_bounces=1)
        sc.close()
        images.append(accum[..., :3] / accum[..., 3:4])
    assert images[0].max() > 0.1
    assert np.allclose(images[1], 3.0 * images[0], rtol=1e-5, atol=1e-7)


def test_oracle_coverage_texture_controls_visibility():
    """A cutout texture that is 0 on one half of a quad lets rays through that half only (Material::get_coverage)."""
    from bifrost3d_b200 import capi, scenes
    quad = scenes.plane(1)
    tex = np.zeros((2, 2, 1), np.uint8); tex[:, 1] = 255           # u < 0.5 transparent, u > 0.5 opaque
    mat = scenes.material((1, 1, 1), 1.0, thin_walled=True); mat["flags"] = 3; mat["coverage"] = 0.5; mat["coverage_texture_id"] = 1
    scene = {"meshes": {0: quad}, "materials": np.array([scenes.material((0, 0, 0), 0), mat], capi.MATERIAL_DTYPE),
             "instances": np.array([scenes._instance(0, 1, scenes.affine())], capi.INSTANCE_DTYPE), "lights": np.zeros(0, capi.LIGHT_DTYPE),
             "environment": {"tint": (0, 0, 0)}, "textures": {1: {"pixels": tex, "linear": False, "wrap_u": capi.WRAP_CLAMP, "wrap_v": capi.WRAP_CLAMP}}}
    sc = oracle_lib.OracleScene(scene)
    x = np.float32([-0.25, 0.25])                                    # texcoord u = x + 0.5
    o = np.stack([x, np.full(2, 1.0, np.float32), np.zeros(2, np.float32)], axis=1)
    d = np.tile(np.float32([0, -1, 0]), (2, 1))
    _, _, _, occluded = sc.intersect(o, d, brute=True)
    sc.close()
    assert list(occluded) == [0, 1]
