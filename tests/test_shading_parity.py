"""GPU parity of the BSDF / shading-model / light / RNG kernels against the reference's host-compiled
headers (oracle/_ref, kind "reference") and against the committed golden fixtures. All calls go through
the C ABI (libbpt.so)."""
import numpy as np
import pytest

from tests import oracle_lib
from tests.parity import REL_TOL, pdf_class, rel_err, report
from bifrost3d_b200.workloads import bsdf_tuples, transmissive_tuples
from bifrost3d_b200 import capi

pytestmark = pytest.mark.gpu

KINDS = {"default": 0, "ggx_r": 1, "oren_nayar": 2, "burley": 3}

# The north star's bound is 1e-5 relative. What B200 measures against the host-compiled reference headers (round 2, 262 144
# tuples per BSDF, printed by every run of these tests):
#   evaluate_with_PDF (f and PDF): every element within 1e-5 (max 5e-7; bit-exact for GGX_R, Oren-Nayar, Burley) -> NO allowance.
#   sample (f, PDF, direction): at most 6 of 786 432 elements above 1e-5, the largest 1.3e-4 (Oren-Nayar PDF). All of them are
#     tuples where the sampled direction lies within ~1e-3 of the horizon: the samplers end in sqrt(1 - x^2 - y^2) (clipped LTC,
#     Distributions.h:214-216; bounded VNDF, :411-414) whose relative condition number is 1 / (1 - x^2 - y^2), and x, y carry the
#     1 ulp by which glibc's sinf / cosf (not correctly rounded in ~1e-3 of the arguments) differ from the correctly rounded
#     values the device computes. The allowance below is sized to that: 2e-5 of the elements, none above 5e-4; the offending
#     tuples are printed so that a change in their number or kind shows up in the log.
SAMPLE_OUTLIER_FRACTION = 2e-5
SAMPLE_OUTLIER_REL_TOL = 5e-4


def allowed_outliers(n):
    return max(2, int(np.ceil(SAMPLE_OUTLIER_FRACTION * n)))


def compare_bsdf(got, want, what):
    msgs = []
    for key in ("eval_pdf", "sample_pdf"):
        cg, cw = pdf_class(got[key]), pdf_class(want[key])
        mismatch = cg != cw
        # a PDF straddling the 1e-6 validity threshold by rounding is a legitimate class flip
        near_threshold = np.abs(np.abs(want[key]) - 1e-6) < 1e-9
        count = int(np.sum(mismatch & ~near_threshold))
        limit = 0 if key == "eval_pdf" else allowed_outliers(mismatch.size)
        assert count <= limit, f"{what}.{key}: {count} PDF class mismatches"
    for key, floor in (("eval_f", 1e-6), ("eval_pdf", 1e-6), ("sample_f", 1e-6), ("sample_pdf", 1e-6)):
        # Compare samples only where both agree the sample is usable (the reflectance of an invalid sample is unspecified).
        e = rel_err(got[key], want[key], floor)
        rows = np.arange(e.shape[0])
        if key.startswith("sample"):
            valid = pdf_class(want["sample_pdf"]) == pdf_class(got["sample_pdf"])
            e = e[valid]; rows = rows[valid]
        msgs.append(report(f"{what}.{key}", e, REL_TOL))
        bad = np.argwhere(e > REL_TOL)
        for index in bad[:8]:
            row = rows[index[0]]
            msgs.append(f"    tuple {row}: got {np.atleast_1d(got[key][row])} want {np.atleast_1d(want[key][row])} sampled direction z {want['sample_dir'][row][2]:.3e}")
        if key.startswith("eval"):
            assert bad.shape[0] == 0, "\n".join(msgs[-9:])
        else:
            assert bad.shape[0] <= allowed_outliers(e.size), "\n".join(msgs[-9:])
            assert np.all(e[np.isfinite(e)] <= SAMPLE_OUTLIER_REL_TOL), "\n".join(msgs[-9:])
    valid = (pdf_class(want["sample_pdf"]) == 3) & (pdf_class(got["sample_pdf"]) == 3)
    d = np.abs(got["sample_dir"][valid].astype(np.float64) - want["sample_dir"][valid])
    msgs.append(f"{what}.sample_dir: max abs err {d.max() if d.size else 0:.3e}, {int(np.sum(d > 1e-5))}/{d.size} above 1e-5")
    assert np.sum(d > 1e-5) == 0, msgs[-1]
    return msgs


@pytest.mark.parametrize("name", list(KINDS))
def test_bsdf_matches_golden_fixture(bpt, name):
    from tests.golden import make_golden
    stored = np.load(make_golden.HERE / "bsdf_c1_small.npz")
    t = make_golden.inputs()
    got = bpt.bsdf_eval_sample_pdf(KINDS[name], t["wo"], t["wi"], t["tint"], t["rms"], t["u"], coat=t["coat"] if name == "default" else None)
    want = {k: stored[f"{name}_{k}"] for k in got}
    for m in compare_bsdf(got, want, name):
        print(m)


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", list(KINDS))
@pytest.mark.parametrize("n", [1, 255, 1 << 18])
def test_bsdf_matches_reference(bpt, ref, name, n):
    t = bsdf_tuples(n, seed=1234 + n, with_coat=(name == "default"))
    got = bpt.bsdf_eval_sample_pdf(KINDS[name], t["wo"], t["wi"], t["tint"], t["rms"], t["u"], coat=t["coat"])
    want = ref.bsdf_eval_sample_pdf(KINDS[name], t["wo"], t["wi"], t["tint"], t["rms"], t["u"], coat=t["coat"])
    for m in compare_bsdf(got, want, f"{name}[n={n}]"):
        print(m)


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
def test_default_shading_matches_reference_at_the_full_c1_size(bpt, ref):
    """BASELINE.json configs[0] at the size bench.py --workload bsdf runs it: 2^22 DefaultShading tuples (with coat) against the
    reference's host-compiled headers, same tolerances as the smaller batches."""
    n = 1 << 22
    t = bsdf_tuples(n, seed=1234, with_coat=True)
    got = bpt.bsdf_eval_sample_pdf(KINDS["default"], t["wo"], t["wi"], t["tint"], t["rms"], t["u"], coat=t["coat"])
    want = ref.bsdf_eval_sample_pdf(KINDS["default"], t["wo"], t["wi"], t["tint"], t["rms"], t["u"], coat=t["coat"])
    for m in compare_bsdf(got, want, f"default[n={n}]"):
        print(m)


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
def test_default_shading_regression_vectors_on_gpu(bpt, ref):
    """The reference's own golden vectors (DefaultShadingTest.h:410-447), 1e-4 relative as in the reference test."""
    from tests.test_oracle_reference import regression_inputs
    t = regression_inputs(ref)
    got = bpt.bsdf_eval_sample_pdf(0, **t)
    want = ref.bsdf_eval_sample_pdf(0, **t)
    assert np.all(rel_err(got["sample_f"], want["sample_f"], 1e-12) <= 1e-4)
    assert np.all(rel_err(np.abs(got["sample_pdf"]), np.abs(want["sample_pdf"]), 1e-12) <= 1e-4)


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,kind", [("transmissive_shading", 4), ("ggx", 5)])
@pytest.mark.parametrize("n", [1, 255, 1 << 18])
def test_transmissive_bsdfs_match_reference(bpt, ref, name, kind, n):
    """TransmissiveShading (TransmissiveShading.h:22-98) and the combined reflection + transmission GGX (GGX.h:258-443)."""
    t = transmissive_tuples(n, seed=77 + n, combined_ggx=(kind == 5))
    got = bpt.bsdf_eval_sample_pdf(kind, t["wo"], t["wi"], t["tint"], t["rms"], t["u"])
    want = ref.bsdf_eval_sample_pdf(kind, t["wo"], t["wi"], t["tint"], t["rms"], t["u"])
    for m in compare_bsdf(got, want, f"{name}[n={n}]"):
        print(m)
    if n > 1000:  # the batch covers reflection and refraction, both directions through the interface
        valid = np.isin(pdf_class(got["sample_pdf"]), (1, 3))
        refracted = valid & (got["sample_dir"][:, 2] * t["wo"][:, 2] < 0)
        assert 0.2 < refracted.sum() / max(1, valid.sum()) < 0.98


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
def test_transmissive_shading_regression_vectors_on_gpu(bpt, ref):
    """The reference's own golden vectors (TransmissiveShadingTest.h:203-238), 1e-4 relative as in the reference test."""
    from tests.test_oracle_reference import TRANSMISSIVE_REGRESSION, transmissive_regression_inputs
    got = bpt.bsdf_eval_sample_pdf(4, **transmissive_regression_inputs(ref))
    assert np.all(rel_err(got["sample_f"], TRANSMISSIVE_REGRESSION[:, :3], 1e-12) <= 1e-4)
    assert np.all(rel_err(np.abs(got["sample_pdf"]), TRANSMISSIVE_REGRESSION[:, 3], 1e-12) <= 1e-4)


def test_empty_batch_is_a_no_op(bpt):
    z3 = np.zeros((0, 3), np.float32)
    out = bpt.bsdf_eval_sample_pdf(0, z3, z3, z3, z3, z3)
    assert out["eval_f"].shape == (0, 3)


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
def test_default_shading_with_path_regularization(bpt, ref):
    n = 1 << 16
    t = bsdf_tuples(n, seed=99, with_coat=True)
    rng = np.random.default_rng(5)
    m = np.zeros(n, capi.MATERIAL_DTYPE)
    m["tint"] = t["tint"]; m["roughness"] = t["rms"][:, 0]; m["metallic"] = t["rms"][:, 1]; m["specularity"] = t["rms"][:, 2]
    m["coverage"] = 1.0
    m["coat"] = (np.clip(t["coat"][:, 0], 0, 1) * 65535.0 + 0.5).astype(np.uint16)
    m["coat_roughness"] = (np.clip(t["coat"][:, 1], 0, 1) * 65535.0 + 0.5).astype(np.uint16)
    hint = np.exp(rng.uniform(np.log(1e-3), np.log(1e4), n)).astype(np.float32)
    hint[rng.random(n) < 0.2] *= -1.0  # delta dirac / MIS-disabled previous bounce
    scale = rng.integers(0, 256, (n, 4)).astype(np.float32) / np.float32(255.0)
    got = bpt.default_shading_regularized(m, hint, t["wo"], t["wi"], t["u"], scale)
    want = ref.default_shading_regularized(m, hint, t["wo"], t["wi"], t["u"], scale)
    for msg in compare_bsdf(got, want, "regularized"):
        print(msg)


# ---- RNG: integers, bit exact ---------------------------------------------------------------------

def test_sobol_matches_golden_fixture_bit_exact(bpt):
    from tests.golden import make_golden
    stored = np.load(make_golden.HERE / "bsdf_c1_small.npz")
    acc, ph, dim = make_golden.rng_inputs()
    ui, f = bpt.rng_sample4(acc, ph, dim)
    assert np.array_equal(ui, stored["sobol_ui"])
    assert np.array_equal(f, stored["sobol_f"])


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
def test_sobol_matches_reference_bit_exact(bpt, ref):
    rng = np.random.default_rng(3)
    n = 1 << 20
    acc = rng.integers(0, 1 << 32, n, dtype=np.uint32)
    acc[: 1 << 16] = np.arange(1 << 16, dtype=np.uint32)
    ph = rng.integers(0, 1 << 32, n, dtype=np.uint32)
    dim = rng.integers(0, 8 * 12, n, dtype=np.uint32)
    ui, f = bpt.rng_sample4(acc, ph, dim)
    rui, rf = ref.sobol_sample4(acc, ph, dim)
    assert np.array_equal(ui, rui)
    assert np.array_equal(f, rf)
    assert f.max() <= 1.0 and f.min() >= 0.0


# ---- lights -----------------------------------------------------------------------------------------

def make_lights(rng, n):
    lights = np.zeros(n, capi.LIGHT_DTYPE)
    kind = rng.integers(0, 3, n)
    power = rng.uniform(0.1, 50.0, (n, 3)).astype(np.float32)
    pos = rng.uniform(-5, 5, (n, 3)).astype(np.float32)
    radius = np.where(rng.random(n) < 0.15, 0.0, rng.uniform(0.01, 2.0, n)).astype(np.float32)
    direction = rng.normal(size=(n, 3)).astype(np.float32)
    direction /= np.linalg.norm(direction, axis=1, keepdims=True).astype(np.float32)
    # the core stores the spot cos angle as unorm16 and the radius as half (Scene/LightSource.cpp:105-106)
    cos_angle = (np.floor(rng.uniform(0.05, 0.99, n) * 65535 + 0.5) / 65535).astype(np.float32)
    lights["data"][:, 0:3] = power
    lights["data"][:, 3:6] = pos
    lights["data"][:, 6] = radius
    lights["data"][:, 7:10] = direction
    lights["data"][:, 10] = cos_angle
    is_dir = kind == 2
    lights["data"][is_dir, 3:6] = direction[is_dir]
    lights["data"][is_dir, 6:] = 0
    lights["flags"] = np.where(kind == 0, capi.LIGHT_SPHERE, np.where(kind == 1, capi.LIGHT_SPOT, capi.LIGHT_DIRECTIONAL))
    return lights


@pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
def test_lights_match_reference(bpt, ref):
    rng = np.random.default_rng(11)
    n = 1 << 16
    lights = make_lights(rng, n)
    position = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    u2 = rng.random((n, 2), dtype=np.float32)
    # query the pdf/evaluate functions towards the sampled direction of the oracle (on the light) for half of the elements
    q = rng.normal(size=(n, 3)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    rs, _, _ = ref.light_sample_pdf_evaluate(lights, position, u2, q)
    towards = rng.random(n) < 0.5
    ok = towards & np.all(np.isfinite(rs[:, 4:7]), axis=1)
    q[ok] = rs[ok, 4:7]
    rs, rp, rr = ref.light_sample_pdf_evaluate(lights, position, u2, q)
    gs, gp, gr = bpt.light_sample_pdf_evaluate(lights, position, u2, q)
    g = np.concatenate([gs["radiance"], gs["pdf"][:, None], gs["direction_to_light"], gs["distance"][:, None]], axis=1)

    # Measured on B200 (round 2): every one of the 65 536 sphere / spot / directional elements within 1e-5 (sample radiance and
    # PDF max 3e-7, direction 1.2e-7 absolute, distance 2.5e-7, pdf() and evaluate() bit-exact), so there is NO allowance; a PDF
    # within rounding of the 1e-6 validity threshold may change class.
    def class_mismatches(a, b):
        mismatch = pdf_class(a) != pdf_class(b)
        return int(np.sum(mismatch & ~(np.abs(np.abs(b) - 1e-6) < 1e-9)))
    assert class_mismatches(g[:, 3], rs[:, 3]) == 0
    e = rel_err(g[:, 0:4], rs[:, 0:4], 1e-6)
    print(report("light.sample radiance/pdf", e, REL_TOL)); assert np.all(e <= REL_TOL)
    d = np.abs(g[:, 4:7].astype(np.float64) - rs[:, 4:7]); d = d[np.isfinite(d)]
    print("light.sample direction max abs", d.max()); assert np.all(d <= 1e-5)
    e = rel_err(g[:, 7], rs[:, 7], 1e-6)
    print(report("light.sample distance", e, REL_TOL)); assert np.all(e <= REL_TOL)
    assert class_mismatches(gp, rp) == 0
    e = rel_err(gp, rp, 1e-6)
    print(report("light.pdf", e, REL_TOL)); assert np.all(e <= REL_TOL)
    e = rel_err(gr, rr, 1e-6)
    print(report("light.evaluate", e, REL_TOL)); assert np.all(e <= REL_TOL)
