"""Tonemapped resolve (north star item 6): the device operators against the core's own camera-effects header
(core/Bifrost/Bifrost/Math/CameraEffects.h, compiled on the host by oracle/ref_api.cpp), 1e-5 relative."""
import numpy as np
import pytest

from tests import oracle_lib
from bifrost3d_b200 import capi, scenes

needs_oracle = pytest.mark.skipif(not oracle_lib.available(), reason="oracle/_ref not built")
FILMIC_PRESETS = {"aces": (0.0, 0.53, 0.91, 0.23, 0.035), "uncharted2": (0.0, 0.55, 0.63, 0.47, 0.01), "hp": (0.0, 0.63, 0.65, 0.45, 0.0),
                  "legacy": (0.0, 0.3, 0.98, 0.22, 0.025)}  # TonemappingSettings, CameraEffects.h:28-31


def color_err(got, want):
    """Error relative to the colour's largest channel: a channel that nearly cancels in the AP1 -> sRGB matrix (saturated
    colours) carries no relative accuracy of its own in either implementation."""
    # AgX ends in pow(c, 2.2) of a value that is slightly negative for very dark colours (CameraEffects.h:249,264): NaN in
    # the reference and here alike.
    assert np.array_equal(np.isnan(got), np.isnan(want))
    scale = np.maximum(np.nanmax(np.abs(want), axis=-1, keepdims=True), 1e-4)
    return np.nan_to_num(np.abs(got.astype(np.float64) - want) / scale, nan=0.0)


def colors(n, seed):
    rng = np.random.default_rng(seed)
    c = np.exp(rng.uniform(np.log(1e-4), np.log(64.0), (n, 3))).astype(np.float32)  # 20 stops
    c[: n // 8] = np.exp(rng.uniform(np.log(1e-3), np.log(8.0), (n // 8, 1))).astype(np.float32)  # greys
    return c


@needs_oracle
def test_reference_operators_known_properties(ref):
    """Sanity of the oracle itself: the operators keep black black-ish, are monotonic on greys and stay below ~1."""
    grey = np.repeat(np.geomspace(1e-3, 32, 64, dtype=np.float32)[:, None], 3, axis=1)
    for mode in (1, 2, 3):
        out = ref.tonemap(mode, grey)
        assert np.all(np.diff(out[:, 1]) >= -1e-6), mode
        assert out.max() < 1.1 and out.min() > -0.01


@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("mode", ["linear", "filmic", "agx", "khronos_neutral"])
def test_tonemapping_operators_match_the_core_header(bpt, ref, mode):
    c = colors(1 << 16, 5)
    for exposure in (1.0, 0.35):
        got = bpt.tonemap_colors(c, mode, exposure)
        want = ref.tonemap(capi.TONEMAP[mode], c, exposure)
        e = color_err(got, want)
        print(f"{mode} exposure {exposure}: max rel err {e.max():.2e}, above 1e-5: {np.mean(e > 1e-5):.2e}")
        assert np.mean(e > 1e-5) < 1e-4 and e.max() < 1e-4


@pytest.mark.gpu
@needs_oracle
@pytest.mark.parametrize("preset", list(FILMIC_PRESETS))
def test_filmic_presets_match_the_core_header(bpt, ref, preset):
    c = colors(1 << 14, 6)
    got = bpt.tonemap_colors(c, "filmic", 1.0, FILMIC_PRESETS[preset])
    want = ref.tonemap(1, c, 1.0, FILMIC_PRESETS[preset])
    e = color_err(got, want)
    print(f"{preset}: max rel err {e.max():.2e}, above 1e-5: {np.mean(e > 1e-5):.2e}")
    assert np.mean(e > 1e-5) < 1e-4 and e.max() < 1e-4


@pytest.mark.gpu
@needs_oracle
def test_resolve_tonemapped_is_the_operator_applied_to_the_mean(bpt, ref):
    scene = scenes.cornell_box(sphere_quads=(8, 4))
    scenes.upload(bpt, scene)
    bpt.render(scene["camera"], 32, 24, 0, 4, reset=True)
    mean = bpt.resolve_float4()
    toned = bpt.resolve_tonemapped("filmic", exposure=2.0)
    want = ref.tonemap(1, mean[..., :3].reshape(-1, 3), 2.0).reshape(24, 32, 3)
    assert np.all(color_err(toned[..., :3], want) < 1e-5) and np.all(toned[..., 3] == 1.0)
    rgba8 = bpt.resolve_tonemapped("filmic", exposure=2.0, rgba8=True)
    encoded = ref.linear_to_srgb(np.clip(want, 0, None).reshape(-1)).reshape(24, 32, 3)
    expected = np.floor(np.clip(encoded, 0, 1) * 255 + 0.5)
    assert np.abs(rgba8[..., :3].astype(np.int32) - expected).max() <= 1 and np.all(rgba8[..., 3] == 255)
    with pytest.raises(capi.BptError, match="bad tonemapping settings"):
        bpt.resolve_tonemapped(7)
