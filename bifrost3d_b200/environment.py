"""Host-side environment-map preparation (numpy): what the reference's host code computes once per environment and
uploads:
  * importance (r+g+b) * sin(theta) per texel, 3x3 tent filter (20/2/1 over 32) because the sampler is linear filtered
    (core/Bifrost/Bifrost/Assets/InfiniteAreaLight.cpp:24-121),
  * row prefix sums -> conditional CDF (W+1) x H and marginal CDF H+1, fp32, sequential (Math/Distribution2D.h:172-207),
  * per pixel solid angle PDF without the 1/sin(theta) factor (InfiniteAreaLight.cpp:140-157),
  * 8192 presampled lights (OptiXRenderer/PresampledEnvironmentMap.cpp:60-96) through InfiniteAreaLight::sample
    (InfiniteAreaLight.h:90-101) and Distribution2D::sample_continuous (Distribution2D.h:128-145).
The per pixel PDF and `sample()` are checked against the reference's own code in tests/test_environment.py. The point
set that is presampled is data: the reference draws PMJ blue-noise points, this module draws the (0,2) Sobol sequence
(equally stratified); a host that wants the reference's exact sample set passes its own to bpt_set_environment.
"""
import numpy as np

from . import capi

MINIMUM_PDF_HEIGHT = 128
F = np.float32


def procedural_sky(width=2048, height=1024, seed=7):
    """Gradient sky + Gaussian sun (peak 5e4, sigma 1 degree) + seeded value-noise clouds, RGBA float latlong."""
    from .scenes import value_noise
    v = (np.arange(height, dtype=F) + F(0.5)) / F(height)
    u = (np.arange(width, dtype=F) + F(0.5)) / F(width)
    uu, vv = np.meshgrid(u, v)
    phi, theta = uu * F(2 * np.pi), vv * F(np.pi)
    d = -np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], axis=-1)  # latlong_texcoord_to_direction
    up = np.clip(d[..., 1], -1, 1)
    horizon = np.exp(-np.abs(up) * 4.0)
    sky = np.stack([0.25 + 0.5 * horizon, 0.4 + 0.45 * horizon, 0.8 + 0.2 * horizon], axis=-1) * np.where(up[..., None] > 0, 1.0, 0.15)
    clouds = np.clip(value_noise(uu * 2, vv * 2, seed, 5) * 2.5 + 0.2, 0, 1)[..., None] * np.where(up[..., None] > 0.02, 1.0, 0.0)
    sky = sky * (1 - 0.6 * clouds) + 0.9 * clouds
    sun_dir = np.array([0.4, 0.6, -0.69], np.float64); sun_dir /= np.linalg.norm(sun_dir)
    cos_angle = np.clip(d @ sun_dir, -1, 1)
    angle = np.arccos(cos_angle)
    sun = 5.0e4 * np.exp(-0.5 * (angle / np.radians(1.0)) ** 2)
    rgb = sky + sun[..., None] * np.array([1.0, 0.95, 0.85])
    return np.concatenate([rgb, np.ones((height, width, 1))], axis=-1).astype(F)


def per_pixel_importance(texels):
    h, w = texels.shape[:2]
    assert h >= MINIMUM_PDF_HEIGHT, "environment maps lower than 128 rows need the resampling branch (InfiniteAreaLight.cpp:45-52)"
    sin_theta = np.sin(F(np.pi) * (np.arange(h, dtype=F) + F(0.5)) / F(h)).astype(F)
    imp = ((texels[..., 0] + texels[..., 1]) + texels[..., 2]) * sin_theta[:, None]
    imp = imp.astype(F)
    # tent filter in the reference's accumulation order: left column (low, 2*mid, up), right column, middle column (2*low, 2*up, 20*centre)
    left, right = np.roll(imp, 1, axis=1), np.roll(imp, -1, axis=1)
    def low(a): return np.concatenate([a[:1], a[:-1]], axis=0)
    def upp(a): return np.concatenate([a[1:], a[-1:]], axis=0)
    acc = np.zeros_like(imp)
    for term in (low(left), left * F(2), upp(left), low(right), right * F(2), upp(right), low(imp) * F(2), upp(imp) * F(2), imp * F(20)):
        acc = (acc + term).astype(F)
    return (acc / F(32)).astype(F)


def build_cdfs(function):
    h, w = function.shape
    conditional = np.zeros((h, w + 1), F)
    conditional[:, 1:] = np.cumsum(function, axis=1, dtype=F)
    marginal = np.zeros(h + 1, F)
    marginal[1:] = np.cumsum(conditional[:, w], dtype=F)
    integral = marginal[h] / F(w * h)
    marginal[1:h] = marginal[1:h] / marginal[h]
    marginal[h] = 1.0
    total = conditional[:, w].copy()
    ok = total > 0
    conditional[ok, 1:w] = conditional[ok, 1:w] / total[ok, None]
    conditional[:, w] = 1.0
    return marginal, conditional, integral


def solid_angle_pdf_sans_sin_theta(marginal, conditional):
    h, w = conditional.shape[0], conditional.shape[1] - 1
    scale = F(F(w * h) * (F(1.0) / (F(2.0) * F(np.pi) * F(np.pi))))
    marginal_pdf = (marginal[1:] - marginal[:-1]).astype(F)
    conditional_pdf = (conditional[:, 1:] - conditional[:, :-1]).astype(F)
    return ((marginal_pdf[:, None] * conditional_pdf) * scale).astype(F)


def _search(cdf, x, count):
    """Distribution2D::binary_search: the last index in [0, count) with cdf[index] <= x."""
    return np.clip(np.searchsorted(cdf[:count + 1], x, side="right") - 1, 0, count - 1)


def bilinear_latlong(texels, uv):
    """Assets::sample2D (Texture.cpp:114-170) for a linear-filtered texture with repeat in u and clamp in v."""
    h, w = texels.shape[:2]
    u = uv[:, 0].astype(F); v = uv[:, 1].astype(F)
    u = u - np.trunc(u); u = np.where(u < 0, u + F(1), u).astype(F)
    v = np.clip(v, F(0), np.nextafter(F(1), F(0)))
    x = (u * F(w) - F(0.5)).astype(F); y = (v * F(h) - F(0.5)).astype(F)
    x0 = np.trunc(x).astype(np.int64); y0 = np.trunc(y).astype(np.int64)
    tx = (x - x0.astype(F)).astype(F); tx = np.where(tx < 0, tx + F(1), tx).astype(F)
    ty = (y - y0.astype(F)).astype(F); ty = np.where(ty < 0, ty + F(1), ty).astype(F)
    # int() truncates towards zero, so x in (-0.5, 0) gives lower texel 0 with lerp weight x + 1: the reference's quirk
    def px(ix, iy):
        return texels[np.clip(iy, 0, h - 1), (ix + w) % w, :3]
    lower = px(x0, y0) + (px(x0 + 1, y0) - px(x0, y0)) * tx[:, None]
    upper = px(x0, y0 + 1) + (px(x0 + 1, y0 + 1) - px(x0, y0 + 1)) * tx[:, None]
    return (lower + (upper - lower) * ty[:, None]).astype(F)


def sample(texels, marginal, conditional, points):
    """InfiniteAreaLight::sample for an array of random points -> LIGHT_SAMPLE_DTYPE records."""
    h, w = conditional.shape[0], conditional.shape[1] - 1
    rx, ry = points[:, 0].astype(F), points[:, 1].astype(F)
    y = _search(marginal, ry, h)
    cdf_y = marginal[y]
    dy = (ry - cdf_y) / (marginal[y + 1] - cdf_y)
    rows = conditional[y]
    x = np.array([_search(rows[i], rx[i], w) for i in range(len(rx))]) if len(rx) < 64 else _search_rows(rows, rx, w)
    cdf_x = rows[np.arange(len(rx)), x]
    dx = (rx - cdf_x) / (rows[np.arange(len(rx)), x + 1] - cdf_x)
    marginal_pdf = marginal[y + 1] - marginal[y]
    conditional_pdf = rows[np.arange(len(rx)), x + 1] - rows[np.arange(len(rx)), x]
    pdf = ((marginal_pdf * conditional_pdf) * F(w)) * F(h)
    uv = np.stack([(x.astype(F) + dx.astype(F)) / F(w), (y.astype(F) + dy.astype(F)) / F(h)], axis=1).astype(F)
    phi, theta = uv[:, 0] * F(2.0) * F(np.pi), uv[:, 1] * F(np.pi)
    sin_theta = np.sin(theta).astype(F)
    direction = -np.stack([sin_theta * np.cos(phi).astype(F), np.cos(theta).astype(F), sin_theta * np.sin(phi).astype(F)], axis=1).astype(F)
    out = np.zeros(len(rx), capi.LIGHT_SAMPLE_DTYPE)
    out["direction_to_light"] = direction
    out["distance"] = 1e30
    out["radiance"] = bilinear_latlong(texels, uv)
    s = np.abs(np.sqrt(F(1.0) - direction[:, 1] * direction[:, 1])).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = pdf.astype(F) / (F(2.0) * F(np.pi) * F(np.pi) * s)
    out["pdf"] = np.where(s == 0, F(0), p)
    return out


def _search_rows(rows, x, count):
    # vectorised per-row binary search: last index with rows[i, index] <= x[i]
    le = rows[:, :count + 1] <= x[:, None]
    return np.clip(le.sum(axis=1) - 1, 0, count - 1)


def sobol02(n):
    """First n points of the (0,2) sequence (van der Corput + Sobol), fp32 in [0, 1)."""
    i = np.arange(n, dtype=np.uint64)
    def reverse_bits(v):
        r = np.zeros_like(v)
        for b in range(32):
            r |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(31 - b)
        return r
    x = reverse_bits(i)
    y = np.zeros_like(i)
    vdir = np.uint64(1 << 31)
    k = i.copy()
    for _ in range(32):
        y ^= np.where(k & np.uint64(1), vdir, np.uint64(0))
        k >>= np.uint64(1)
        vdir ^= vdir >> np.uint64(1)
    scale = 1.0 / 4294967296.0
    return np.minimum(np.stack([x * scale, y * scale], axis=1), np.nextafter(1.0, 0.0)).astype(np.float64)


def build_environment(texels, tint=(1.0, 1.0, 1.0), sample_count=8192):
    """-> dict for scenes (tint, texels, per_pixel_pdf, samples) as bpt_set_environment consumes them."""
    texels = np.ascontiguousarray(texels, F)
    marginal, conditional, integral = build_cdfs(per_pixel_importance(texels))
    pdf = solid_angle_pdf_sans_sin_theta(marginal, conditional)
    if integral < 0.00001 or sample_count == 0:  # PresampledEnvironmentMap.cpp:27-29: importance sampling disabled
        samples = np.zeros(1, capi.LIGHT_SAMPLE_DTYPE); samples["direction_to_light"] = (0, 1, 0); samples["pdf"] = -0.0
        pdf = np.zeros((1, 1), F)
    else:
        count = max(2, 1 << int(np.ceil(np.log2(sample_count))))
        exponent = int(np.log2(count))
        points = sobol02(count)
        order = np.array([int(format(i, f"0{exponent}b")[::-1], 2) for i in range(count)])  # bit-reversed access (:81)
        samples = sample(texels, marginal, conditional, points[order].astype(F))
    return {"tint": tuple(tint), "texels": texels, "per_pixel_pdf": pdf, "samples": samples,
            "marginal_cdf": marginal, "conditional_cdf": conditional, "integral": float(integral)}
