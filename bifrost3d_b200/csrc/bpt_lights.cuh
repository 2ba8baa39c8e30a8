// Light sources: sampling, PDF and evaluation.
//   Sphere       extensions/OptiXRenderer/OptiXRenderer/Shading/LightSources/SphereLightImpl.h:22-103
//   Spot         .../SpotLightImpl.h:22-130
//   Directional  .../DirectionalLightImpl.h:17-52
//   dispatch     .../LightImpl.h:23-108
//   ray/sphere, ray/plane, ray/disk  .../Intersect.h:23-67
//   TBN          .../TBN.h:27-58, compute_tangents Utils.h:347-356
//   environment  .../PresampledEnvironmentLightImpl.h:17-55, latlong mapping Utils.h:288-301
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#pragma once
#include "bpt_context.h"
#include "bpt_shading.cuh"

namespace bpt {

struct LightSample {
    float3 radiance;
    Pdf pdf;
    float3 direction_to_light;
    float distance;
};

BPT_D LightSample light_sample_none() {
    LightSample s;
    s.radiance = f3(0.0f);
    s.pdf = Pdf::delta_dirac(0.0f);
    s.direction_to_light = f3(0.0f, 1.0f, 0.0f);
    s.distance = 0.0f;
    return s;
}

// Orthonormal basis (Duff et al.), Utils.h:347-356.
struct Tbn {
    float3 tangent, bitangent, normal;
    BPT_D Tbn() {}
    BPT_D explicit Tbn(float3 n) : normal(n) {
        float sign = copysignf(1.0f, n.z);
        const float a = fdiv(-1.0f, sign + n.z);
        const float b = n.x * n.y * a;
        tangent = f3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
        bitangent = f3(b, sign + n.y * n.y * a, -n.y);
    }
    // world -> local
    BPT_D float3 to_local(float3 v) const { return f3(dot(tangent, v), dot(bitangent, v), dot(normal, v)); }
    // local -> world
    BPT_D float3 to_world(float3 v) const { return v.x * tangent + v.y * bitangent + v.z * normal; }
};

namespace isect {
BPT_D float ray_sphere(float3 ray_origin, float3 ray_direction, float3 sphere_center, float sphere_radius) {
    float3 direction_to_sphere = ray_origin - sphere_center;
    float b = dot(direction_to_sphere, ray_direction);
    float radius_squared = sphere_radius * sphere_radius;
    float3 fbd = direction_to_sphere - b * ray_direction;
    float d = radius_squared - dot(fbd, fbd);
    if (d > 0.0f)
        return -b - fsqrt(d);
    return nanf("");
}
BPT_D float ray_plane(float3 ray_origin, float3 ray_direction, float3 plane_point, float3 plane_normal) {
    float d = dot(plane_normal, plane_point);
    float n_dot_o = dot(plane_normal, ray_origin);
    float n_dot_d = dot(plane_normal, ray_direction);
    return fdiv(d - n_dot_o, n_dot_d);
}
BPT_D float ray_disk(float3 ray_origin, float3 ray_direction, float3 disk_center, float3 disk_normal, float disk_radius) {
    float distance_to_plane = ray_plane(ray_origin, ray_direction, disk_center, disk_normal);
    float3 plane_intersection = ray_origin + ray_direction * distance_to_plane;
    float3 v = plane_intersection - disk_center;
    float distance_squared = dot(v, v);
    if (distance_squared <= disk_radius * disk_radius && distance_to_plane >= 0.0f)
        return distance_to_plane;
    return nanf("");
}
BPT_D float point_distance_to_plane(float3 point, float3 plane_point, float3 plane_normal) {
    return dot(plane_normal, plane_point) - dot(plane_normal, point);
}
} // namespace isect

// ---- typed views on the 48 byte Light POD -------------------------------------------------------
struct SphereLight { float3 power, position; float radius; };
struct SpotLight { float3 power, position; float radius; float3 direction; float cos_angle; };
struct DirectionalLight { float3 radiance, direction; };

BPT_D uint32_t light_type(const Light& l) { return l.flags & BPT_LIGHT_TYPE_MASK; }
BPT_D SphereLight as_sphere(const Light& l) {
    SphereLight s; s.power = f3(l.data[0], l.data[1], l.data[2]); s.position = f3(l.data[3], l.data[4], l.data[5]); s.radius = l.data[6]; return s;
}
BPT_D SpotLight as_spot(const Light& l) {
    SpotLight s; s.power = f3(l.data[0], l.data[1], l.data[2]); s.position = f3(l.data[3], l.data[4], l.data[5]); s.radius = l.data[6];
    s.direction = f3(l.data[7], l.data[8], l.data[9]); s.cos_angle = l.data[10]; return s;
}
BPT_D DirectionalLight as_directional(const Light& l) {
    DirectionalLight d; d.radiance = f3(l.data[0], l.data[1], l.data[2]); d.direction = f3(l.data[3], l.data[4], l.data[5]); return d;
}

// ---- sphere light ------------------------------------------------------------------------------
namespace sphere_light {
constexpr float small_sin_theta_squared = 0.0f;

BPT_D float surface_area(const SphereLight& l) { return 4.0f * PI_F * l.radius * l.radius; }

BPT_D bool is_delta(const SphereLight& l, float3 position) {
    float3 v = l.position - position;
    float sin_theta_squared = fdiv(l.radius * l.radius, dot(v, v));
    return sin_theta_squared <= small_sin_theta_squared;
}

BPT_D LightSample sample_radiance(const SphereLight& l, float3 position, float2 u) {
    float3 vector_to_light = l.position - position;
    float sin_theta_squared = fdiv(l.radius * l.radius, dot(vector_to_light, vector_to_light));

    LightSample s;
    if (sin_theta_squared <= small_sin_theta_squared) {
        s.direction_to_light = vector_to_light;
        s.distance = length(s.direction_to_light);
        s.direction_to_light /= s.distance;
        s.radiance = l.power / (4.0f * PI_F * s.distance * s.distance);
        s.distance -= l.radius;
        s.pdf = Pdf::delta_dirac(1.0f);
    } else {
        float cos_theta = fsqrt(1.0f - sin_theta_squared);
        DirectionalSample cone = dist::cone_sample(cos_theta, u);
        const Tbn tbn(normalize(vector_to_light));
        s.direction_to_light = tbn.to_world(cone.direction);
        s.pdf = Pdf(cone.pdf);
        s.distance = isect::ray_sphere(position, s.direction_to_light, l.position, l.radius);
        if (s.distance <= 0.0f)
            s.distance = dot(vector_to_light, s.direction_to_light);
        float inv_divisor = fdiv(1.0f, PI_F * surface_area(l));
        s.radiance = l.power * inv_divisor;
    }
    s.distance = nextafterf(s.distance, 0.0f);
    return s;
}

BPT_D Pdf pdf(const SphereLight& l, float3 lit_position, float3 direction_to_light) {
    float3 v = l.position - lit_position;
    float sin_theta_squared = fdiv(l.radius * l.radius, dot(v, v));
    if (sin_theta_squared < small_sin_theta_squared)
        return Pdf::delta_dirac(0.0f);
    float cos_theta_max = fsqrt(1.0f - sin_theta_squared);
    float cos_theta = dot(direction_to_light, normalize(v));
    float valid_direction = cos_theta >= cos_theta_max ? 1.0f : 0.0f;
    return Pdf(dist::cone_pdf(cos_theta_max) * valid_direction);
}

BPT_D float3 evaluate(const SphereLight& l, float3 position) {
    float inv_divisor = fdiv(1.0f, is_delta(l, position) ? (4.0f * PI_F) : (PI_F * surface_area(l)));
    return l.power * inv_divisor;
}
} // namespace sphere_light

// ---- spot light --------------------------------------------------------------------------------
namespace spot_light {
constexpr float min_cone_angle_to_sample = 1e-5f;

BPT_D float surface_area(const SpotLight& l) { return PI_F * pow2(l.radius); }
BPT_D bool is_delta(const SpotLight& l) { return l.radius == 0.0f; }

BPT_D Pdf pdf(const SpotLight& l, float3 lit_position, float3 direction_to_light) {
    float cos_theta = -dot(l.direction, direction_to_light);
    if (cos_theta > 0.0f && !is_delta(l)) {
        float t = isect::ray_plane(lit_position, -l.direction, l.position, l.direction);
        float cone_radius_at_intersection = fdiv(t * fsqrt(1.0f - pow2(l.cos_angle)), l.cos_angle);
        if (l.radius > cone_radius_at_intersection && l.cos_angle > min_cone_angle_to_sample)
            return Pdf(dist::cone_pdf(l.cos_angle));
        float td = isect::ray_disk(lit_position, direction_to_light, l.position, l.direction, l.radius);
        if (td >= 0.0f) {
            float area_PDF_to_solid_angle_PDF = fdiv(td * td, cos_theta);
            return Pdf(dist::disk_pdf(l.radius) * area_PDF_to_solid_angle_PDF);
        }
    }
    return Pdf::delta_dirac(0.0f);
}

BPT_D float3 evaluate(const SpotLight& l, float3 lit_position, float3 direction_to_light) {
    float cos_theta = -dot(l.direction, direction_to_light);
    float normalization = TWO_PI_F * (1.0f - l.cos_angle);
    if (is_delta(l)) {
        float3 d = l.position - lit_position;
        normalization *= dot(d, d);
    } else
        normalization *= surface_area(l) * cos_theta;
    float3 radiance = l.power / normalization;
    return (cos_theta > l.cos_angle) ? radiance : f3(0.0f);
}

BPT_D LightSample sample_radiance(const SpotLight& l, float3 lit_position, float2 u) {
    LightSample s;
    if (is_delta(l)) {
        s.direction_to_light = l.position - lit_position;
        s.distance = length(s.direction_to_light);
        s.direction_to_light /= s.distance;
        s.pdf = Pdf(1.0f); // the reference leaves this positive (SpotLightImpl.h:84)
        s.radiance = evaluate(l, lit_position, s.direction_to_light);
        return s;
    }
    const Tbn light_to_world(l.direction);

    float t = isect::ray_plane(lit_position, -l.direction, l.position, l.direction);
    float cone_radius_at_intersection = fdiv(t * fsqrt(1.0f - pow2(l.cos_angle)), l.cos_angle);
    if (l.radius > cone_radius_at_intersection && l.cos_angle > min_cone_angle_to_sample) {
        DirectionalSample cone = dist::cone_sample(l.cos_angle, u);
        s.direction_to_light = light_to_world.to_world(-cone.direction);
        s.distance = isect::ray_plane(lit_position, s.direction_to_light, l.position, l.direction);
        s.pdf = Pdf(cone.pdf);
        s.radiance = f3(0.0f);
        float3 sample_position_on_light = lit_position + s.direction_to_light * s.distance;
        float3 light_pos_to_sample_pos = sample_position_on_light - l.position;
        if (dot(light_pos_to_sample_pos, light_pos_to_sample_pos) < pow2(l.radius))
            s.radiance = evaluate(l, lit_position, s.direction_to_light);
    } else {
        float2 disk_position = dist::disk_sample(l.radius, u);
        float3 sampled_position = l.position + light_to_world.to_world(f3(disk_position.x, disk_position.y, 0.0f));
        s.direction_to_light = sampled_position - lit_position;
        s.distance = length(s.direction_to_light);
        s.direction_to_light /= s.distance;
        float cos_theta = -dot(l.direction, s.direction_to_light);
        float area_PDF_to_solid_angle_PDF = fdiv(pow2(s.distance), cos_theta);
        s.pdf = Pdf(dist::disk_pdf(l.radius) * area_PDF_to_solid_angle_PDF);
        s.radiance = evaluate(l, lit_position, s.direction_to_light);
    }
    s.distance = nextafterf(s.distance, 0.0f);
    return s;
}
} // namespace spot_light

// ---- directional light -------------------------------------------------------------------------
namespace directional_light {
BPT_D LightSample sample_radiance(const DirectionalLight& l) {
    LightSample s;
    s.radiance = l.radiance;
    s.pdf = Pdf::delta_dirac(1.0f);
    s.direction_to_light = -l.direction;
    s.distance = 1e30f;
    return s;
}
} // namespace directional_light

// ---- environment (presampled) ------------------------------------------------------------------
// Device view of the scene environment. `texels` is a latlong RGBA float image sampled bilinearly with
// wrap in u and clamp in v, matching the sampler the reference creates (EnvironmentMap.cpp). `per_pixel_pdf`
// is sampled with nearest lookup (PresampledEnvironmentMap.cpp:98-100).
struct EnvironmentView {
    const float4* texels;
    int width, height;
    const float* per_pixel_pdf;
    int pdf_width, pdf_height;
    const bpt_light_sample* samples;
    int sample_count;
    float3 tint;
    // 2-D distribution of the map for CDF inversion on the device (EnvironmentLightImpl.h:22-83): marginal CDF, pdf_height + 1
    // floats, and one conditional CDF row of pdf_width + 1 floats per PDF row (Distribution2D.h:172-207). nullptr = not uploaded.
    const float* marginal_cdf;
    const float* conditional_cdf;
};

// Utils.h:288-292, with one deviation: direction.y is clamped to [-1, 1] first. A normalised bounce direction can come
// out as y = 1 + 1 ulp at the zenith; asinf then returns NaN, which the reference hands to the texture unit (rtTex2D
// tolerates a NaN coordinate) but which a software bilinear fetch turns into NaN radiance that the fp64 sum never loses.
BPT_D float2 direction_to_latlong_texcoord(float3 direction) {
    float u = fdiv((atan2f(direction.z, direction.x) + PI_F) * 0.5f, PI_F);
    float v = fdiv(asinf(fminf(fmaxf(direction.y, -1.0f), 1.0f)) + PI_F * 0.5f, PI_F);
    return f2(u, v);
}

namespace environment_light {

BPT_D float3 fetch_bilinear(const EnvironmentView& e, float2 uv) {
    // CUDA linear filtering convention: texel centres at (i + 0.5) / N.
    float x = uv.x * e.width - 0.5f;
    float y = uv.y * e.height - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float tx = x - fx, ty = y - fy;
    int x0 = int(fx), y0 = int(fy);
    int x1 = x0 + 1, y1 = y0 + 1;
    x0 = ((x0 % e.width) + e.width) % e.width;
    x1 = ((x1 % e.width) + e.width) % e.width;
    y0 = max(0, min(e.height - 1, y0));
    y1 = max(0, min(e.height - 1, y1));
    float4 p00 = e.texels[y0 * e.width + x0], p10 = e.texels[y0 * e.width + x1];
    float4 p01 = e.texels[y1 * e.width + x0], p11 = e.texels[y1 * e.width + x1];
    float3 a = lerp(f3(p00), f3(p10), tx);
    float3 b = lerp(f3(p01), f3(p11), tx);
    return lerp(a, b, ty);
}

BPT_D float fetch_pdf_nearest(const EnvironmentView& e, float2 uv) {
    int x = int(floorf(uv.x * e.pdf_width));
    int y = int(floorf(uv.y * e.pdf_height));
    x = max(0, min(e.pdf_width - 1, x)); // RT_WRAP_CLAMP_TO_EDGE in both directions (PresampledEnvironmentMap.cpp:46-47)
    y = max(0, min(e.pdf_height - 1, y));
    return e.per_pixel_pdf[y * e.pdf_width + x];
}

BPT_D LightSample sample_radiance(const EnvironmentView& e, float2 u) {
    int index = int(u.x * e.sample_count);
    bpt_light_sample raw = e.samples[index];
    LightSample s;
    s.radiance = f3(raw.radiance[0], raw.radiance[1], raw.radiance[2]) * e.tint;
    s.pdf = Pdf(raw.pdf);
    s.direction_to_light = f3(raw.direction_to_light[0], raw.direction_to_light[1], raw.direction_to_light[2]);
    s.distance = raw.distance;
    return s;
}

// Utils.h:294-301
BPT_D float3 latlong_texcoord_to_direction(float2 uv) {
    float phi = uv.x * 2.0f * PI_F;
    float theta = uv.y * PI_F;
    float sin_theta, cos_theta, sin_phi, cos_phi;
    sincos_(theta, sin_theta, cos_theta);
    sincos_(phi, sin_phi, cos_phi);
    return -f3(sin_theta * cos_phi, cos_theta, sin_theta * sin_phi);
}

// sample_CDFs_for_uv, EnvironmentLightImpl.h:22-66: two binary searches (marginal CDF for the row, that row's conditional CDF
// for the column; the reference reads them through unfiltered integer-coordinate textures) and an inverse lerp inside the
// bracketing pair. 10 + 11 dependent loads for a 2048 x 1024 map, all but the last few from a handful of hot cache lines.
BPT_D float2 sample_cdfs_for_uv(const EnvironmentView& e, float2 random_sample) {
    float2 uv;
    int conditional_row;
    {
        int lowerbound = 0, upperbound = e.pdf_height;
        while (lowerbound + 1 != upperbound) {
            int middlebound = (lowerbound + upperbound) / 2;
            float cdf = __ldg(e.marginal_cdf + middlebound);
            if (random_sample.y < cdf) upperbound = middlebound; else lowerbound = middlebound;
        }
        conditional_row = lowerbound;
        float cdf_at_lowerbound = __ldg(e.marginal_cdf + lowerbound);
        float dv = random_sample.y - cdf_at_lowerbound;
        dv = fdiv(dv, __ldg(e.marginal_cdf + lowerbound + 1) - cdf_at_lowerbound);
        uv.y = fdiv(float(lowerbound) + dv, float(e.pdf_height));
    }
    {
        const float* __restrict__ row = e.conditional_cdf + (long long)conditional_row * (e.pdf_width + 1);
        int lowerbound = 0, upperbound = e.pdf_width;
        while (lowerbound + 1 != upperbound) {
            int middlebound = (lowerbound + upperbound) / 2;
            float cdf = __ldg(row + middlebound);
            if (random_sample.x < cdf) upperbound = middlebound; else lowerbound = middlebound;
        }
        float cdf_at_lowerbound = __ldg(row + lowerbound);
        float du = random_sample.x - cdf_at_lowerbound;
        du = fdiv(du, __ldg(row + lowerbound + 1) - cdf_at_lowerbound);
        uv.x = fdiv(float(lowerbound) + du, float(e.pdf_width));
    }
    return uv;
}

// sample_radiance(EnvironmentLight), EnvironmentLightImpl.h:69-83: importance sampling by CDF inversion.
BPT_D LightSample sample_radiance_cdf(const EnvironmentView& e, float2 u) {
    if (e.marginal_cdf == nullptr) return light_sample_none();
    float2 uv = sample_cdfs_for_uv(e, u);
    LightSample s;
    s.direction_to_light = latlong_texcoord_to_direction(uv);
    s.distance = 1e30f;
    s.radiance = fetch_bilinear(e, uv) * e.tint;
    float sin_theta = fsqrt(fmaxf(0.0f, 1.0f - s.direction_to_light.y * s.direction_to_light.y));
    float p = fdiv(fetch_pdf_nearest(e, uv), sin_theta);
    s.pdf = Pdf(sin_theta == 0.0f ? 0.0f : p);
    return s;
}

BPT_D Pdf pdf(const EnvironmentView& e, float3 direction_to_light) {
    float2 uv = direction_to_latlong_texcoord(direction_to_light);
    float sin_theta = fsqrt(fmaxf(0.0f, 1.0f - direction_to_light.y * direction_to_light.y)); // |y| = 1 + 1 ulp: see above
    float p = fdiv(fetch_pdf_nearest(e, uv), sin_theta);
    return sin_theta == 0.0f ? Pdf::delta_dirac(0.0f) : Pdf(p);
}

BPT_D float3 evaluate(const EnvironmentView& e, float3 direction_to_light) {
    float2 uv = direction_to_latlong_texcoord(direction_to_light);
    return e.tint * fetch_bilinear(e, uv);
}

} // namespace environment_light

// ---- dispatch (LightImpl.h:38-108) -------------------------------------------------------------
BPT_CALL1 LightSample light_sample_radiance(const Light& light, const EnvironmentView& env, float3 position, float2 u) {
    switch (light_type(light)) {
    case BPT_LIGHT_SPHERE: return sphere_light::sample_radiance(as_sphere(light), position, u);
    case BPT_LIGHT_DIRECTIONAL: return directional_light::sample_radiance(as_directional(light));
    case BPT_LIGHT_ENVIRONMENT: return environment_light::sample_radiance_cdf(env, u);
    case BPT_LIGHT_PRESAMPLED_ENVIRONMENT: return environment_light::sample_radiance(env, u);
    case BPT_LIGHT_SPOT: return spot_light::sample_radiance(as_spot(light), position, u);
    }
    return light_sample_none();
}

BPT_D Pdf light_pdf(const Light& light, const EnvironmentView& env, float3 lit_position, float3 direction_to_light) {
    switch (light_type(light)) {
    case BPT_LIGHT_SPHERE: return sphere_light::pdf(as_sphere(light), lit_position, direction_to_light);
    case BPT_LIGHT_DIRECTIONAL: return Pdf::delta_dirac(0.0f);
    case BPT_LIGHT_ENVIRONMENT:
    case BPT_LIGHT_PRESAMPLED_ENVIRONMENT: return environment_light::pdf(env, direction_to_light);
    case BPT_LIGHT_SPOT: return spot_light::pdf(as_spot(light), lit_position, direction_to_light);
    }
    return Pdf::invalid();
}

BPT_D float3 light_evaluate(const Light& light, const EnvironmentView& env, float3 position, float3 direction_to_light) {
    switch (light_type(light)) {
    case BPT_LIGHT_SPHERE: return sphere_light::evaluate(as_sphere(light), position);
    case BPT_LIGHT_DIRECTIONAL: return f3(0.0f);
    case BPT_LIGHT_ENVIRONMENT:
    case BPT_LIGHT_PRESAMPLED_ENVIRONMENT: return environment_light::evaluate(env, direction_to_light);
    case BPT_LIGHT_SPOT: return spot_light::evaluate(as_spot(light), position, direction_to_light);
    }
    return f3(0.0f);
}

// The context's environment as the kernels see it; CDFs only when the host uploaded them and asked for CDF inversion.
inline EnvironmentView environment_view(const Context* ctx, bool with_cdfs) {
    EnvironmentView e = {};
    e.tint = f3(ctx->env_tint[0], ctx->env_tint[1], ctx->env_tint[2]);
    if (ctx->env_width > 0) {
        e.texels = ctx->env_texels.ptr; e.width = ctx->env_width; e.height = ctx->env_height;
        e.per_pixel_pdf = ctx->env_pdf.ptr; e.pdf_width = ctx->env_pdf_width; e.pdf_height = ctx->env_pdf_height;
        e.samples = ctx->env_samples.ptr; e.sample_count = ctx->env_sample_count;
        if (with_cdfs && ctx->env_has_cdfs) { e.marginal_cdf = ctx->env_marginal_cdf.ptr; e.conditional_cdf = ctx->env_conditional_cdf.ptr; }
    }
    return e;
}

// MonteCarlo.h:20-35
BPT_D float balance_heuristic(float pdf1, float pdf2) {
    float divisor = pdf1 + pdf2;
    float result = fdiv(pdf1, divisor);
    bool result_is_invalid = isinf(divisor) || isnan(result);
    return result_is_invalid ? (pdf1 <= pdf2 ? 0.0f : 1.0f) : result;
}
BPT_D float mis_weight(Pdf pdf1, Pdf pdf2) { return balance_heuristic(pdf1.value(), pdf2.value()); }

} // namespace bpt
