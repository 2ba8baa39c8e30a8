// BSDFs and the default shading model, restated for sm_100a.
//
// Arithmetic follows the reference's host-compilable headers so results agree with the
// host-compiled oracle to ~1e-5 relative (IEEE div/sqrt, no fast-math; see DESIGN.md "FP contract"):
//   PDF wrapper            extensions/OptiXRenderer/OptiXRenderer/Types.h:155-204
//   Cosine / Uniform / Cone / Disk / OrenNayar clipped-LTC / GGX (bounded) VNDF
//                          .../Distributions.h:34-98,135-181,186-260,304-461
//   GGX_R                  .../Shading/BSDFs/GGX.h:33-133
//   OrenNayar (EON)        .../Shading/BSDFs/OrenNayar.h:30-127
//   Burley                 .../Shading/BSDFs/Burley.h:29-71
//   SpecularRho / GGXMinimumRoughness   .../Shading/ShadingModels/Utils.h:27-57,104-130
//   DefaultShading         .../Shading/ShadingModels/DefaultShading.h:41-298
//   fresnel / coat helpers .../Utils.h:129-190,363-367
// The rho / alpha tables are sampled like the reference's HOST branch: float tables with the
// software bilinear of core/Bifrost/Bifrost/Math/ImageSampling.h:18-41 (not unorm16 textures).
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#pragma once
#include "bpt_math.cuh"
#include "bpt_types.h"

namespace bpt {

// ------------------------------------------------------------------------------------------------
// PDF wrapper: negative = delta dirac / MIS disabled, NaN = invalid.
// ------------------------------------------------------------------------------------------------
constexpr float MIN_VALID_PDF = 0.000001f;

struct Pdf {
    float v;
    BPT_HD Pdf() {}
    BPT_HD Pdf(float pdf) : v(pdf) {}
    BPT_HD static Pdf invalid() { return Pdf(nanf("")); }
    BPT_HD static Pdf delta_dirac(float pdf = 1.0f) { return Pdf(-pdf); }
    BPT_HD float value() const { return fabsf(v); }
    BPT_HD bool is_valid() const { return value() > MIN_VALID_PDF; }
    BPT_HD bool is_delta_dirac() const { return !(v >= 0.0f); }
    BPT_HD void disable_MIS() { if (v >= 0.0f) v = -v; }
    BPT_HD bool is_valid_and_not_delta_dirac() const { return v > MIN_VALID_PDF; }
    BPT_HD bool invalid_or_delta_dirac() const { return !(v > MIN_VALID_PDF); }
    BPT_HD bool use_for_MIS() const { return is_valid_and_not_delta_dirac(); }
};

struct BsdfResponse { float3 reflectance; Pdf pdf; };
struct BsdfSample { float3 reflectance; Pdf pdf; float3 direction; };

BPT_HD BsdfResponse bsdf_response_none() { BsdfResponse r; r.reflectance = f3(0.0f); r.pdf = Pdf(0.0f); return r; }
BPT_HD BsdfSample bsdf_sample_none() { BsdfSample s; s.reflectance = f3(0.0f); s.pdf = Pdf(0.0f); s.direction = f3(0.0f); return s; }

struct DirectionalSample { float3 direction; float pdf; };

// The two powf calls of the shading set-up (Oren-Nayar lobe selection probability: roughness^0.1, coat roughness modulation:
// x^0.25) feed cancellation-prone expressions, and CUDA's native powf is only accurate to a few ulp, so they are evaluated in
// double precision and rounded once: the correctly rounded float in all but vanishingly rare halfway cases. x^0.25 is two
// double square roots and x^y is exp(y log x) in double (error ~1e-15 relative, 7 orders of magnitude below half a float
// ulp) instead of the general pow(double, double), which is several times as long. Checked on the host over 5e7 arguments:
// both forms round to the same float as pow(double, double) EVERYWHERE, and all three differ from glibc's powf - the oracle's -
// in the same 0.08 % of arguments, where glibc's result is the one that is not correctly rounded (1 ulp).
BPT_CALL float pow_quarter_exact(float x) { return (float)sqrt(sqrt((double)x)); }
BPT_CALL float pow_exact(float x, float y) { return (float)exp((double)y * log((double)x)); }

// sin/cos of a float angle evaluated in double precision and rounded once: correctly rounded results, which
// is what the oracle's glibc sinf/cosf deliver in all but ~1e-3 of cases (CUDA's sincosf is 2 ulp). Several
// samplers square and subtract the results (sqrt(1 - x*x - y*y) in the clipped-LTC and VNDF samplers), which
// amplifies a 1 ulp difference beyond the 1e-5 parity bound; B200's FP64 rate (half of FP32) makes this cheap.
BPT_CALL void sincos_(float theta, float& s, float& c) {
    double sd, cd;
    sincos((double)theta, &sd, &cd);
    s = (float)sd; c = (float)cd;
}

// ------------------------------------------------------------------------------------------------
// Table sampling (ImageSampling.h:18-41). `pixels` is row-major [height][width].
// ------------------------------------------------------------------------------------------------
BPT_D float bilinear(const float* __restrict__ pixels, int width, int height, float u, float v) {
    u = clampf(u, 0.0f, 1.0f); // Bifrost clamp: min(max(v, lo), hi); identical for non-NaN input
    float u_coord = u * (width - 1);
    int lower_u = int(u_coord);
    int upper_u = min(lower_u + 1, width - 1);

    v = clampf(v, 0.0f, 1.0f);
    float v_coord = v * (height - 1);
    int lower_v = int(v_coord);
    int upper_v = min(lower_v + 1, height - 1);

    float u_t = u_coord - lower_u;
    const float* lower_row = pixels + lower_v * width;
    float a0 = lower_row[lower_u], a1 = lower_row[upper_u];
    float lower_pixel = a0 + (a1 - a0) * u_t;
    const float* upper_row = pixels + upper_v * width;
    float b0 = upper_row[lower_u], b1 = upper_row[upper_u];
    float upper_pixel = b0 + (b1 - b0) * u_t;

    float v_t = v_coord - lower_v;
    return lower_pixel + (upper_pixel - lower_pixel) * v_t;
}

// Pointers to the three 32x32 float tables; they may live in shared or global memory.
struct ShadingTables {
    const float* ggx_with_fresnel_rho; // [roughness][cos_theta]
    const float* ggx_rho;              // [roughness][cos_theta]
    const float* estimate_alpha;       // [cos_theta][encoded max PDF]
};
constexpr int RHO_TABLE_DIM = 32;
constexpr int TABLE_FLOATS = RHO_TABLE_DIM * RHO_TABLE_DIM;

struct SpecularRho {
    float base, full;
    BPT_D float rho(float specularity) const { return lerp(base, full, specularity); }
    BPT_D float3 rho(float3 s) const { return f3(rho(s.x), rho(s.y), rho(s.z)); }
    BPT_D float energy_loss_adjustment() const { return fdiv(1.0f, full); }
    BPT_D static SpecularRho fetch(const ShadingTables& t, float abs_cos_theta, float roughness) {
        SpecularRho r;
        r.base = bilinear(t.ggx_with_fresnel_rho, RHO_TABLE_DIM, RHO_TABLE_DIM, abs_cos_theta, roughness);
        r.full = bilinear(t.ggx_rho, RHO_TABLE_DIM, RHO_TABLE_DIM, abs_cos_theta, roughness);
        return r;
    }
};

// ------------------------------------------------------------------------------------------------
// Distributions
// ------------------------------------------------------------------------------------------------
namespace dist {

BPT_D float cone_pdf(float cos_theta_max) { return fdiv(1.0f, 2.0f * PI_F * (1.0f - cos_theta_max)); }

BPT_D DirectionalSample cone_sample(float cos_theta_max, float2 u) {
    float cos_theta = (1.0f - u.x) + u.x * cos_theta_max;
    float sin_theta = fsqrt(1.0f - cos_theta * cos_theta);
    float phi = 2.0f * PI_F * u.y;
    float sin_phi, cos_phi;
    sincos_(phi, sin_phi, cos_phi);
    DirectionalSample res;
    res.direction = f3(cos_phi * sin_theta, sin_phi * sin_theta, cos_theta);
    res.pdf = cone_pdf(cos_theta_max);
    return res;
}

BPT_D float disk_pdf(float radius) { return fdiv(1.0f, PI_F * pow2(radius)); }

BPT_D float2 disk_sample(float radius, float2 u) {
    float r = fsqrt(u.x) * radius;
    float phi = 2.0f * PI_F * u.y;
    float sin_phi, cos_phi;
    sincos_(phi, sin_phi, cos_phi);
    return f2(r * cos_phi, r * sin_phi);
}

BPT_D float uniform_hemisphere_pdf() { return 0.5f * RECIP_PI_F; }

BPT_D DirectionalSample uniform_hemisphere_sample(float2 u) {
    float z = u.x;
    float r = fsqrt(fmaxf(0.0f, 1.0f - z * z));
    float phi = TWO_PI_F * u.y;
    float sin_phi, cos_phi;
    sincos_(phi, sin_phi, cos_phi);
    DirectionalSample res;
    res.direction = f3(r * cos_phi, r * sin_phi, z);
    res.pdf = uniform_hemisphere_pdf();
    return res;
}

BPT_D float cosine_pdf(float abs_cos_theta) { return abs_cos_theta * RECIP_PI_F; }

BPT_D DirectionalSample cosine_sample(float2 u) {
    float r2 = u.x;
    float r = fsqrt(1.0f - r2);
    float z = fsqrt(r2);
    float phi = 2.0f * PI_F * u.y;
    float sin_phi, cos_phi;
    sincos_(phi, sin_phi, cos_phi);
    DirectionalSample res;
    res.direction = f3(r * cos_phi, r * sin_phi, z);
    res.pdf = z * RECIP_PI_F;
    return res;
}

// --- Clipped linearly transformed cosine for Oren-Nayar (Distributions.h:186-260) ---------------
// The 2x2 tangent basis has columns X and Y = (-X.y, X.x).
BPT_D float2 ltc_tangent_x(float3 w) {
    float2 wh = f2(w.x, w.y);
    float len_sqr = dot(wh, wh);
    return len_sqr > 0.0f ? wh / fsqrt(len_sqr) : f2(1.0f, 0.0f);
}

BPT_D void oren_nayar_ltc_coefficients(float cos_theta, float roughness, float& a, float& b, float& c, float& d) {
    a = 1.0f + roughness * (0.303392f + (-0.518982f + 0.111709f * cos_theta) * cos_theta + (-0.276266f + 0.335918f * cos_theta) * roughness);
    b = fdiv(roughness * (-1.16407f + 1.15859f * cos_theta + (0.150815f - 0.150105f * cos_theta) * roughness), cos_theta * cos_theta * cos_theta - 1.43545f);
    c = 1.0f + (0.20013f + (-0.506373f + 0.261777f * cos_theta) * cos_theta) * roughness;
    d = fdiv((0.540852f + (-1.01625f + 0.475392f * cos_theta) * cos_theta) * roughness, -1.0743f + cos_theta * (0.0725628f + cos_theta));
}

BPT_D DirectionalSample oren_nayar_cltc_sample(float roughness, float3 wo, float2 u) {
    float a, b, c, d;
    oren_nayar_ltc_coefficients(wo.z, roughness, a, b, c, d);

    float radius = fsqrt(u.x);
    float phi = 2.0f * PI_F * u.y;
    float sin_phi, cos_phi;
    sincos_(phi, sin_phi, cos_phi);
    float x = radius * cos_phi;
    float y = radius * sin_phi;

    float vz = fdiv(1.0f, fsqrt(d * d + 1.0f));
    float s = 0.5f * (1.0f + vz);
    x = -lerp(fsqrt(1.0f - y * y), x, s);
    float3 wh = f3(x, y, fsqrt(fmaxf(1.0f - (x * x + y * y), 0.0f)));
    float pdf_wh = fdiv(wh.z, PI_F * s);
    float3 wi = f3(a * wh.x + b * wh.z, c * wh.y, d * wh.x + wh.z);
    float wi_magnitude = length(wi);
    float determinant_M = c * (a - b * d);
    float pdf_wi = fdiv(pdf_wh * wi_magnitude * wi_magnitude * wi_magnitude, determinant_M);
    // wi -> local space: [X Y] * wi.xy
    float2 X = ltc_tangent_x(wo);
    float2 Y = f2(-X.y, X.x);
    float2 xy = f2(X.x * wi.x + Y.x * wi.y, X.y * wi.x + Y.y * wi.y);
    wi = normalize(f3(xy.x, xy.y, wi.z));

    DirectionalSample res;
    res.direction = wi;
    res.pdf = pdf_wi;
    return res;
}

BPT_D float oren_nayar_cltc_pdf(float roughness, float3 wo, float3 wi_shading) {
    // wi -> LTC space: transpose([X Y]) * wi.xy
    float2 X = ltc_tangent_x(wo);
    float2 Y = f2(-X.y, X.x);
    float3 wi = f3(X.x * wi_shading.x + X.y * wi_shading.y, Y.x * wi_shading.x + Y.y * wi_shading.y, wi_shading.z);

    float a, b, c, d;
    oren_nayar_ltc_coefficients(wo.z, roughness, a, b, c, d);

    float determinant_M = c * (a - b * d);
    float3 wh = f3(c * (wi.x - b * wi.z), (a - b * d) * wi.y, -c * (d * wi.x - a * wi.z));
    float wh_magnitude_squared = dot(wh, wh);
    float vz = fdiv(1.0f, fsqrt(d * d + 1.0f));
    float s = 0.5f * (1.0f + vz);
    return fdiv(fdiv(determinant_M * determinant_M, pow2(wh_magnitude_squared)) * fmaxf(wh.z, 0.0f), PI_F * s);
}

// --- GGX visible normals (Distributions.h:304-461) -----------------------------------------------
BPT_D float ggx_D(float alpha, float3 halfway) {
    float m = pow2(fdiv(halfway.x, alpha)) + pow2(fdiv(halfway.y, alpha)) + pow2(halfway.z);
    return fdiv(1.0f, PI_F * alpha * alpha * pow2(m));
}

BPT_D float ggx_lambda(float alpha, float3 w) {
    return 0.5f * (-1.0f + fsqrt(1.0f + fdiv(pow2(alpha * w.x) + pow2(alpha * w.y), pow2(w.z))));
}

// Bounded VNDF (Eto et al. 2023), isotropic alpha.
BPT_D float3 ggx_bounded_vndf_sample_reflection(float alpha, float3 wo, float2 u) {
    float3 wo_std = normalize(f3(wo.x * alpha, wo.y * alpha, wo.z));

    float phi = 2.0f * PI_F * u.y;
    float a = fminf(alpha, alpha);
    float s = 1.0f + length(f2(wo.x, wo.y));
    float a2 = a * a; float s2 = s * s;
    float k = fdiv((1.0f - a2) * s2, s2 + a2 * wo.z * wo.z);
    float b = wo.z >= 0.0f ? k * wo_std.z : wo_std.z;
    float z = fmaf(1.0f - u.x, 1.0f + b, -b);
    float sin_theta = fsqrt(fmaxf(1.0f - z * z, 0.0f));
    float sin_phi, cos_phi;
    sincos_(phi, sin_phi, cos_phi);
    float3 o_std = f3(sin_theta * cos_phi, sin_theta * sin_phi, z);

    float3 halfway_std = wo_std + o_std;
    float3 halfway = normalize(f3(halfway_std.x * alpha, halfway_std.y * alpha, halfway_std.z));
    return reflect(-wo, halfway);
}

BPT_D float ggx_bounded_vndf_reflection_pdf(float alpha, float3 wo, float3 wi) {
    float3 halfway = normalize(wo + wi);
    float ndf = ggx_D(alpha, halfway);
    float2 ao = f2(alpha * wo.x, alpha * wo.y);
    float len2 = dot(ao, ao);
    float t = fsqrt(len2 + wo.z * wo.z);
    if (wo.z >= 0.0f) {
        float min_alpha = fminf(alpha, alpha);
        float s = 1.0f + length(f2(wo.x, wo.y));
        float min_alpha_squared = min_alpha * min_alpha; float s2 = s * s;
        float k = fdiv((1.0f - min_alpha_squared) * s2, s2 + min_alpha_squared * wo.z * wo.z);
        return fdiv(ndf, 2.0f * (k * wo.z + t));
    }
    return fdiv(ndf * (t - wo.z), 2.0f * len2);
}

} // namespace dist

// ------------------------------------------------------------------------------------------------
// Fresnel and specularity helpers (Utils.h:129-190)
// ------------------------------------------------------------------------------------------------
constexpr float COAT_SPECULARITY = 0.04f;
constexpr float COAT_IOR = 1.5f;
constexpr float AIR_IOR = 1.0f;

BPT_D float3 schlick_fresnel(float3 incident_specular, float abs_cos_theta) {
    float t = pow5(1.0f - abs_cos_theta);
    return (1.0f - t) * incident_specular + t;
}

BPT_D float dielectric_specularity(float ior_o, float ior_i) { return pow2(fdiv(ior_o - ior_i, ior_o + ior_i)); }
BPT_D float dielectric_ior_from_specularity(float specularity) { return fdiv(2.0f, 1.0f - fsqrt(specularity)) - 1.0f; }
BPT_D float adjust_dielectric_specularity_to_exterior_medium(float exterior_ior, float specularity_through_air) {
    float base_ior = dielectric_ior_from_specularity(specularity_through_air);
    return dielectric_specularity(exterior_ior, base_ior);
}

// Conductor variants with the extinction coefficient fixed to zero by the only caller
// (DefaultShading.h:98-100); the general expressions are kept so the rounding is the same.
BPT_D float3 conductor_ior_from_specularity(float3 specularity, float3 ext_i) {
    float3 a = specularity - 1.0f;
    float3 b = 2.0f * specularity + 2.0f;
    float3 c = a + (specularity - 1.0f) * (ext_i * ext_i);
    float3 d = b * b - 4.0f * a * c;
    float3 sqrt_d = f3(fsqrt(d.x), fsqrt(d.y), fsqrt(d.z));
    return (-b + sqrt_d) / (2.0f * a);
}
BPT_D float3 conductor_specularity(float3 ior_o, float3 ior_i, float3 ext_i) {
    float3 ext_i_sqrd = ext_i * ext_i;
    float3 dm = ior_o - ior_i, dp = ior_o + ior_i;
    return (dm * dm + ext_i_sqrd) / (dp * dp + ext_i_sqrd);
}
BPT_D float3 adjust_conductor_specularity_to_exterior_medium(float3 exterior_ior, float3 specularity_through_air, float3 extinction) {
    float3 base_ior = conductor_ior_from_specularity(specularity_through_air, extinction);
    return conductor_specularity(exterior_ior, base_ior, extinction);
}

BPT_D float modulate_roughness_under_coat(float base_roughness, float coat_roughness) {
    float x_coat = 1.0f - AIR_IOR / COAT_IOR;
    float adjusted_roughness4 = fminf(1.0f, pow4(base_roughness) + 2.0f * x_coat * pow4(coat_roughness));
    return pow_quarter_exact(adjusted_roughness4);
}

// ------------------------------------------------------------------------------------------------
// GGX reflection (GGX.h:33-133)
// ------------------------------------------------------------------------------------------------
namespace ggx {
constexpr float MIN_ALPHA = 1e-4f;
BPT_D float alpha_from_roughness(float roughness) { return fmaxf(MIN_ALPHA, roughness * roughness); }
BPT_D float roughness_from_alpha(float alpha) { return fsqrt(alpha); }
BPT_D bool effectively_smooth(float alpha) { return alpha <= MIN_ALPHA; }
BPT_D float height_correlated_G(float alpha, float3 wo, float3 wi) {
    return fdiv(1.0f, 1.0f + dist::ggx_lambda(alpha, wo) + dist::ggx_lambda(alpha, wi));
}
} // namespace ggx

namespace ggx_r {

BPT_D float3 evaluate(float alpha, float3 specularity, float3 wo, float3 wi) {
    if (ggx::effectively_smooth(alpha))
        return f3(0.0f);
    if (wo.z * wi.z <= 0.0f)
        return f3(0.0f);

    float3 halfway = normalize(wo + wi);
    float G = ggx::height_correlated_G(alpha, wo, wi);
    float D = dist::ggx_D(alpha, halfway);
    float3 F = schlick_fresnel(specularity, dot(wo, halfway));
    return F * fdiv(D * G, 4.0f * wo.z * wi.z);
}

BPT_D Pdf pdf(float alpha, float3 wo, float3 wi) {
    if (ggx::effectively_smooth(alpha))
        return Pdf::invalid();
    return Pdf(dist::ggx_bounded_vndf_reflection_pdf(alpha, wo, wi));
}

BPT_CALL BsdfResponse evaluate_with_pdf(float alpha, float3 specularity, float3 wo, float3 wi) {
    BsdfResponse r;
    r.reflectance = evaluate(alpha, specularity, wo, wi);
    r.pdf = pdf(alpha, wo, wi);
    return r;
}

BPT_CALL BsdfSample sample(float alpha, float3 specularity, float3 wo, float2 u) {
    BsdfSample s;
    if (ggx::effectively_smooth(alpha)) {
        s.direction = f3(-wo.x, -wo.y, wo.z);
        s.pdf = Pdf::delta_dirac(1.0f);
        s.reflectance = schlick_fresnel(specularity, fabsf(wo.z)) / fabsf(s.direction.z);
        return s;
    }
    s.direction = dist::ggx_bounded_vndf_sample_reflection(alpha, wo, u);
    s.pdf = Pdf(dist::ggx_bounded_vndf_reflection_pdf(alpha, wo, s.direction));
    s.reflectance = evaluate(alpha, specularity, wo, s.direction);
    bool energyloss = s.direction.z < 0.0f;
    return energyloss ? bsdf_sample_none() : s;
}

} // namespace ggx_r

// ------------------------------------------------------------------------------------------------
// Energy-preserving Oren-Nayar (OrenNayar.h:30-127), approximate E_FON.
// ------------------------------------------------------------------------------------------------
namespace oren_nayar {

BPT_D float E_FON_approx(float cos_theta, float A, float B) {
    float mucomp = 1.0f - cos_theta;
    float GoverPi = 0.0f;
    GoverPi = mucomp * (0.0714429953f + GoverPi);
    GoverPi = mucomp * (-0.332181442f + GoverPi);
    GoverPi = mucomp * (0.491881867f + GoverPi);
    GoverPi = mucomp * (0.0571085289f + GoverPi);
    return A + B * GoverPi;
}

BPT_D float evaluate(float roughness, float3 wo, float3 wi) {
    const float constant1_FON = 0.5f - 2.0f / (3.0f * PI_F);
    const float constant2_FON = 2.0f / 3.0f - 28.0f / (15.0f * PI_F);

    float cos_theta_i = wi.z;
    float cos_theta_o = wo.z;
    float s = dot(wi, wo) - cos_theta_i * cos_theta_o;
    float s_over_t = s > 0.0f ? fdiv(s, fmaxf(cos_theta_i, cos_theta_o)) : s;
    float A = fdiv(1.0f, 1.0f + constant1_FON * roughness);
    float B = roughness * A;

    float f_single_scatter = RECIP_PI_F * A * (1.0f + roughness * s_over_t);

    float EF_o = E_FON_approx(cos_theta_o, A, B);
    float EF_i = E_FON_approx(cos_theta_i, A, B);
    float average_EF = A * (1.0f + constant2_FON * roughness);
    float multi_scatter_rho = fdiv(average_EF, 1.0f - (1.0f - average_EF));
    float f_multi_scatter = fdiv((multi_scatter_rho * RECIP_PI_F) * fabsf(1.0f - EF_o) * fabsf(1.0f - EF_i),
                                 fmaxf(1.0e-7f, 1.0f - average_EF));
    return f_single_scatter + f_multi_scatter;
}

// pow(roughness, 0.1): the only transcendental in the lobe selection probability; hoisted so that callers
// that evaluate the BSDF several times for one (roughness, wo) pay for it once.
BPT_D float uniform_lobe_roughness_factor(float roughness) { return pow_exact(roughness, 0.1f); }

BPT_D float uniform_lobe_probability(float roughness_factor, float cos_theta) {
    return roughness_factor * (0.162925f + cos_theta * (-0.372058f + (0.538233f - 0.290822f * cos_theta) * cos_theta));
}

BPT_D Pdf pdf(float roughness, float roughness_factor, float3 wo, float3 wi) {
    float uniform_probability = uniform_lobe_probability(roughness_factor, wo.z);
    float cltc_probability = 1.0f - uniform_probability;
    float cltc_PDF = dist::oren_nayar_cltc_pdf(roughness, wo, wi);
    float uniform_PDF = dist::uniform_hemisphere_pdf();
    return Pdf(uniform_probability * uniform_PDF + cltc_probability * cltc_PDF);
}

BPT_CALL BsdfResponse evaluate_with_pdf(float3 albedo, float roughness, float roughness_factor, float3 wo, float3 wi) {
    BsdfResponse r;
    r.reflectance = albedo * evaluate(roughness, wo, wi);
    r.pdf = pdf(roughness, roughness_factor, wo, wi);
    return r;
}

BPT_CALL BsdfSample sample(float3 albedo, float roughness, float roughness_factor, float3 wo, float2 u) {
    float uniform_probability = uniform_lobe_probability(roughness_factor, wo.z);
    float cltc_probability = 1.0f - uniform_probability;

    DirectionalSample ds;
    float cltc_PDF;
    if (u.x <= uniform_probability) {
        u.x = fdiv(u.x, uniform_probability);
        ds = dist::uniform_hemisphere_sample(u);
        cltc_PDF = dist::oren_nayar_cltc_pdf(roughness, wo, ds.direction);
    } else {
        u.x = fdiv(u.x - uniform_probability, cltc_probability);
        ds = dist::oren_nayar_cltc_sample(roughness, wo, u);
        cltc_PDF = ds.pdf;
    }
    float uniform_PDF = dist::uniform_hemisphere_pdf();

    BsdfSample s;
    s.direction = ds.direction;
    s.pdf = Pdf(uniform_probability * uniform_PDF + cltc_probability * cltc_PDF);
    s.reflectance = albedo * evaluate(roughness, wo, s.direction);
    return s;
}

} // namespace oren_nayar

// ------------------------------------------------------------------------------------------------
// Burley diffuse (Burley.h:29-71). Not used by the renderer; part of the C1 workload.
// ------------------------------------------------------------------------------------------------
namespace burley {

BPT_D float schlick(float abs_cos_theta) { return pow5(fmaxf(1.0f - abs_cos_theta, 0.0f)); }

BPT_D float evaluate(float roughness, float3 wo, float3 wi) {
    float3 halfway = normalize(wi + wo);
    float wi_dot_halfway = dot(wi, halfway);
    float fd90 = 0.5f + 2.0f * wi_dot_halfway * wi_dot_halfway * roughness;
    float fresnel_wo = schlick(wo.z);
    float fresnel_wi = schlick(wi.z);
    float normalizer = fdiv(1.0f, lerp(0.969371021f, 1.04337633f, roughness));
    return lerp(1.0f, fd90, fresnel_wo) * lerp(1.0f, fd90, fresnel_wi) * RECIP_PI_F * normalizer;
}

BPT_D BsdfResponse evaluate_with_pdf(float3 tint, float roughness, float3 wo, float3 wi) {
    BsdfResponse r;
    r.reflectance = tint * evaluate(roughness, wo, wi);
    r.pdf = Pdf(dist::cosine_pdf(wi.z));
    return r;
}

BPT_D BsdfSample sample(float3 tint, float roughness, float3 wo, float2 u) {
    DirectionalSample ds = dist::cosine_sample(u);
    BsdfSample s;
    s.direction = ds.direction;
    s.pdf = Pdf(ds.pdf);
    s.reflectance = tint * evaluate(roughness, wo, s.direction);
    return s;
}

} // namespace burley

// ------------------------------------------------------------------------------------------------
// Combined reflection + transmission GGX for dielectrics (GGX.h:258-443), the BSDF of ShadingModel::Transmissive.
// Out of line: only transmissive materials pay for it.
// ------------------------------------------------------------------------------------------------
BPT_D float signf(float v) { return v >= 0.0f ? 1.0f : -1.0f; }                     // Utils.h:76-78
BPT_D bool same_hemisphere(float3 wo, float3 wi) { return wo.z * wi.z >= 0.0f; }     // Utils.h:55-57

// Utils.h:192-204
BPT_D float dielectric_schlick_fresnel(float incident_specular, float abs_cos_theta, float ior_i_over_o) {
    float sin2_theta = 1.0f - pow2(abs_cos_theta);
    if (sin2_theta >= pow2(ior_i_over_o))
        return 1.0f;
    float t = pow5(1.0f - abs_cos_theta);
    return (1.0f - t) * incident_specular + t;
}

// Utils.h:242-256: refraction about the normal (0, 0, 1).
BPT_D bool refract_z(float3& refraction_direction, float3 wi, float ior_i_over_o) {
    float normal_z = 1.0f;
    float cos_theta_i = wi.z;
    if (cos_theta_i > 0.0f) {
        normal_z = -1.0f;
        cos_theta_i = -cos_theta_i;
    } else
        ior_i_over_o = 1.0f / ior_i_over_o;
    float k = 1.0f - ior_i_over_o * ior_i_over_o * (1.0f - cos_theta_i * cos_theta_i);
    refraction_direction = ior_i_over_o * wi - f3(0.0f, 0.0f, (ior_i_over_o * cos_theta_i + fsqrt(k)) * normal_z);
    return k >= 0.0f;
}

// optix::refract(r, i, n, ior) as the reference's host build uses it (GGX.h:232,415).
BPT_D bool refract_n(float3& r, float3 i, float3 n, float ior) {
    float3 nn = n;
    float negNdotV = dot(i, nn);
    float eta;
    if (negNdotV > 0.0f) {
        eta = ior;
        nn = -n;
        negNdotV = -negNdotV;
    } else
        eta = 1.0f / ior;
    const float k = 1.0f - eta * eta * (1.0f - negNdotV * negNdotV);
    if (k < 0.0f) {
        r = f3(0.0f);
        return false;
    }
    r = normalize(eta * i - (eta * negNdotV + fsqrt(k)) * nn);
    return true;
}

namespace dist {
// Sampling Visible GGX Normals with Spherical Caps (Distributions.h:347-381), isotropic alpha.
BPT_D float3 ggx_vndf_sample_halfway(float alpha, float3 wo, float2 u) {
    float3 wo_std = normalize(f3(alpha * wo.x, alpha * wo.y, wo.z));
    float phi = 2.0f * PI_F * u.y;
    float z = fmaf(1.0f - u.x, 1.0f + wo_std.z, -wo_std.z);
    float sin_theta = fsqrt(clampf(1.0f - z * z, 0.0f, 1.0f));
    float sin_phi, cos_phi;
    sincos_(phi, sin_phi, cos_phi);
    float3 c = f3(sin_theta * cos_phi, sin_theta * sin_phi, z);
    float3 wi_std = c + wo_std;
    return normalize(f3(alpha * wi_std.x, alpha * wi_std.y, fmaxf(0.0f, wi_std.z)));
}
BPT_D float ggx_vndf_pdf(float alpha, float3 wo, float3 halfway) {
    float recip_G1 = 1.0f + ggx_lambda(alpha, wo);
    float D = ggx_D(alpha, halfway);
    return dot(wo, halfway) * D / (recip_G1 * fabsf(wo.z));
}
} // namespace dist

namespace ggx_rt {

BPT_D float transmission_pdf_scale(float ior_i_over_o, float3 wo, float3 wi, float3 halfway) {
    float sqrt_denom = dot(wo, halfway) + ior_i_over_o * dot(wi, halfway);
    return pow2(ior_i_over_o / sqrt_denom) * fabsf(dot(wi, halfway));
}

BPT_D float3 compute_halfway_vector(float ior_i_over_o, float3 wo, float3 wi) {
    float3 halfway = normalize(wo + ior_i_over_o * wi);
    if (halfway.z < 0.0f)
        halfway = -halfway;
    return halfway;
}

BPT_D float normalize_reflection_probability(float reflection_probability, float3 transmission_tint) {
    float transmission_probability = 1.0f - reflection_probability;
    float scaled_transmission_probability = sum(transmission_tint) * transmission_probability;
    float scaled_reflection_probability = 3.0f * reflection_probability;
    return scaled_reflection_probability / (scaled_reflection_probability + scaled_transmission_probability);
}

BPT_D float evaluate(float alpha, float specularity, float ior_i_over_o, float3 wo, float3 wi) {
    if (ggx::effectively_smooth(alpha) || wo.z == 0.0f || wi.z == 0.0f)
        return 0.0f;
    bool entering = wo.z >= 0.0f;
    if (!entering) {
        wo.z = -wo.z;
        wi.z = -wi.z;
    }
    bool is_reflection = same_hemisphere(wo, wi);
    float halfway_ior = is_reflection ? 1.0f : ior_i_over_o;
    float3 halfway = compute_halfway_vector(halfway_ior, wo, wi);

    float G = ggx::height_correlated_G(alpha, wo, wi);
    float D = dist::ggx_D(alpha, halfway);
    float F = dielectric_schlick_fresnel(specularity, dot(wo, halfway), ior_i_over_o);

    if (is_reflection)
        return F * D * G / (4.0f * wo.z * wi.z);
    if (dot(wi, halfway) * wi.z <= 0.0f || dot(wo, halfway) * wo.z <= 0.0f)
        return 0.0f;
    float f1 = fabsf(dot(wo, halfway) * dot(wi, halfway) / (wo.z * wi.z));
    float f2 = (1.0f - F) * G * D * pow2(ior_i_over_o / (dot(wo, halfway) + ior_i_over_o * dot(wi, halfway)));
    return f1 * f2;
}

BPT_D float3 evaluate(float3 transmission_tint, float alpha, float specularity, float ior_i_over_o, float3 wo, float3 wi) {
    float f = evaluate(alpha, specularity, ior_i_over_o, wo, wi);
    bool is_transmission = signf(wo.z) != signf(wi.z);
    return f * (is_transmission ? transmission_tint : f3(1.0f));
}

BPT_D Pdf pdf(float3 transmission_tint, float alpha, float specularity, float ior_i_over_o, float3 wo, float3 wi) {
    if (ggx::effectively_smooth(alpha))
        return Pdf::invalid();
    bool entering = wo.z >= 0.0f;
    if (!entering) {
        wo.z = -wo.z;
        wi.z = -wi.z;
    }
    bool is_reflection = same_hemisphere(wo, wi);
    float halfway_ior = is_reflection ? 1.0f : ior_i_over_o;
    float3 halfway = compute_halfway_vector(halfway_ior, wo, wi);

    bool backfacing_microfacet = !is_reflection && (dot(wo, halfway) < 0.0f || dot(wi, halfway) >= 0.0f);
    if (backfacing_microfacet)
        return Pdf::invalid();

    float p = dist::ggx_vndf_pdf(alpha, wo, halfway);
    float reflection_probability = dielectric_schlick_fresnel(specularity, dot(wo, halfway), ior_i_over_o);
    float normalized_reflection_probability = normalize_reflection_probability(reflection_probability, transmission_tint);
    p *= is_reflection ? normalized_reflection_probability : (1.0f - normalized_reflection_probability);
    if (is_reflection)
        p *= 1.0f / (4.0f * dot(wo, halfway));
    else
        p *= transmission_pdf_scale(ior_i_over_o, wo, wi, halfway);
    return Pdf(p);
}

BPT_CALL BsdfResponse evaluate_with_pdf(float3 transmission_tint, float alpha, float specularity, float ior_i_over_o, float3 wo, float3 wi) {
    BsdfResponse r;
    r.reflectance = evaluate(transmission_tint, alpha, specularity, ior_i_over_o, wo, wi);
    r.pdf = pdf(transmission_tint, alpha, specularity, ior_i_over_o, wo, wi);
    return r;
}

BPT_CALL BsdfSample sample(float3 transmission_tint, float alpha, float specularity, float ior_i_over_o, float3 wo, float3 u) {
    BsdfSample s;
    s.reflectance = f3(0.0f); s.pdf = Pdf(0.0f); s.direction = f3(0.0f);

    bool entering = wo.z >= 0.0f;
    if (!entering)
        wo.z = -wo.z;

    if (ggx::effectively_smooth(alpha)) {
        float reflection_probability = dielectric_schlick_fresnel(specularity, fabsf(wo.z), ior_i_over_o);
        float normalized_reflection_probability = normalize_reflection_probability(reflection_probability, transmission_tint);
        bool is_reflection = u.z < normalized_reflection_probability;
        if (is_reflection) {
            s.pdf = Pdf::delta_dirac(normalized_reflection_probability);
            s.direction = f3(-wo.x, -wo.y, wo.z);
        } else {
            s.pdf = Pdf::delta_dirac(1.0f - normalized_reflection_probability);
            if (!refract_z(s.direction, -wo, ior_i_over_o))
                return bsdf_sample_none();
        }
        float reflectance = (is_reflection ? reflection_probability : (1.0f - reflection_probability)) / fabsf(s.direction.z);
        s.reflectance = f3(reflectance);
    } else {
        float3 halfway = dist::ggx_vndf_sample_halfway(alpha, wo, f2(u.x, u.y));
        float p = dist::ggx_vndf_pdf(alpha, wo, halfway);

        float reflection_probability = dielectric_schlick_fresnel(specularity, dot(wo, halfway), ior_i_over_o);
        float normalized_reflection_probability = normalize_reflection_probability(reflection_probability, transmission_tint);
        bool is_reflection = u.z < normalized_reflection_probability;

        if (is_reflection) {
            s.direction = reflect(-wo, halfway);
            p *= normalized_reflection_probability / (4.0f * dot(wo, halfway));
        } else {
            if (!refract_n(s.direction, -wo, halfway, ior_i_over_o))
                return bsdf_sample_none();
            p *= 1.0f - normalized_reflection_probability;
            p *= transmission_pdf_scale(ior_i_over_o, wo, s.direction, halfway);
        }
        s.pdf = Pdf(p);

        bool energyloss = is_reflection ? s.direction.z < 0.0f : s.direction.z >= 0.0f;
        if (energyloss)
            return bsdf_sample_none();

        float f = evaluate(alpha, specularity, ior_i_over_o, wo, s.direction);
        s.reflectance = f3(f);
    }

    bool is_transmission = signf(wo.z) != signf(s.direction.z);
    if (is_transmission)
        s.reflectance *= transmission_tint;
    if (!entering)
        s.direction.z = -s.direction.z;
    return s;
}

} // namespace ggx_rt

// Dielectric rho tables (Fittings.h:27-46, DielectricGGXRho.cpp:1086-1110): two 16 x 16 x 16 tables of
// {total rho, reflected rho}, sampled trilinearly like Math::ImageSampling::trilinear (ImageSampling.h:43-60).
constexpr int DIELECTRIC_DIM = 16;
constexpr int DIELECTRIC_TABLE_FLOAT2S = DIELECTRIC_DIM * DIELECTRIC_DIM * DIELECTRIC_DIM;

BPT_D float2 bilinear2(const float2* __restrict__ pixels, int width, int height, float u, float v) {
    u = clampf(u, 0.0f, 1.0f);
    float u_coord = u * (width - 1);
    int lower_u = int(u_coord);
    int upper_u = min(lower_u + 1, width - 1);
    v = clampf(v, 0.0f, 1.0f);
    float v_coord = v * (height - 1);
    int lower_v = int(v_coord);
    int upper_v = min(lower_v + 1, height - 1);
    float u_t = u_coord - lower_u;
    const float2* lower_row = pixels + lower_v * width;
    float2 a0 = lower_row[lower_u], a1 = lower_row[upper_u];
    float2 lower_pixel = make_float2(a0.x + (a1.x - a0.x) * u_t, a0.y + (a1.y - a0.y) * u_t);
    const float2* upper_row = pixels + upper_v * width;
    float2 b0 = upper_row[lower_u], b1 = upper_row[upper_u];
    float2 upper_pixel = make_float2(b0.x + (b1.x - b0.x) * u_t, b0.y + (b1.y - b0.y) * u_t);
    float v_t = v_coord - lower_v;
    return make_float2(lower_pixel.x + (upper_pixel.x - lower_pixel.x) * v_t, lower_pixel.y + (upper_pixel.y - lower_pixel.y) * v_t);
}

// tables: [into_light 16^3 | into_dense 16^3] float2. Returns {total_rho, reflected_rho}.
BPT_CALL float2 dielectric_rho_fetch(const float2* __restrict__ tables, float abs_cos_theta, float roughness, float ior_i_over_o) {
    const float2* table = tables;
    float w;
    if (ior_i_over_o < 1.0f)
        w = (ior_i_over_o - 0.331492f) / 0.457982f;
    else {
        w = (ior_i_over_o - 1.26667f) / 1.75f;
        table = tables + DIELECTRIC_TABLE_FLOAT2S;
    }
    w = clampf(w, 0.0f, 1.0f);
    float w_coord = w * (DIELECTRIC_DIM - 1);
    int lower_w = int(w_coord);
    int upper_w = min(lower_w + 1, DIELECTRIC_DIM - 1);
    float2 lower_pixel = bilinear2(table + lower_w * DIELECTRIC_DIM * DIELECTRIC_DIM, DIELECTRIC_DIM, DIELECTRIC_DIM, abs_cos_theta, roughness);
    float2 upper_pixel = bilinear2(table + upper_w * DIELECTRIC_DIM * DIELECTRIC_DIM, DIELECTRIC_DIM, DIELECTRIC_DIM, abs_cos_theta, roughness);
    float w_t = w_coord - lower_w;
    return make_float2(lower_pixel.x + (upper_pixel.x - lower_pixel.x) * w_t, lower_pixel.y + (upper_pixel.y - lower_pixel.y) * w_t);
}

// ------------------------------------------------------------------------------------------------
// Path regularisation: smallest GGX roughness whose bounded-VNDF peak PDF stays below a hint
// (ShadingModels/Utils.h:104-130, host branch; EstimateGGXBoundedVNDFAlpha.cpp encode_PDF / estimate_alpha).
// ------------------------------------------------------------------------------------------------
BPT_D float ggx_min_roughness_from_pdf(const ShadingTables& t, float abs_cos_theta, Pdf max_pdf) {
    if (max_pdf.is_delta_dirac())
        return 0.0f;
    float pdf = max_pdf.value();
    float non_linear_PDF = fdiv(pdf, 1.0f + pdf);
    float encoded_PDF = fdiv(non_linear_PDF - 0.13f, 0.87f);
    if (isnan(encoded_PDF))
        encoded_PDF = 1.0f;
    float min_alpha = bilinear(t.estimate_alpha, RHO_TABLE_DIM, RHO_TABLE_DIM, encoded_PDF, abs_cos_theta);
    return ggx::roughness_from_alpha(min_alpha);
}

// ------------------------------------------------------------------------------------------------
// Default shading: EON diffuse + GGX specular (+ optional GGX coat). DefaultShading.h:41-298
// ------------------------------------------------------------------------------------------------
struct DefaultShading {
    float3 diffuse_tint;
    float roughness;
    float3 specularity;
    float specular_scale;
    float coat_scale;
    float coat_alpha;
    float diffuse_roughness_factor; // pow(roughness, 0.1), hoisted out of the Oren-Nayar PDF
    unsigned short specular_probability_q; // quantised to 16 bit exactly like the reference
    unsigned short coat_probability_q;
    bool diffuse_only; // ShadingModel::Diffuse (DiffuseShading.h:21-49): the Oren-Nayar lobe alone, no regularisation

    BPT_D static float compute_specular_properties(const ShadingTables& t, float roughness, float specularity, float scale, float abs_cos_theta_o,
                                                   float& alpha, float& reflection_scale, float& transmission_scale) {
        alpha = ggx::alpha_from_roughness(roughness);
        SpecularRho rho_computation = SpecularRho::fetch(t, abs_cos_theta_o, roughness);
        reflection_scale = scale * rho_computation.energy_loss_adjustment();
        float specular_rho = rho_computation.rho(specularity) * reflection_scale;
        transmission_scale = 1.0f - specular_rho;
        return specular_rho;
    }

    BPT_D void setup_shading(const ShadingTables& t, float3 tint, float roughness_in, float dielectric_spec, float metallic,
                             float coat_scale_in, float coat_roughness, float cos_theta_o, float& coat_rho) {
        float abs_cos_theta_o = fabsf(cos_theta_o);

        roughness = roughness_in;
        float3 conductor_spec = tint;

        if (coat_scale_in > 0.0f) {
            float coat_modulated_roughness = modulate_roughness_under_coat(roughness_in, coat_roughness);
            roughness = lerp(roughness_in, coat_modulated_roughness, coat_scale_in);

            if (dielectric_spec < 1.0f) {
                float coated = adjust_dielectric_specularity_to_exterior_medium(COAT_IOR, dielectric_spec);
                dielectric_spec = lerp(dielectric_spec, coated, coat_scale_in);
            }

            if (metallic > 0.0f) {
                float3 coated = adjust_conductor_specularity_to_exterior_medium(f3(COAT_IOR), conductor_spec, f3(0.0f));
                conductor_spec = lerp(conductor_spec, coated, coat_scale_in);
                conductor_spec.x = isnan(conductor_spec.x) ? 1.0f : conductor_spec.x;
                conductor_spec.y = isnan(conductor_spec.y) ? 1.0f : conductor_spec.y;
                conductor_spec.z = isnan(conductor_spec.z) ? 1.0f : conductor_spec.z;
            }
        }

        float specular_alpha, dielectric_specular_transmission;
        compute_specular_properties(t, roughness, dielectric_spec, 1.0f, abs_cos_theta_o,
                                    specular_alpha, specular_scale, dielectric_specular_transmission);
        float3 dielectric_tint = tint * dielectric_specular_transmission;

        specularity = lerp(f3(dielectric_spec), conductor_spec, metallic);
        diffuse_tint = dielectric_tint * (1.0f - metallic);

        if (coat_scale_in > 0.0f) {
            float coat_transmission;
            coat_rho = compute_specular_properties(t, coat_roughness, COAT_SPECULARITY, coat_scale_in, abs_cos_theta_o,
                                                   coat_alpha, coat_scale, coat_transmission);
            specular_scale *= coat_transmission;
            diffuse_tint *= coat_transmission;
        } else {
            coat_rho = 0.0f;
            coat_scale = 0.0f;
            coat_alpha = 0.0f;
        }
    }

    BPT_D float3 specular_rho(const ShadingTables& t, float abs_cos_theta) const {
        return SpecularRho::fetch(t, abs_cos_theta, roughness).rho(specularity) * specular_scale;
    }
    BPT_D float coat_rho_at(const ShadingTables& t, float abs_cos_theta) const {
        float coat_roughness = ggx::roughness_from_alpha(coat_alpha);
        return SpecularRho::fetch(t, abs_cos_theta, coat_roughness).rho(COAT_SPECULARITY) * coat_scale;
    }
    BPT_D float3 rho(const ShadingTables& t, float abs_cos_theta) const {
        float3 r = diffuse_tint + specular_rho(t, abs_cos_theta);
        if (coat_scale > 0.0f)
            r = r + coat_rho_at(t, abs_cos_theta);
        return r;
    }

    BPT_D void setup_sampling_probabilities(const ShadingTables& t, float abs_cos_theta_o, float coat_rho) {
        const float USHORT_MAX_F = 65535.0f;
        float diffuse_rho_sum = sum(diffuse_tint);
        float specular_rho_sum = sum(specular_rho(t, abs_cos_theta_o));
        float coat_rho_sum = 3.0f * coat_rho;
        float recip_total_rho = fdiv(1.0f, diffuse_rho_sum + specular_rho_sum + coat_rho_sum);

        float specular_probability = specular_rho_sum * recip_total_rho;
        specular_probability_q = (unsigned short)(specular_probability * USHORT_MAX_F + 0.5f);
        float coat_probability = coat_rho_sum * recip_total_rho;
        coat_probability_q = (unsigned short)(coat_probability * USHORT_MAX_F + 0.5f);
    }

    // Host constructor equivalent (DefaultShading.h:149-153).
    BPT_D static DefaultShading create(const ShadingTables& t, float3 tint, float roughness, float specularity, float metallic,
                                       float coat, float coat_roughness, float abs_cos_theta_o) {
        DefaultShading s;
        float coat_rho;
        s.setup_shading(t, tint, roughness, specularity, metallic, coat, coat_roughness, abs_cos_theta_o, coat_rho);
        s.setup_sampling_probabilities(t, abs_cos_theta_o, coat_rho);
        s.diffuse_roughness_factor = oren_nayar::uniform_lobe_roughness_factor(s.roughness);
        s.diffuse_only = false;
        return s;
    }

    // Renderer constructor (DefaultShading.h:155-179) for untextured materials: per-vertex tint/roughness
    // scale, roughness floor from the previous bounce's BSDF PDF.
    BPT_D static DefaultShading create_regularized(const ShadingTables& t, const Material& m, float4 tint_and_roughness_scale,
                                                   float abs_cos_theta_o, Pdf max_pdf_hint) {
        float min_roughness = ggx_min_roughness_from_pdf(t, abs_cos_theta_o, max_pdf_hint);
        float coat_roughness = fmaxf(unorm16_to_float(m.coat_roughness), min_roughness);
        float metallic = m.metallic;
        float3 tint = f3(m.tint[0] * tint_and_roughness_scale.x, m.tint[1] * tint_and_roughness_scale.y, m.tint[2] * tint_and_roughness_scale.z);
        float roughness = fmaxf(m.roughness * tint_and_roughness_scale.w, min_roughness);
        return create(t, tint, roughness, m.specularity, metallic, unorm16_to_float(m.coat), coat_roughness, abs_cos_theta_o);
    }

    // DiffuseMaterialCreator::create, MonteCarlo.cu:250-255.
    BPT_D static DefaultShading create_diffuse(const Material& m, float4 tint_and_roughness_scale) {
        DefaultShading s;
        s.diffuse_tint = f3(m.tint[0] * tint_and_roughness_scale.x, m.tint[1] * tint_and_roughness_scale.y, m.tint[2] * tint_and_roughness_scale.z);
        s.roughness = m.roughness * tint_and_roughness_scale.w;
        s.specularity = f3(0.0f); s.specular_scale = 0.0f; s.coat_scale = 0.0f; s.coat_alpha = 0.0f;
        s.specular_probability_q = 0; s.coat_probability_q = 0;
        s.diffuse_roughness_factor = oren_nayar::uniform_lobe_roughness_factor(s.roughness);
        s.diffuse_only = true;
        return s;
    }

    BPT_D float get_specular_alpha() const { return ggx::alpha_from_roughness(roughness); }
    BPT_D float get_diffuse_probability() const { return 1.0f - fdiv(float(specular_probability_q + coat_probability_q), 65535.0f); }
    BPT_D float get_specular_probability() const { return fdiv(float(specular_probability_q), 65535.0f); }
    BPT_D float get_coat_probability() const { return fdiv(float(coat_probability_q), 65535.0f); }

    BPT_D BsdfResponse evaluate_with_pdf(float3 wo, float3 wi) const {
        if (wo.z < 0.000001f || wi.z < 0.000001f)
            return bsdf_response_none();
        if (diffuse_only)
            return oren_nayar::evaluate_with_pdf(diffuse_tint, roughness, diffuse_roughness_factor, wo, wi);

        BsdfResponse diffuse_response = oren_nayar::evaluate_with_pdf(diffuse_tint, roughness, diffuse_roughness_factor, wo, wi);
        BsdfResponse specular_response = ggx_r::evaluate_with_pdf(get_specular_alpha(), specularity, wo, wi);
        specular_response.reflectance *= specular_scale;

        BsdfResponse response;
        response.reflectance = diffuse_response.reflectance + specular_response.reflectance;

        float diffuse_probability = get_diffuse_probability();
        float specular_probability = get_specular_probability();
        response.pdf = Pdf(diffuse_response.pdf.v * diffuse_probability + specular_response.pdf.v * specular_probability);

        if (coat_scale > 0.0f) {
            float coat_probability = get_coat_probability();
            BsdfResponse coat_response = ggx_r::evaluate_with_pdf(coat_alpha, f3(COAT_SPECULARITY), wo, wi);
            response.reflectance += coat_scale * coat_response.reflectance;
            response.pdf.v += coat_response.pdf.v * coat_probability;
        }
        return response;
    }

    BPT_D BsdfSample sample(float3 wo, float3 u) const {
        if (wo.z < 0.000001f)
            return bsdf_sample_none();
        if (diffuse_only)
            return oren_nayar::sample(diffuse_tint, roughness, diffuse_roughness_factor, wo, f2(u.x, u.y));

        float specular_probability = get_specular_probability();
        float coat_probability = get_coat_probability();
        float diffuse_probability = 1.0f - coat_probability - specular_probability;

        bool sample_coat = u.z < coat_probability;
        bool sample_specular = !sample_coat && u.z < (coat_probability + specular_probability);
        bool sample_diffuse = !sample_coat && !sample_specular;

        BsdfSample s;
        if (sample_diffuse) {
            s = oren_nayar::sample(diffuse_tint, roughness, diffuse_roughness_factor, wo, f2(u.x, u.y));
            s.pdf.v *= diffuse_probability;
        } else if (sample_specular) {
            s = ggx_r::sample(get_specular_alpha(), specularity, wo, f2(u.x, u.y));
            s.reflectance *= specular_scale;
            s.pdf.v *= specular_probability;
        } else {
            s = ggx_r::sample(coat_alpha, f3(COAT_SPECULARITY), wo, f2(u.x, u.y));
            s.reflectance *= coat_scale;
            s.pdf.v *= coat_probability;
        }

        if (s.pdf.invalid_or_delta_dirac())
            return s;

        if (!sample_diffuse) {
            BsdfResponse r = oren_nayar::evaluate_with_pdf(diffuse_tint, roughness, diffuse_roughness_factor, wo, s.direction);
            if (r.pdf.is_valid_and_not_delta_dirac()) {
                s.reflectance += r.reflectance;
                s.pdf.v += r.pdf.v * diffuse_probability;
            }
        }
        if (!sample_specular) {
            BsdfResponse r = ggx_r::evaluate_with_pdf(get_specular_alpha(), specularity, wo, s.direction);
            if (r.pdf.is_valid_and_not_delta_dirac()) {
                s.reflectance += r.reflectance * specular_scale;
                s.pdf.v += r.pdf.v * specular_probability;
            }
        }
        if (!sample_coat && coat_scale > 0.0f) {
            BsdfResponse r = ggx_r::evaluate_with_pdf(coat_alpha, f3(COAT_SPECULARITY), wo, s.direction);
            if (r.pdf.is_valid_and_not_delta_dirac()) {
                s.reflectance += coat_scale * r.reflectance;
                s.pdf.v += r.pdf.v * coat_probability;
            }
        }
        return s;
    }
};

// ------------------------------------------------------------------------------------------------
// Transmissive shading: rough dielectric interface with tinted transmission. TransmissiveShading.h:22-98
// ------------------------------------------------------------------------------------------------
struct TransmissiveShading {
    float3 transmission_tint;
    float specularity;
    float ggx_alpha;
    float ior_i_over_o;
    float energy_loss_adjustment;

    BPT_D static TransmissiveShading create(const float2* __restrict__ dielectric_tables, float3 tint, float roughness, float specularity_, float cos_theta_o) {
        TransmissiveShading s;
        s.transmission_tint = tint;
        s.specularity = specularity_;
        s.ggx_alpha = ggx::alpha_from_roughness(roughness);
        float medium_ior = dielectric_ior_from_specularity(specularity_);
        bool entering = cos_theta_o >= 0.0f;
        float ior_o = entering ? AIR_IOR : medium_ior;
        float ior_i = entering ? medium_ior : AIR_IOR;
        s.ior_i_over_o = ior_i / ior_o;
        float rho = dielectric_rho_fetch(dielectric_tables, fabsf(cos_theta_o), roughness, s.ior_i_over_o).x;
        s.energy_loss_adjustment = 1.0f / rho;
        return s;
    }

    // Renderer constructor (TransmissiveShading.h:51-66) for untextured materials.
    BPT_D static TransmissiveShading create_regularized(const ShadingTables& t, const float2* __restrict__ dielectric_tables, const Material& m,
                                                        float4 tint_and_roughness_scale, float cos_theta_o, Pdf max_pdf_hint) {
        float min_roughness = ggx_min_roughness_from_pdf(t, fabsf(cos_theta_o), max_pdf_hint);
        float3 tint = f3(m.tint[0] * tint_and_roughness_scale.x, m.tint[1] * tint_and_roughness_scale.y, m.tint[2] * tint_and_roughness_scale.z);
        float roughness = fmaxf(m.roughness * tint_and_roughness_scale.w, min_roughness);
        return create(dielectric_tables, tint, roughness, m.specularity, cos_theta_o);
    }

    BPT_D BsdfResponse evaluate_with_pdf(float3 wo, float3 wi) const {
        if (wo.z < 0.000001f)
            return bsdf_response_none();
        BsdfResponse r = ggx_rt::evaluate_with_pdf(transmission_tint, ggx_alpha, specularity, ior_i_over_o, wo, wi);
        r.reflectance *= energy_loss_adjustment;
        return r;
    }

    BPT_D BsdfSample sample(float3 wo, float3 u) const {
        if (wo.z < 0.000001f)
            return bsdf_sample_none();
        BsdfSample s = ggx_rt::sample(transmission_tint, ggx_alpha, specularity, ior_i_over_o, wo, u);
        s.reflectance *= energy_loss_adjustment;
        return s;
    }

    BPT_D float3 rho(const float2* __restrict__ dielectric_tables, float abs_cos_theta_o) const {
        float roughness = ggx::roughness_from_alpha(ggx_alpha);
        float2 r = dielectric_rho_fetch(dielectric_tables, abs_cos_theta_o, roughness, ior_i_over_o);
        float reflection = r.y / r.x;
        return f3(reflection) + (1.0f - reflection) * transmission_tint;
    }
};

} // namespace bpt
