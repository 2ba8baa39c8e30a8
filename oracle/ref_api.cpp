// C entry points over the REFERENCE's own host-compilable headers (kind: "reference").
//
// TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg may load the library built from this file. It is compiled against the
// staged, syntactically patched copy of the reference in baseline/_ref (see stage_reference.py)
// with the same flags that make the reference's own gtest suite pass (oracle/Makefile), so every
// number it returns comes out of the reference's code:
//   RNG            extensions/OptiXRenderer/OptiXRenderer/RNG.h:127-144,238-293
//   GGX_R          .../Shading/BSDFs/GGX.h:69-133
//   OrenNayar      .../Shading/BSDFs/OrenNayar.h:61-127
//   Burley         .../Shading/BSDFs/Burley.h:35-71
//   DefaultShading .../Shading/ShadingModels/DefaultShading.h:149-280
//   TransmissiveShading .../Shading/ShadingModels/TransmissiveShading.h:22-98, combined GGX .../BSDFs/GGX.h:258-443
//   lights         .../Shading/LightSources/{Sphere,Spot,Directional}LightImpl.h
//   rho tables     core/Bifrost/Bifrost/Assets/Shading/Fittings.h:16-76
// Array arguments are plain host pointers; vectors are packed xyz triples.
#include <OptiXRenderer/MonteCarlo.h>
#include <OptiXRenderer/RNG.h>
#include <OptiXRenderer/Shading/BSDFs/Burley.h>
#include <OptiXRenderer/Shading/BSDFs/GGX.h>
#include <OptiXRenderer/Shading/BSDFs/OrenNayar.h>
#include <OptiXRenderer/Shading/LightSources/DirectionalLightImpl.h>
#include <OptiXRenderer/Shading/LightSources/SphereLightImpl.h>
#include <OptiXRenderer/Shading/LightSources/SpotLightImpl.h>
#include <OptiXRenderer/Types.h>
// The renderer builds DefaultShading through a GPU_DEVICE-only constructor (DefaultShading.h:155-179).
// To run exactly that code path on the host we call the class' private setup functions directly.
#include <Bifrost/Assets/Shading/Fittings.h>
#include <sstream>
#include <string>
#include <vector>
#include <iostream>
#include <functional>
#define private public
#include <OptiXRenderer/Shading/ShadingModels/DefaultShading.h>
#undef private
#include <OptiXRenderer/Shading/BSDFs/GGX.h>
#include <OptiXRenderer/Shading/ShadingModels/Utils.h> // DielectricRho, which TransmissiveShading.h uses without including
#include <OptiXRenderer/Shading/ShadingModels/TransmissiveShading.h>

#include <Bifrost/Assets/Shading/Fittings.h>
#include <Bifrost/Assets/Image.h>
#include <Bifrost/Assets/InfiniteAreaLight.h>
#include <Bifrost/Assets/Texture.h>
#include <Bifrost/Math/CameraEffects.h>
#include <Bifrost/Math/OctahedralNormal.h>
#include <Bifrost/Math/RNG.h>
#include <ImageOperations/Compare.h>

#include <omp.h>
#include <cstdint>
#include <cstring>

using namespace OptiXRenderer;
using namespace optix;

static_assert(sizeof(Material) == 64, "Material must be 64 bytes");
static_assert(sizeof(Light) == 48, "Light must be 48 bytes");
static_assert(sizeof(LightSample) == 32, "LightSample must be 32 bytes");
static_assert(sizeof(BSDFSample) == 32, "BSDFSample must be 32 bytes");
static_assert(sizeof(BSDFResponse) == 16, "BSDFResponse must be 16 bytes");

namespace {

inline float3 ld3(const float* p, int64_t i) { return make_float3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
inline void st3(float* p, int64_t i, float3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

inline Material make_material(const float* tint, const float* rms, const float* coat, int64_t i) {
    Material m = {};
    m.tint = ld3(tint, i);
    m.roughness = rms[3 * i];
    m.metallic = rms[3 * i + 1];
    m.specularity = rms[3 * i + 2];
    m.coverage = 1.0f;
    if (coat != nullptr) {
        m.coat = UNorm16(coat[2 * i]);
        m.coat_roughness = UNorm16(coat[2 * i + 1]);
    }
    return m;
}

} // namespace

// The core's image / texture managers are process-wide singletons: allocated once, shared by every entry point below.
static void ensure_image_managers() {
    static bool allocated = false;
    if (!allocated) { Bifrost::Assets::Images::allocate(8u); Bifrost::Assets::Textures::allocate(4u); allocated = true; }
}

extern "C" {

int ref_sizeof(const char* name) {
    if (!strcmp(name, "Material")) return sizeof(Material);
    if (!strcmp(name, "Light")) return sizeof(Light);
    if (!strcmp(name, "LightSample")) return sizeof(LightSample);
    if (!strcmp(name, "BSDFSample")) return sizeof(BSDFSample);
    if (!strcmp(name, "BSDFResponse")) return sizeof(BSDFResponse);
    if (!strcmp(name, "MonteCarloPayload")) return sizeof(MonteCarloPayload);
    if (!strcmp(name, "VertexGeometry")) return sizeof(VertexGeometry);
    if (!strcmp(name, "DefaultShading")) return sizeof(Shading::ShadingModels::DefaultShading);
    return -1;
}

int ref_max_threads() { return omp_get_max_threads(); }

// ---------------------------------------------------------------------------------------------
// Tables (Fittings.h). Each is 32x32 floats, row-major [row = second coordinate][col = first].
// ---------------------------------------------------------------------------------------------
void ref_get_tables(float* ggx_with_fresnel_rho, float* ggx_rho, float* estimate_alpha, int* dims /*[6]*/) {
    using namespace Bifrost::Assets::Shading;
    dims[0] = Rho::GGX_with_fresnel_angle_sample_count; dims[1] = Rho::GGX_with_fresnel_roughness_sample_count;
    dims[2] = Rho::GGX_angle_sample_count; dims[3] = Rho::GGX_roughness_sample_count;
    dims[4] = Estimate_GGX_bounded_VNDF_alpha::max_PDF_sample_count; dims[5] = Estimate_GGX_bounded_VNDF_alpha::wo_dot_normal_sample_count;
    if (ggx_with_fresnel_rho) memcpy(ggx_with_fresnel_rho, Rho::GGX_with_fresnel, sizeof(float) * dims[0] * dims[1]);
    if (ggx_rho) memcpy(ggx_rho, Rho::GGX, sizeof(float) * dims[2] * dims[3]);
    if (estimate_alpha) memcpy(estimate_alpha, Estimate_GGX_bounded_VNDF_alpha::alphas, sizeof(float) * dims[4] * dims[5]);
}

// Dielectric GGX rho (Fittings.h:36-46): two 16x16x16 tables of {total_rho, reflected_rho}.
void ref_get_dielectric_tables(float* into_light_medium, float* into_dense_medium, int* dims /*[3]*/) {
    using namespace Bifrost::Assets::Shading;
    dims[0] = Rho::dielectric_GGX_angle_sample_count; dims[1] = Rho::dielectric_GGX_roughness_sample_count;
    dims[2] = Rho::dielectric_GGX_ior_i_over_o_sample_count;
    size_t bytes = sizeof(float) * 2 * dims[0] * dims[1] * dims[2];
    if (into_light_medium) memcpy(into_light_medium, Rho::dielectric_GGX_into_light_medium, bytes);
    if (into_dense_medium) memcpy(into_dense_medium, Rho::dielectric_GGX_into_dense_medium, bytes);
}

void ref_sample_dielectric_rho(int64_t n, const float* cos_theta, const float* roughness, const float* ior_i_over_o, float* out_total_reflected) {
    for (int64_t i = 0; i < n; ++i) {
        auto rho = Bifrost::Assets::Shading::Rho::sample_dielectric_GGX(cos_theta[i], roughness[i], ior_i_over_o[i]);
        out_total_reflected[2 * i] = rho.total_rho; out_total_reflected[2 * i + 1] = rho.reflected_rho;
    }
}

// ---------------------------------------------------------------------------------------------
// Tonemapping operators of the core (Math/CameraEffects.h:161-291). mode: TonemappingMode (Linear, Filmic, AgX, KhronosNeutral);
// filmic_settings: {black_clip, toe, slope, shoulder, white_clip}.
// ---------------------------------------------------------------------------------------------
void ref_tonemap(int mode, float exposure, const float* filmic_settings, int64_t n, const float* rgb_in, float* rgb_out) {
    using namespace Bifrost::Math;
    for (int64_t i = 0; i < n; ++i) {
        RGB c = RGB(rgb_in[3 * i], rgb_in[3 * i + 1], rgb_in[3 * i + 2]) * exposure;
        switch (mode) {
        case 1: c = CameraEffects::filmic(c, filmic_settings[2], filmic_settings[1], filmic_settings[3], filmic_settings[0], filmic_settings[4]); break;
        case 2: c = CameraEffects::agx(c); break;
        case 3: c = CameraEffects::khronos_neutral_tone_mapping(c); break;
        default: break;
        }
        rgb_out[3 * i] = c.r; rgb_out[3 * i + 1] = c.g; rgb_out[3 * i + 2] = c.b;
    }
}

void ref_linear_to_srgb(int64_t n, const float* in, float* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = Bifrost::Math::linear_to_sRGB(in[i]);
}

// ---------------------------------------------------------------------------------------------
// RNG. Integer outputs: parity must be bit exact.
// ---------------------------------------------------------------------------------------------
void ref_pcg2d(int64_t n, const uint32_t* x, const uint32_t* y, uint32_t* out_xy) {
    for (int64_t i = 0; i < n; ++i) {
        uint2 r = RNG::pcg2d(x[i], y[i]);
        out_xy[2 * i] = r.x; out_xy[2 * i + 1] = r.y;
    }
}

void ref_sobol_sample4(int64_t n, const uint32_t* accumulation, const uint32_t* pixel_hash, const uint32_t* dimension,
                       uint32_t* out_ui4, float* out_f4) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint4 r = RNG::PracticalScrambledSobol::sample4ui(accumulation[i], pixel_hash[i], dimension[i]);
        if (out_ui4) { out_ui4[4 * i] = r.x; out_ui4[4 * i + 1] = r.y; out_ui4[4 * i + 2] = r.z; out_ui4[4 * i + 3] = r.w; }
        if (out_f4) {
            float4 f = RNG::PracticalScrambledSobol::sample4f(accumulation[i], pixel_hash[i], dimension[i]);
            out_f4[4 * i] = f.x; out_f4[4 * i + 1] = f.y; out_f4[4 * i + 2] = f.z; out_f4[4 * i + 3] = f.w;
        }
    }
}

void ref_reverse_halton4(int n, float* out_f4) {
    for (int i = 0; i < n; ++i) {
        float4 f = RNG::ReverseHalton(i).sample4f(); // Renderer.cpp:323-336
        out_f4[4 * i] = f.x; out_f4[4 * i + 1] = f.y; out_f4[4 * i + 2] = f.z; out_f4[4 * i + 3] = f.w;
    }
}

void ref_lcg_fill(uint32_t seed, int64_t n, float* out) {
    RNG::LinearCongruential rng(seed);
    for (int64_t i = 0; i < n; ++i) out[i] = rng.sample1f();
}

void ref_sample02(int64_t n, float* out_f2) {
    for (int64_t i = 0; i < n; ++i) { float2 s = RNG::sample02((unsigned int)i); out_f2[2 * i] = s.x; out_f2[2 * i + 1] = s.y; }
}

void ref_octahedral_encode_precise(int64_t n, const float* normals, int16_t* out_s2) {
    for (int64_t i = 0; i < n; ++i) {
        auto e = Bifrost::Math::OctahedralNormal::encode_precise(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
        out_s2[2 * i] = e.encoding.x; out_s2[2 * i + 1] = e.encoding.y;
    }
}

void ref_octahedral_decode(int64_t n, const int16_t* s2, float* out_normals) {
    for (int64_t i = 0; i < n; ++i) {
        OctahedralNormal o; o.encoding.x = s2[2 * i]; o.encoding.y = s2[2 * i + 1];
        st3(out_normals, i, o.decode());
    }
}

// ---------------------------------------------------------------------------------------------
// BSDFs. kind: 0 DefaultShading, 1 GGX_R (alpha = roughness^2 clamped as alpha_from_roughness,
// specularity = tint), 2 OrenNayar (albedo = tint), 3 Burley (tint).
// rms = packed {roughness, metallic, specularity}; coat = packed {coat, coat_roughness} or null.
// threads <= 0 -> all OpenMP threads.
// ---------------------------------------------------------------------------------------------
void ref_bsdf_eval_sample_pdf(int kind, int64_t n, const float* wo, const float* wi, const float* tint, const float* rms,
                              const float* coat, const float* u,
                              float* eval_f, float* eval_pdf, float* sample_f, float* sample_pdf, float* sample_dir, int threads) {
    using namespace Shading;
    if (threads <= 0) threads = omp_get_max_threads();
    #pragma omp parallel for schedule(dynamic, 4096) num_threads(threads)
    for (int64_t i = 0; i < n; ++i) {
        float3 o = ld3(wo, i), in = ld3(wi, i), rnd = ld3(u, i);
        BSDFResponse r; BSDFSample s;
        if (kind == 0) {
            Material m = make_material(tint, rms, coat, i);
            ShadingModels::DefaultShading shading(m, o.z);
            r = shading.evaluate_with_PDF(o, in);
            s = shading.sample(o, rnd);
        } else if (kind == 1) {
            float alpha = BSDFs::GGX::alpha_from_roughness(rms[3 * i]);
            float3 specularity = ld3(tint, i);
            r = BSDFs::GGX_R::evaluate_with_PDF(alpha, specularity, o, in);
            s = BSDFs::GGX_R::sample(alpha, specularity, o, make_float2(rnd));
        } else if (kind == 2) {
            r = BSDFs::OrenNayar::evaluate_with_PDF(ld3(tint, i), rms[3 * i], o, in);
            s = BSDFs::OrenNayar::sample(ld3(tint, i), rms[3 * i], o, make_float2(rnd));
        } else if (kind == 3) {
            r = BSDFs::Burley::evaluate_with_PDF(ld3(tint, i), rms[3 * i], o, in);
            s = BSDFs::Burley::sample(ld3(tint, i), rms[3 * i], o, make_float2(rnd));
        } else if (kind == 4) {
            // rms = {roughness, cos_theta_o (signed: entering / leaving), specularity}
            Material m = make_material(tint, rms, nullptr, i);
            ShadingModels::TransmissiveShading shading(m, rms[3 * i + 1]);
            r = shading.evaluate_with_PDF(o, in);
            s = shading.sample(o, rnd);
        } else {
            // rms = {roughness, ior_i_over_o, specularity}
            float alpha = BSDFs::GGX::alpha_from_roughness(rms[3 * i]);
            r = BSDFs::GGX::evaluate_with_PDF(ld3(tint, i), alpha, rms[3 * i + 2], rms[3 * i + 1], o, in);
            s = BSDFs::GGX::sample(ld3(tint, i), alpha, rms[3 * i + 2], rms[3 * i + 1], o, rnd);
        }
        st3(eval_f, i, r.reflectance); eval_pdf[i] = r.PDF.m_PDF;
        st3(sample_f, i, s.reflectance); sample_pdf[i] = s.PDF.m_PDF; st3(sample_dir, i, s.direction);
    }
}

// DefaultShading with the path regularisation used by the renderer: the GPU-only constructor
// (DefaultShading.h:155-179) restated with the host branch of GGXMinimumRoughness::from_PDF
// (ShadingModels/Utils.h:108-122) feeding the host constructor (DefaultShading.h:149-153).
// materials: 64-byte reference Material PODs (one per element).
void ref_default_shading_regularized(int64_t n, const void* materials, const float* tint_roughness_scale /*4n or null*/,
                                     const float* max_pdf_hint, const float* wo, const float* wi, const float* u,
                                     float* eval_f, float* eval_pdf, float* sample_f, float* sample_pdf, float* sample_dir) {
    using namespace Shading;
    const Material* mats = (const Material*)materials;
    #pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < n; ++i) {
        float3 o = ld3(wo, i), in = ld3(wi, i), rnd = ld3(u, i);
        Material m = mats[i];
        float abs_cos_theta_o = o.z;
        float min_roughness = ShadingModels::GGXMinimumRoughness::from_PDF(abs_cos_theta_o, PDF(max_pdf_hint[i]));
        float4 scale = tint_roughness_scale ? make_float4(tint_roughness_scale[4 * i], tint_roughness_scale[4 * i + 1], tint_roughness_scale[4 * i + 2], tint_roughness_scale[4 * i + 3]) : make_float4(1.0f);
        // DefaultShading.h:159-176, texture lookups omitted (untextured materials).
        float coat_roughness = fmaxf(float(m.coat_roughness), min_roughness);
        float metallic = m.metallic;
        float4 tint_roughness = make_float4(m.tint, m.roughness) * scale;
        float3 tint = make_float3(tint_roughness);
        float roughness = fmaxf(tint_roughness.w, min_roughness);
        ShadingModels::DefaultShading shading(m, abs_cos_theta_o);
        float coat_rho;
        shading.setup_shading(tint, roughness, m.specularity, metallic, m.coat, coat_roughness, abs_cos_theta_o, coat_rho);
        shading.setup_sampling_probabilities(abs_cos_theta_o, coat_rho);
        BSDFResponse r = shading.evaluate_with_PDF(o, in);
        BSDFSample s = shading.sample(o, rnd);
        st3(eval_f, i, r.reflectance); eval_pdf[i] = r.PDF.m_PDF;
        st3(sample_f, i, s.reflectance); sample_pdf[i] = s.PDF.m_PDF; st3(sample_dir, i, s.direction);
    }
}

// ---------------------------------------------------------------------------------------------
// Lights. lights: 48-byte reference Light PODs, one per element (light_stride = 1) or a single
// light broadcast (light_stride = 0). Only Sphere / Spot / Directional are host compilable.
// ---------------------------------------------------------------------------------------------
void ref_light_sample_pdf_evaluate(int64_t n, const void* lights, int light_stride, const float* position, const float* u2,
                                   const float* query_direction,
                                   float* out_light_samples /*8n: radiance.xyz, pdf, dir.xyz, distance*/,
                                   float* out_pdf /*n: pdf(query_direction)*/, float* out_radiance /*3n: evaluate(query_direction)*/) {
    const Light* L = (const Light*)lights;
    for (int64_t i = 0; i < n; ++i) {
        const Light& light = L[light_stride ? i : 0];
        float3 p = ld3(position, i);
        float2 rnd = make_float2(u2[2 * i], u2[2 * i + 1]);
        float3 q = ld3(query_direction, i);
        LightSample s = LightSample::none();
        PDF pdf = PDF::invalid();
        float3 e = make_float3(0.0f);
        switch (light.get_type()) {
        case Light::Sphere:
            s = LightSources::sample_radiance(light.sphere, p, rnd);
            pdf = LightSources::pdf(light.sphere, p, q);
            e = LightSources::evaluate(light.sphere, p, q);
            break;
        case Light::Spot:
            s = LightSources::sample_radiance(light.spot, p, rnd);
            pdf = LightSources::pdf(light.spot, p, q);
            e = LightSources::evaluate(light.spot, p, q);
            break;
        case Light::Directional:
            s = LightSources::sample_radiance(light.directional, rnd);
            pdf = LightSources::pdf(light.directional, q);
            e = LightSources::evaluate(light.directional, q);
            break;
        default: break;
        }
        float* o = out_light_samples + 8 * i;
        o[0] = s.radiance.x; o[1] = s.radiance.y; o[2] = s.radiance.z; o[3] = s.PDF.m_PDF;
        o[4] = s.direction_to_light.x; o[5] = s.direction_to_light.y; o[6] = s.direction_to_light.z; o[7] = s.distance;
        out_pdf[i] = pdf.m_PDF;
        st3(out_radiance, i, e);
    }
}

// ---------------------------------------------------------------------------------------------
// Environment map: the reference's own CPU build of the sampling data (run once per environment).
//   per pixel PDF : Assets::InfiniteAreaLight ctor (InfiniteAreaLight.cpp:24-110: (r+g+b)*sin(theta) importance, 3x3 tent filter,
//                   Distribution2D CDFs) + reconstruct_solid_angle_PDF_sans_sin_theta (InfiniteAreaLight.cpp:140-157)
//   light samples : PresampledEnvironmentMap ctor (PresampledEnvironmentMap.cpp:60-96): PMJ blue-noise points in
//                   bit-reversed order through InfiniteAreaLight::sample (InfiniteAreaLight.h:90-101)
// texels: width*height RGBA float, latlong. per_pixel_pdf must hold width * max(height, 128) floats.
// Returns the sample count actually produced (1 = importance sampling disabled for a dark image) or -1 on error.
// ---------------------------------------------------------------------------------------------
int ref_environment_build(const float* texels, int width, int height, int requested_sample_count,
                          int* pdf_width, int* pdf_height, float* per_pixel_pdf, float* out_samples /*8 floats each*/, float* image_integral) {
    using namespace Bifrost;
    using namespace Bifrost::Assets;
    ensure_image_managers();
    Image image = Image::create2D("environment", PixelFormat::RGBA_Float, false, Math::Vector2ui(width, height));
    memcpy(image.get_pixels(), texels, sizeof(float) * 4 * size_t(width) * height);
    // Latlong sampler as SimpleViewer creates it: linear filtering, repeat in u, clamp in v.
    Texture latlong = Textures::create2D(image.get_ID(), MagnificationFilter::Linear, MinificationFilter::Linear, WrapMode::Repeat, WrapMode::Clamp);
    int produced = -1;
    {
        InfiniteAreaLight light(latlong);
        *pdf_width = light.get_PDF_width(); *pdf_height = light.get_PDF_height();
        if (image_integral) *image_integral = light.image_integral();
        InfiniteAreaLightUtils::reconstruct_solid_angle_PDF_sans_sin_theta(light, per_pixel_pdf);

        bool is_dark_image = light.image_integral() < 0.00001f;
        unsigned int sample_count = std::max(2u, Math::next_power_of_two((unsigned int)requested_sample_count));
        int exponent = (int)log2(sample_count);
        if (is_dark_image || requested_sample_count == 0) {
            OptiXRenderer::LightSample none = OptiXRenderer::LightSample::none();
            memcpy(out_samples, &none, sizeof(none));
            produced = 1;
        } else {
            std::vector<Math::Vector2f> rng_samples(sample_count);
            Math::RNG::fill_progressive_multijittered_bluenoise_samples(rng_samples.data(), rng_samples.data() + sample_count);
            #pragma omp parallel for schedule(dynamic, 16)
            for (int i = 0; i < int(sample_count); ++i) {
                int adjusted_sample_index = Math::reverse_bits(i) >> (32 - exponent);
                Assets::LightSample sample = light.sample(rng_samples[adjusted_sample_index]);
                float* o = out_samples + 8 * i;
                o[0] = sample.radiance.r; o[1] = sample.radiance.g; o[2] = sample.radiance.b; o[3] = sample.PDF;
                o[4] = sample.direction_to_light.x; o[5] = sample.direction_to_light.y; o[6] = sample.direction_to_light.z; o[7] = sample.distance;
            }
            produced = int(sample_count);
        }
    }
    Textures::destroy(latlong.get_ID());
    Images::destroy(image.get_ID());
    return produced;
}

// The reference's own importance sampling of an environment map at caller-supplied random points:
// InfiniteAreaLight::sample (InfiniteAreaLight.h:90-101, Distribution2D::sample_continuous) and InfiniteAreaLight::PDF of the
// sampled direction (:103-110). Also exports the distribution's CDFs (marginal: pdf_height + 1 floats, conditional:
// pdf_height x (pdf_width + 1) floats) so that the product can be fed the reference's own tables.
int ref_environment_sample(const float* texels, int width, int height, int64_t n, const float* points /*2n*/, float* out_samples /*8n*/,
                           float* out_pdf_of_direction /*n*/, float* marginal_cdf, float* conditional_cdf) {
    using namespace Bifrost;
    using namespace Bifrost::Assets;
    ensure_image_managers();
    Image image = Image::create2D("environment", PixelFormat::RGBA_Float, false, Math::Vector2ui(width, height));
    memcpy(image.get_pixels(), texels, sizeof(float) * 4 * size_t(width) * height);
    Texture latlong = Textures::create2D(image.get_ID(), MagnificationFilter::Linear, MinificationFilter::Linear, WrapMode::Repeat, WrapMode::Clamp);
    {
        InfiniteAreaLight light(latlong);
        int pw = light.get_PDF_width(), ph = light.get_PDF_height();
        if (marginal_cdf) memcpy(marginal_cdf, light.get_image_marginal_CDF(), sizeof(float) * (ph + 1));
        if (conditional_cdf) memcpy(conditional_cdf, light.get_image_conditional_CDF(), sizeof(float) * size_t(pw + 1) * ph);
        #pragma omp parallel for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            Assets::LightSample sample = light.sample(Math::Vector2f(points[2 * i], points[2 * i + 1]));
            float* o = out_samples + 8 * i;
            o[0] = sample.radiance.r; o[1] = sample.radiance.g; o[2] = sample.radiance.b; o[3] = sample.PDF;
            o[4] = sample.direction_to_light.x; o[5] = sample.direction_to_light.y; o[6] = sample.direction_to_light.z; o[7] = sample.distance;
            if (out_pdf_of_direction) out_pdf_of_direction[i] = light.PDF(sample.direction_to_light);
        }
    }
    Textures::destroy(latlong.get_ID());
    Images::destroy(image.get_ID());
    return 0;
}

// The reference's own image comparison (extensions/ImageOperations/ImageOperations/Compare.h) on two RGBA float images.
int ref_compare_images(int width, int height, const float* reference_rgba, const float* target_rgba, int mssim_support,
                       float* out_rms, float* out_ssim, float* out_mssim, float* out_rms_diff /*nullable*/, float* out_mssim_diff /*nullable*/) {
    using namespace Bifrost;
    using namespace Bifrost::Assets;
    ensure_image_managers();
    auto make = [&](const char* name, const float* pixels) {
        Image image = Image::create2D(name, PixelFormat::RGBA_Float, false, Math::Vector2ui(width, height));
        if (pixels) memcpy(image.get_pixels(), pixels, sizeof(float) * 4 * size_t(width) * height);
        return image;
    };
    Image reference = make("reference", reference_rgba), target = make("target", target_rgba);
    Image rms_diff = out_rms_diff ? make("rms diff", nullptr) : Image(), mssim_diff = out_mssim_diff ? make("mssim diff", nullptr) : Image();
    if (out_rms) *out_rms = ImageOperations::Compare::rms(reference, target, rms_diff);
    if (out_ssim) *out_ssim = ImageOperations::Compare::ssim(reference, target);
    if (out_mssim && mssim_support > 0) *out_mssim = ImageOperations::Compare::mssim(reference, target, mssim_support, mssim_diff);
    if (out_rms_diff) memcpy(out_rms_diff, rms_diff.get_pixels(), sizeof(float) * 4 * size_t(width) * height);
    if (out_mssim_diff) memcpy(out_mssim_diff, mssim_diff.get_pixels(), sizeof(float) * 4 * size_t(width) * height);
    for (Image image : { reference, target, rms_diff, mssim_diff }) if (image.exists()) Images::destroy(image.get_ID());
    return 0;
}

// MIS balance heuristic (MonteCarlo.h:20-35).
void ref_balance_heuristic(int64_t n, const float* pdf1, const float* pdf2, float* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = MonteCarlo::balance_heuristic(pdf1[i], pdf2[i]);
}

// Utils.h:372-397
void ref_offset_ray_origin(int64_t n, const float* origin, const float* direction, const float* normal, float* out) {
    for (int64_t i = 0; i < n; ++i) st3(out, i, offset_ray_origin(ld3(origin, i), ld3(direction, i), ld3(normal, i)));
}

} // extern "C"
