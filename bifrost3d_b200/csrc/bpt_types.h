// Device-side view of the PODs declared in include/bpt_c_api.h.
#pragma once
#include "../../include/bpt_c_api.h"
#include "bpt_math.cuh"

namespace bpt {

typedef bpt_material Material;
typedef bpt_light Light;

static_assert(sizeof(Material) == 64, "Material must match Types.h:353-416 (64 bytes)");
static_assert(sizeof(Light) == 48, "Light must match Types.h:290-312 (48 bytes)");
static_assert(sizeof(bpt_light_sample) == 32, "LightSample must match Types.h:210-222 (32 bytes)");

enum MaterialFlags : uint16_t { MATERIAL_THIN_WALLED = 1u, MATERIAL_CUTOUT = 2u };
enum ShadingModelId : uint16_t { SHADING_DEFAULT = 0u, SHADING_DIFFUSE = 1u, SHADING_TRANSMISSIVE = 2u };

// UNorm16, Types.h:76-93
BPT_HD float unorm16_to_float(uint16_t raw) { return raw / 65535.0f; }
BPT_HD uint16_t float_to_unorm16(float v) { return (uint16_t)(saturate(v) * 65535.0f + 0.5f); }

BPT_HD bool material_is_thin_walled(const Material& m) { return (m.flags & (MATERIAL_CUTOUT | MATERIAL_THIN_WALLED)) != 0; }
BPT_HD bool material_is_cutout(const Material& m) { return (m.flags & MATERIAL_CUTOUT) != 0; }
BPT_HD bool material_is_transmissive(const Material& m) { return m.shading_model == SHADING_TRANSMISSIVE; }
// Material::get_coverage (Types.h:405-414) without a coverage texture.
BPT_HD float material_coverage(const Material& m) {
    if (material_is_cutout(m))
        return 1.0f < m.coverage ? 0.0f : 1.0f;
    return m.coverage * 1.0f;
}

// Entry of the per-material coverage table read by any-hit rays: negative = "sample the coverage texture".
BPT_HD float material_coverage_table_entry(const Material& m) { return m.coverage_texture_id ? -1.0f : material_coverage(m); }

BPT_HD bool material_is_textured(const Material& m) {
    return m.tint_roughness_texture_id || m.roughness_texture_id || m.metallic_texture_id || m.coverage_texture_id;
}

#ifdef __CUDACC__
// Texture lookups of Material (Types.h:388-414). `objects` maps texture ids to cudaTextureObject_t.
struct TextureView {
    const unsigned long long* __restrict__ objects;
    const float2* __restrict__ uv; // 3 texcoords per primitive (instance-major order) or nullptr when no mesh has texcoords
};

// interpolate_attributes, TriangleAttributes.cu:57-64: (0, 0) for meshes without texcoords.
BPT_D float2 interpolate_texcoord(const TextureView& tv, int primitive, float bx, float by) {
    if (tv.uv == nullptr) return make_float2(0.0f, 0.0f);
    const float2 t0 = __ldg(tv.uv + 3ll * primitive), t1 = __ldg(tv.uv + 3ll * primitive + 1), t2 = __ldg(tv.uv + 3ll * primitive + 2);
    const float bz = 1.0f - bx - by;
    return make_float2(t1.x * bx + t2.x * by + t0.x * bz, t1.y * bx + t2.y * by + t0.y * bz);
}

BPT_D float4 material_tint_roughness(const Material& m, const TextureView& tv, float2 texcoord) {
    float4 tint_roughness = make_float4(m.tint[0], m.tint[1], m.tint[2], m.roughness);
    if (m.tint_roughness_texture_id) {
        float4 s = tex2D<float4>(tv.objects[m.tint_roughness_texture_id], texcoord.x, texcoord.y);
        tint_roughness.x *= s.x; tint_roughness.y *= s.y; tint_roughness.z *= s.z; tint_roughness.w *= s.w;
    }
    if (m.roughness_texture_id)
        tint_roughness.w *= tex2D<float>(tv.objects[m.roughness_texture_id], texcoord.x, texcoord.y);
    return tint_roughness;
}

BPT_D float material_metallic(const Material& m, const TextureView& tv, float2 texcoord) {
    if (m.metallic_texture_id)
        return m.metallic * tex2D<float>(tv.objects[m.metallic_texture_id], texcoord.x, texcoord.y);
    return m.metallic;
}

BPT_D float material_coverage(const Material& m, const TextureView& tv, float2 texcoord) {
    float coverage_tex_sample = 1.0f;
    if (m.coverage_texture_id)
        coverage_tex_sample = tex2D<float>(tv.objects[m.coverage_texture_id], texcoord.x, texcoord.y);
    if (material_is_cutout(m))
        return coverage_tex_sample < m.coverage ? 0.0f : 1.0f;
    return m.coverage * coverage_tex_sample;
}

// The material with its textures applied at `texcoord`: what DefaultShading / DiffuseShading / TransmissiveShading read
// through get_tint_roughness and get_metallic (DefaultShading.h:159-166, MonteCarlo.cu:250-255, TransmissiveShading.h:55-58).
BPT_D Material material_at(const Material& m, const TextureView& tv, float2 texcoord) {
    if (!material_is_textured(m)) return m;
    Material r = m;
    float4 tr = material_tint_roughness(m, tv, texcoord);
    r.tint[0] = tr.x; r.tint[1] = tr.y; r.tint[2] = tr.z; r.roughness = tr.w;
    r.metallic = material_metallic(m, tv, texcoord);
    return r;
}
#endif

} // namespace bpt
