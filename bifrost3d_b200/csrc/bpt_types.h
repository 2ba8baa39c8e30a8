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

} // namespace bpt
