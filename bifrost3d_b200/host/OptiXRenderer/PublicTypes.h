// Public POD types of the renderer, API compatible with the reference's
// extensions/OptiXRenderer/OptiXRenderer/PublicTypes.h:20-58 (same names, enumerators and members).
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#ifndef _OPTIXRENDERER_PUBLIC_TYPES_H_
#define _OPTIXRENDERER_PUBLIC_TYPES_H_

#include <Bifrost/Core/Bitmask.h>

namespace OptiXRenderer {

enum class Backend { None, PathTracing, AIDenoisedPathTracing, DepthVisualization, AlbedoVisualization, TintVisualization,
                     RoughnessVisualization, ShadingNormalVisualization, PrimitiveIdVisualization };

// Roughness floor derived from the previous bounce's BSDF PDF; the scale may decay with the accumulation count.
struct PathRegularizationSettings {
    float PDF_scale;
    float scale_decay;
    float PDF_scale_at_accumulation(int accumulation) { return PDF_scale * (1.0f + scale_decay * accumulation); }
};

enum class AIDenoiserFlag : unsigned char { None = 0, LogarithmicFeedback = 1 << 0, VisualizeNoise = 1 << 1, VisualizeAlbedo = 1 << 2, Default = LogarithmicFeedback };
typedef Bifrost::Core::Bitmask<AIDenoiserFlag> AIDenoiserFlags;

} // namespace OptiXRenderer

#endif
