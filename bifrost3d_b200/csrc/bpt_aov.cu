// First-hit AOV backends: depth, albedo, tint, roughness, shading normal, primitive id.
// Replaces the ray generation programs depth_RPG, albedo_RPG, tint_RPG, roughness_RPG, shading_normal_RPG and
// primitive_id_RPG (Shading/SimpleRGPs.cu:227-340) and their use by Renderer::request_auxiliary_buffers
// (Renderer.cpp:1267-1358). Like the reference they follow the camera ray through rejected hits (back faces, stochastic
// coverage: MonteCarlo.cu:146-164) until the first accepted surface, then accumulate the feature in the same fp64 buffer.
// Not a hot path: one thread per pixel, the whole loop in one kernel.
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#include "bpt_context.h"
#include "bpt_lights.cuh"
#include "bpt_rng.cuh"
#include "bpt_trace.cuh"
#include <type_traits>

namespace bpt {

namespace {

struct AovParams {
    bpt_camera camera;
    int width, height;
    unsigned int accumulation_count;
    int kind;
};

__device__ __forceinline__ float4 mul4x4(const float* m, float4 v) {
    return make_float4(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w, m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w,
                       m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w, m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w);
}

__device__ __forceinline__ float3 oct_decode(const int16_t e[2]) {
    float2 fe = f2(float(e[0]), float(e[1]));
    float3 n = f3(fe.x, fe.y, 32767 - fabsf(fe.x) - fabsf(fe.y));
    float t = fmaxf(-n.z, 0.0f);
    n.x += n.x >= 0 ? -t : t;
    n.y += n.y >= 0 ? -t : t;
    return normalize(n);
}

// float_to_unorm8 followed by unorm8_to_float (Utils.h:281-290): the payload carries the vertex scale as uchar4.
__device__ __forceinline__ float through_unorm8(float v) { return float((unsigned char)(saturate(v) * 255.0f + 0.5f)) * (1 / 255.0f); }

// primitive_id_to_color, SimpleRGPs.cu:329-334
__device__ __forceinline__ unsigned int compact_by_2(unsigned int v) {
    v &= 0x09249249; v = (v ^ (v >> 2)) & 0x030c30c3; v = (v ^ (v >> 4)) & 0x0300f00f; v = (v ^ (v >> 8)) & 0xff0000ff; v = (v ^ (v >> 16)) & 0x000003ff;
    return v;
}

template <bool COMPRESSED>
__global__ void __launch_bounds__(TRACE_BLOCK) aov_kernel(AccelView accel, const float4* __restrict__ world_vertices, const ShadeTriangle* __restrict__ shade,
                                                          const float* __restrict__ normal_matrices, const Material* __restrict__ materials,
                                                          const float* __restrict__ coverage, const Light* __restrict__ lights, int analytic_light_count,
                                                          const float* __restrict__ tables_, const float2* __restrict__ dielectric_tables, AovParams f, float4* __restrict__ out) {
    __shared__ __align__(16) int s_stack[STACK_SMEM * TRACE_BLOCK];
    typedef typename std::conditional<COMPRESSED, TraversalCW<false>, Traversal<false>>::type Trav;
    typename Trav::Spill spill;
    const ShadingTables tables = { tables_, tables_ + TABLE_FLOATS, tables_ + 2 * TABLE_FLOATS };
    int64_t pixel_count = (int64_t)f.width * f.height;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pixel_count; p += (int64_t)gridDim.x * blockDim.x) {
        // initialize_monte_carlo_payload, SimpleRGPs.cu:56-72
        int x = int(p % f.width), y = int(p / f.width);
        unsigned int pixel_hash = pcg2d((unsigned int)x, (unsigned int)y).x;
        float2 jitter = f2(0.5f, 0.5f);
        if (f.accumulation_count != 0) { float4 r = path_rng_sample4f(f.accumulation_count, pixel_hash, 0u, DIM_CAMERA); jitter = f2(r.x, r.y); }
        float2 viewport_pos = f2((float(x) + jitter.x) / float(f.width), (float(y) + jitter.y) / float(f.height));
        float4 ndc_near = make_float4(viewport_pos.x * 2.0f - 1.0f, viewport_pos.y * 2.0f - 1.0f, -1.0f, 1.0f);
        float4 near_world = mul4x4(f.camera.inverse_view_projection, ndc_near);
        float3 origin = f3(near_world) / near_world.w;
        float4 far_view = mul4x4(f.camera.inverse_projection, make_float4(ndc_near.x, ndc_near.y, 1.0f, 1.0f));
        const float* r = f.camera.view_to_world_rotation;
        float3 v = f3(far_view);
        float3 direction = normalize(f3(r[0] * v.x + r[1] * v.y + r[2] * v.z, r[3] * v.x + r[4] * v.y + r[5] * v.z, r[6] * v.x + r[7] * v.y + r[8] * v.z));

        float tmin = 0.0f, depth = 0.0f;
        float3 value = f3(0.0f);
        // process_material_intersection, SimpleRGPs.cu:265-280: trace until a surface is accepted or the ray leaves the scene.
        for (int guard = 0; guard < 4096; ++guard) {
            Ray ray; ray.origin = origin; ray.direction = direction; ray.tmin = tmin; ray.tmax = 1e27f;
            Trav tr;
            tr.attach(s_stack + threadIdx.x, spill);
            tr.begin(accel, ray, -1);
            tr.run(accel, coverage, 0x7fffffff);
            Hit h = tr.result();
            float t_closest = h.primitive >= 0 ? h.t : 1e27f;
            bool light_hit = false;
            for (int l = 0; l < analytic_light_count; ++l) {
                Light light = lights[l];
                float t = -1e30f, radius = 0.0f;
                if (light_type(light) == BPT_LIGHT_SPHERE) { SphereLight sl = as_sphere(light); radius = sl.radius; t = isect::ray_sphere(origin, direction, sl.position, sl.radius); }
                else if (light_type(light) == BPT_LIGHT_SPOT) { SpotLight sp = as_spot(light); radius = sp.radius; t = isect::ray_disk(origin, direction, sp.position, sp.direction, sp.radius); }
                if (radius > 0.0f && t > tmin && t < t_closest) { t_closest = t; light_hit = true; }
            }
            if (light_hit) { depth += t_closest; break; }                    // light_closest_hit: position = origin + t * direction, path ends
            if (h.primitive < 0) { depth += 1e30f; break; }                  // miss: position = 1e30 * direction, path ends

            const float3 p0 = f3(world_vertices[3ll * h.primitive]), p1 = f3(world_vertices[3ll * h.primitive + 1]), p2 = f3(world_vertices[3ll * h.primitive + 2]);
            const ShadeTriangle st = shade[h.primitive];
            const float2 texcoord = interpolate_texcoord(accel.textures, h.primitive, h.u, h.v);
            const Material m = material_at(materials[st.material_index], accel.textures, texcoord);
            float3 geometric_normal = normalize(cross(p1 - p0, p2 - p0));
            bool hit_from_front = dot(geometric_normal, direction) < 0.0f;
            bool backside_cull = !hit_from_front && !material_is_thin_walled(m) && !material_is_transmissive(m);
            float4 bsdf_random = path_rng_sample4f(f.accumulation_count, pixel_hash, 0u, DIM_BSDF);
            if (backside_cull || material_coverage(m, accel.textures, texcoord) < bsdf_random.w) { tmin = nextafterf(h.t, INFINITY); continue; }

            const float bx = h.u, by = h.v, bz = 1.0f - bx - by;
            float4 scale = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            if (st.flags & 2u) {
                const float n255 = 1.0f / 255.0f;
                scale.x = (st.t1[0] * bx + st.t2[0] * by + st.t0[0] * bz) * n255; scale.y = (st.t1[1] * bx + st.t2[1] * by + st.t0[1] * bz) * n255;
                scale.z = (st.t1[2] * bx + st.t2[2] * by + st.t0[2] * bz) * n255; scale.w = (st.t1[3] * bx + st.t2[3] * by + st.t0[3] * bz) * n255;
            }
            scale = make_float4(through_unorm8(scale.x), through_unorm8(scale.y), through_unorm8(scale.z), through_unorm8(scale.w));
            float3 shading_normal = geometric_normal;
            if (st.flags & 1u) {
                float3 n = normalize(oct_decode(st.n1) * bx + oct_decode(st.n2) * by + oct_decode(st.n0) * bz);
                const float* nm = normal_matrices + 9 * (st.flags >> 2);
                shading_normal = normalize(f3(nm[0] * n.x + nm[1] * n.y + nm[2] * n.z, nm[3] * n.x + nm[4] * n.y + nm[5] * n.z, nm[6] * n.x + nm[7] * n.y + nm[8] * n.z));
            }
            shading_normal = hit_from_front ? shading_normal : -shading_normal;
            { // fix_backfacing_shading_normal(-direction, n, 0.002), Utils.h:67-74
                float cos_theta = dot(-direction, shading_normal);
                if (cos_theta < 0.002f) shading_normal = normalize(shading_normal - (cos_theta - 0.002f) * (-direction));
            }
            // the reference measures depth to the offset origin of the next ray (<= 2^-16 relative away from the surface)
            depth += length(origin - (p1 * bx + p2 * by + p0 * bz));

            const float3 tint = f3(m.tint[0] * scale.x, m.tint[1] * scale.y, m.tint[2] * scale.z);
            switch (f.kind) {
            case BPT_AOV_TINT: value = tint; break;
            case BPT_AOV_ROUGHNESS: { float rough = m.roughness * scale.w; value = f3(rough); break; }
            case BPT_AOV_SHADING_NORMAL: value = shading_normal * 0.5f + 0.5f; break;
            case BPT_AOV_PRIMITIVE_ID: {
                // instance ids are not kept after flattening: the global primitive index alone keys the colour
                unsigned int primitive_encoding = __brev((unsigned int)h.primitive + 1u) >> 2;
                value = f3(float(compact_by_2(primitive_encoding >> 2)), float(compact_by_2(primitive_encoding >> 1)), float(compact_by_2(primitive_encoding))) / 1023.0f;
                break;
            }
            case BPT_AOV_ALBEDO: {
                float abs_cos_theta = fabsf(dot(direction, shading_normal));
                if (m.shading_model == SHADING_DIFFUSE) value = tint;
                else if (m.shading_model == SHADING_TRANSMISSIVE) {
                    // SimpleRGPs.cu:293-295
                    TransmissiveShading s = TransmissiveShading::create(dielectric_tables, tint, m.roughness * scale.w, m.specularity, abs_cos_theta);
                    value = s.rho(dielectric_tables, abs_cos_theta);
                } else {
                    DefaultShading s = DefaultShading::create(tables, tint, m.roughness * scale.w, m.specularity, m.metallic, unorm16_to_float(m.coat),
                                                              unorm16_to_float(m.coat_roughness), abs_cos_theta);
                    value = s.rho(tables, abs_cos_theta);
                }
                break;
            }
            default: break;
            }
            break;
        }
        if (f.kind == BPT_AOV_DEPTH) value = f3(depth);
        out[p] = f4(value, 0.0f);
    }
}

} // namespace

__global__ void accumulate_kernel_aov(const float4* __restrict__ rad, double* __restrict__ accum, int64_t pixel_count) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pixel_count; p += (int64_t)gridDim.x * blockDim.x) {
        float4 r = rad[p];
        accum[4 * p] += (double)r.x; accum[4 * p + 1] += (double)r.y; accum[4 * p + 2] += (double)r.z; accum[4 * p + 3] += 1.0;
    }
}

int render_aov(Context* ctx, const bpt_camera* camera, int kind, int width, int height, uint32_t first_sample, uint32_t sample_count, int reset_accumulation) {
    if (!camera || width <= 0 || height <= 0 || kind < BPT_AOV_DEPTH || kind > BPT_AOV_PRIMITIVE_ID)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_render_aov: bad arguments");
    if (!ctx->accel.valid) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_render_aov: call bpt_build_accel first");
    if (!ctx->has_tables) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_render_aov: call bpt_set_tables first");
    if (kind == BPT_AOV_ALBEDO && ctx->has_transmissive_materials && !ctx->has_dielectric_tables)
        return ctx->fail(BPT_ERROR_NOT_READY, "bpt_render_aov: transmissive materials need bpt_set_dielectric_tables");
    if (int status = sync_texture_table(ctx)) return status;
    cudaStream_t st = ctx->stream;
    const int64_t pixels = (int64_t)width * height;
    bool size_changed = ctx->width != width || ctx->height != height || ctx->accumulation.size != (size_t)(4 * pixels);
    if (size_changed) { BPT_CUDA_CHECK(ctx, ctx->accumulation.resize(4 * pixels)); ctx->width = width; ctx->height = height; }
    if (size_changed || reset_accumulation) BPT_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->accumulation.ptr, 0, sizeof(double) * 4 * pixels, st));

    std::vector<float> h_cov(ctx->host_materials.size());
    for (size_t i = 0; i < h_cov.size(); ++i) h_cov[i] = material_coverage_table_entry(ctx->host_materials[i]);
    float* d_cov = nullptr; float4* d_out = nullptr;
    BPT_CUDA_CHECK(ctx, cudaMallocAsync((void**)&d_cov, h_cov.size() * sizeof(float), st));
    BPT_CUDA_CHECK(ctx, cudaMallocAsync((void**)&d_out, pixels * sizeof(float4), st));
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(d_cov, h_cov.data(), h_cov.size() * sizeof(float), cudaMemcpyHostToDevice, st));

    ctx->half4_scale = 1.0f;
    if (kind == BPT_AOV_DEPTH) {
        // max depth = distance between the centres of the near and far planes, SimpleRGPs.cu:247-255
        const float* m = camera->inverse_projection;
        float near_z = (m[8] * 0.0f + m[9] * 0.0f + m[10] * -1.0f + m[11]) / (m[12] * 0.0f + m[13] * 0.0f + m[14] * -1.0f + m[15]);
        float far_z = (m[10] * 1.0f + m[11]) / (m[14] * 1.0f + m[15]);
        ctx->half4_scale = far_z - near_z;
    }
    AccelView accel = accel_view(ctx);
    AovParams f = {};
    f.camera = *camera; f.width = width; f.height = height; f.kind = kind;
    const int grid = ctx->sm_count * 8;
    for (uint32_t k = 0; k < sample_count; ++k) {
        f.accumulation_count = first_sample + k;
        if (accel.cw)
            aov_kernel<true><<<grid, TRACE_BLOCK, 0, st>>>(accel, ctx->accel.world_vertices.ptr, ctx->accel.shade.ptr, ctx->accel.normal_matrices.ptr, ctx->materials.ptr,
                                                            d_cov, ctx->lights.ptr, ctx->light_count, ctx->tables.ptr, ctx->dielectric_tables.ptr, f, d_out);
        else
            aov_kernel<false><<<grid, TRACE_BLOCK, 0, st>>>(accel, ctx->accel.world_vertices.ptr, ctx->accel.shade.ptr, ctx->accel.normal_matrices.ptr, ctx->materials.ptr,
                                                             d_cov, ctx->lights.ptr, ctx->light_count, ctx->tables.ptr, ctx->dielectric_tables.ptr, f, d_out);
        accumulate_kernel_aov<<<grid, 256, 0, st>>>(d_out, ctx->accumulation.ptr, pixels);
        ctx->counters.kernel_launches += 2;
    }
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    BPT_CUDA_CHECK(ctx, cudaFreeAsync(d_cov, st));
    BPT_CUDA_CHECK(ctx, cudaFreeAsync(d_out, st));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    return BPT_OK;
}

} // namespace bpt
