// C ABI: context, scene upload and the batched unit entry points (BSDF / light / RNG kernels).
// See include/bpt_c_api.h for the contract of every function and the reference interface it replaces.
#include "bpt_context.h"
#include "bpt_lights.cuh"
#include "bpt_rng.cuh"
#include "bpt_sort.cuh"

#include <math.h>
#include <string.h>

using namespace bpt;

namespace {

// ------------------------------------------------------------------------------------------------
// Batched BSDF kernel (C1 workload): evaluate_with_PDF(wo, wi) + sample(wo, u) per tuple.
// HBM-bound streaming kernel: 60 B in (68 with coat) + 44 B out per tuple. Inputs arrive as packed
// xyz triples; each warp stages its 32 tuples through shared memory so that global loads and stores
// are fully coalesced 128-bit transactions instead of stride-12 scalar accesses.
// ------------------------------------------------------------------------------------------------
#ifndef BPT_BSDF_BLOCK
#define BPT_BSDF_BLOCK 256
#endif
#ifndef BPT_BSDF_MIN_BLOCKS
#define BPT_BSDF_MIN_BLOCKS 4 // 64 registers, 4 x 256 threads per SM: +9 % over 78 registers x 3 CTAs (measured round 2)
#endif
constexpr int BSDF_BLOCK = BPT_BSDF_BLOCK;

template <int FLOATS_PER_ITEM>
__device__ __forceinline__ void stage_in(float* smem, const float* __restrict__ g, int64_t block_first, int64_t n, int block_items) {
    // Copies block_items * FLOATS_PER_ITEM consecutive floats starting at item block_first; vectorised when aligned.
    int64_t first = block_first * FLOATS_PER_ITEM;
    int64_t count = (int64_t)min((int64_t)block_items, n - block_first) * FLOATS_PER_ITEM;
    if (count <= 0) return;
    const float* src = g + first;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        int vec = int(count >> 2);
        const float4* src4 = reinterpret_cast<const float4*>(src);
        float4* dst4 = reinterpret_cast<float4*>(smem);
        for (int i = threadIdx.x; i < vec; i += blockDim.x) dst4[i] = __ldcs(src4 + i);
        for (int i = (vec << 2) + threadIdx.x; i < count; i += blockDim.x) smem[i] = src[i];
    } else
        for (int i = threadIdx.x; i < count; i += blockDim.x) smem[i] = src[i];
}

template <int FLOATS_PER_ITEM>
__device__ __forceinline__ void stage_out(const float* smem, float* __restrict__ g, int64_t block_first, int64_t n, int block_items) {
    int64_t first = block_first * FLOATS_PER_ITEM;
    int64_t count = (int64_t)min((int64_t)block_items, n - block_first) * FLOATS_PER_ITEM;
    if (count <= 0) return;
    float* dst = g + first;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        int vec = int(count >> 2);
        float4* dst4 = reinterpret_cast<float4*>(dst);
        const float4* src4 = reinterpret_cast<const float4*>(smem);
        for (int i = threadIdx.x; i < vec; i += blockDim.x) __stcs(dst4 + i, src4[i]);
        for (int i = (vec << 2) + threadIdx.x; i < count; i += blockDim.x) dst[i] = smem[i];
    } else
        for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = smem[i];
}

struct BsdfBatchArgs {
    const float* wo; const float* wi; const float* tint; const float* rms; const float* coat; const float* u;
    float* eval_f; float* eval_pdf; float* sample_f; float* sample_pdf; float* sample_dir;
    const float* tables;
    const float2* dielectric_tables;
    int64_t n;
};

template <int KIND>
__global__ void __launch_bounds__(BSDF_BLOCK, BPT_BSDF_MIN_BLOCKS) bsdf_batch_kernel(BsdfBatchArgs a) {
    // 3 tables (12 KB) + staging: 5 x 3 floats in, 2 floats coat, 11 floats out -> reuse the input area for output.
    __shared__ __align__(16) float s_tables[3 * TABLE_FLOATS];
    __shared__ __align__(16) float s_wo[BSDF_BLOCK * 3];
    __shared__ __align__(16) float s_wi[BSDF_BLOCK * 3];
    __shared__ __align__(16) float s_tint[BSDF_BLOCK * 3];
    __shared__ __align__(16) float s_rms[BSDF_BLOCK * 3];
    __shared__ __align__(16) float s_u[BSDF_BLOCK * 3];
    __shared__ __align__(16) float s_coat[BSDF_BLOCK * 2];
    __shared__ __align__(16) float s_pdf[BSDF_BLOCK * 2];

    if (KIND == BPT_BSDF_DEFAULT_SHADING)
        for (int i = threadIdx.x; i < 3 * TABLE_FLOATS; i += blockDim.x) s_tables[i] = a.tables[i];
    ShadingTables tables = { s_tables, s_tables + TABLE_FLOATS, s_tables + 2 * TABLE_FLOATS };

    int64_t tile_count = (a.n + BSDF_BLOCK - 1) / BSDF_BLOCK;
    for (int64_t tile = blockIdx.x; tile < tile_count; tile += gridDim.x) {
        int64_t first = tile * BSDF_BLOCK;
        __syncthreads(); // previous tile's stage_out has finished reading shared memory
        stage_in<3>(s_wo, a.wo, first, a.n, BSDF_BLOCK);
        stage_in<3>(s_wi, a.wi, first, a.n, BSDF_BLOCK);
        stage_in<3>(s_tint, a.tint, first, a.n, BSDF_BLOCK);
        stage_in<3>(s_rms, a.rms, first, a.n, BSDF_BLOCK);
        stage_in<3>(s_u, a.u, first, a.n, BSDF_BLOCK);
        if (a.coat) stage_in<2>(s_coat, a.coat, first, a.n, BSDF_BLOCK);
        __syncthreads();

        int t = threadIdx.x;
        BsdfResponse r = bsdf_response_none();
        BsdfSample s = bsdf_sample_none();
        if (first + t < a.n) {
            float3 wo = f3(s_wo[3 * t], s_wo[3 * t + 1], s_wo[3 * t + 2]);
            float3 wi = f3(s_wi[3 * t], s_wi[3 * t + 1], s_wi[3 * t + 2]);
            float3 tint = f3(s_tint[3 * t], s_tint[3 * t + 1], s_tint[3 * t + 2]);
            float roughness = s_rms[3 * t], metallic = s_rms[3 * t + 1], specularity = s_rms[3 * t + 2];
            float3 u = f3(s_u[3 * t], s_u[3 * t + 1], s_u[3 * t + 2]);
            if (KIND == BPT_BSDF_DEFAULT_SHADING) {
                // Material stores coat parameters as UNorm16 (Types.h:381-382); quantise like the host does.
                float coat = 0.0f, coat_roughness = 0.0f;
                if (a.coat) {
                    coat = unorm16_to_float(float_to_unorm16(s_coat[2 * t]));
                    coat_roughness = unorm16_to_float(float_to_unorm16(s_coat[2 * t + 1]));
                }
                DefaultShading shading = DefaultShading::create(tables, tint, roughness, specularity, metallic, coat, coat_roughness, wo.z);
                r = shading.evaluate_with_pdf(wo, wi);
                s = shading.sample(wo, u);
            } else if (KIND == BPT_BSDF_GGX_R) {
                float alpha = ggx::alpha_from_roughness(roughness);
                r = ggx_r::evaluate_with_pdf(alpha, tint, wo, wi);
                s = ggx_r::sample(alpha, tint, wo, f2(u.x, u.y));
            } else if (KIND == BPT_BSDF_OREN_NAYAR) {
                float roughness_factor = oren_nayar::uniform_lobe_roughness_factor(roughness);
                r = oren_nayar::evaluate_with_pdf(tint, roughness, roughness_factor, wo, wi);
                s = oren_nayar::sample(tint, roughness, roughness_factor, wo, f2(u.x, u.y));
            } else if (KIND == BPT_BSDF_BURLEY) {
                r = burley::evaluate_with_pdf(tint, roughness, wo, wi);
                s = burley::sample(tint, roughness, wo, f2(u.x, u.y));
            } else if (KIND == BPT_BSDF_TRANSMISSIVE_SHADING) {
                // rms.y carries the signed cos_theta_o of TransmissiveShading::setup_shading.
                TransmissiveShading shading = TransmissiveShading::create(a.dielectric_tables, tint, roughness, specularity, metallic);
                r = shading.evaluate_with_pdf(wo, wi);
                s = shading.sample(wo, u);
            } else {
                // rms.y carries ior_i_over_o of the combined GGX BSDF.
                float alpha = ggx::alpha_from_roughness(roughness);
                r = ggx_rt::evaluate_with_pdf(tint, alpha, specularity, metallic, wo, wi);
                s = ggx_rt::sample(tint, alpha, specularity, metallic, wo, u);
            }
        }
        __syncthreads(); // everyone has consumed the inputs; reuse the staging areas for the outputs
        s_wo[3 * t] = r.reflectance.x; s_wo[3 * t + 1] = r.reflectance.y; s_wo[3 * t + 2] = r.reflectance.z;
        s_wi[3 * t] = s.reflectance.x; s_wi[3 * t + 1] = s.reflectance.y; s_wi[3 * t + 2] = s.reflectance.z;
        s_tint[3 * t] = s.direction.x; s_tint[3 * t + 1] = s.direction.y; s_tint[3 * t + 2] = s.direction.z;
        s_pdf[t] = r.pdf.v; s_pdf[BSDF_BLOCK + t] = s.pdf.v;
        __syncthreads();
        stage_out<3>(s_wo, a.eval_f, first, a.n, BSDF_BLOCK);
        stage_out<3>(s_wi, a.sample_f, first, a.n, BSDF_BLOCK);
        stage_out<3>(s_tint, a.sample_dir, first, a.n, BSDF_BLOCK);
        stage_out<1>(s_pdf, a.eval_pdf, first, a.n, BSDF_BLOCK);
        stage_out<1>(s_pdf + BSDF_BLOCK, a.sample_pdf, first, a.n, BSDF_BLOCK);
    }
}

__global__ void default_shading_regularized_kernel(int64_t n, const Material* __restrict__ materials, const float* __restrict__ scale,
                                                   const float* __restrict__ max_pdf_hint, const float* __restrict__ wo_, const float* __restrict__ wi_,
                                                   const float* __restrict__ u_, const float* __restrict__ tables_,
                                                   float* eval_f, float* eval_pdf, float* sample_f, float* sample_pdf, float* sample_dir) {
    ShadingTables tables = { tables_, tables_ + TABLE_FLOATS, tables_ + 2 * TABLE_FLOATS };
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float3 wo = f3(wo_[3 * i], wo_[3 * i + 1], wo_[3 * i + 2]);
        float3 wi = f3(wi_[3 * i], wi_[3 * i + 1], wi_[3 * i + 2]);
        float3 u = f3(u_[3 * i], u_[3 * i + 1], u_[3 * i + 2]);
        float4 sc = scale ? make_float4(scale[4 * i], scale[4 * i + 1], scale[4 * i + 2], scale[4 * i + 3]) : make_float4(1.0f, 1.0f, 1.0f, 1.0f);
        Material m = materials[i];
        DefaultShading shading = DefaultShading::create_regularized(tables, m, sc, wo.z, Pdf(max_pdf_hint[i]));
        BsdfResponse r = shading.evaluate_with_pdf(wo, wi);
        BsdfSample s = shading.sample(wo, u);
        eval_f[3 * i] = r.reflectance.x; eval_f[3 * i + 1] = r.reflectance.y; eval_f[3 * i + 2] = r.reflectance.z; eval_pdf[i] = r.pdf.v;
        sample_f[3 * i] = s.reflectance.x; sample_f[3 * i + 1] = s.reflectance.y; sample_f[3 * i + 2] = s.reflectance.z; sample_pdf[i] = s.pdf.v;
        sample_dir[3 * i] = s.direction.x; sample_dir[3 * i + 1] = s.direction.y; sample_dir[3 * i + 2] = s.direction.z;
    }
}

__global__ void light_batch_kernel(int64_t n, const Light* __restrict__ lights, int light_stride, const float* __restrict__ position,
                                   const float* __restrict__ u2, const float* __restrict__ query, bpt_light_sample* out_samples,
                                   float* out_pdf, float* out_radiance, EnvironmentView env) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        Light light = lights[light_stride ? i : 0];
        float3 p = f3(position[3 * i], position[3 * i + 1], position[3 * i + 2]);
        float3 q = f3(query[3 * i], query[3 * i + 1], query[3 * i + 2]);
        float2 u = f2(u2[2 * i], u2[2 * i + 1]);
        LightSample s = light_sample_none();
        Pdf pdf = Pdf::invalid();
        float3 e = f3(0.0f);
        uint32_t type = light_type(light);
        const bool environment = (type == BPT_LIGHT_ENVIRONMENT || type == BPT_LIGHT_PRESAMPLED_ENVIRONMENT) && env.texels != nullptr;
        if (type == BPT_LIGHT_SPHERE || type == BPT_LIGHT_SPOT || type == BPT_LIGHT_DIRECTIONAL || environment) {
            s = light_sample_radiance(light, env, p, u);
            pdf = light_pdf(light, env, p, q);
            e = light_evaluate(light, env, p, q);
        }
        bpt_light_sample o;
        o.radiance[0] = s.radiance.x; o.radiance[1] = s.radiance.y; o.radiance[2] = s.radiance.z; o.pdf = s.pdf.v;
        o.direction_to_light[0] = s.direction_to_light.x; o.direction_to_light[1] = s.direction_to_light.y; o.direction_to_light[2] = s.direction_to_light.z;
        o.distance = s.distance;
        out_samples[i] = o;
        out_pdf[i] = pdf.v;
        out_radiance[3 * i] = e.x; out_radiance[3 * i + 1] = e.y; out_radiance[3 * i + 2] = e.z;
    }
}

__global__ void rng_batch_kernel(int64_t n, const uint32_t* __restrict__ accumulation, const uint32_t* __restrict__ pixel_hash,
                                 const uint32_t* __restrict__ dimension, uint4* out_ui4, float4* out_f4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (out_ui4) out_ui4[i] = sobol_sample4ui(accumulation[i], pixel_hash[i], dimension[i]);
        if (out_f4) out_f4[i] = sobol_sample4f(accumulation[i], pixel_hash[i], dimension[i]);
    }
}

// RNG::ReverseHalton(index).sample4f(), RNG.h:196-231: radical inverse with reversed digits in the first four prime bases.
// NB the reference builds the float4 as make_float4(sample2f(), sample2f()) with sample2f() = make_float2(sample1f(),
// sample1f()), and every sample1f() advances the prime index. C++ leaves the evaluation order of function arguments
// unspecified; MSVC (the reference's compiler) and GCC (the oracle's) both evaluate right to left on x86-64, so the
// components come out as (base 7, base 5, base 3, base 2). The golden fixture tests/golden (halton_offsets) pins this.
void reverse_halton4(int index, float out[4]) {
    static const int primes[4] = { 7, 5, 3, 2 };
    for (int d = 0; d < 4; ++d) {
        const int prime = primes[d];
        double h = 0.0, f = 1.0 / (double)prime, fct = f;
        int i = index;
        while (i > 0) {
            int digit = i % prime;
            h += (digit == 0 ? digit : (prime - digit)) * fct;
            i /= prime;
            fct *= f;
        }
        out[d] = (float)h;
    }
}

// OctahedralNormal::encode_precise, core/Bifrost/Bifrost/Math/OctahedralNormal.h:53-83.
void oct_decode(int16_t ex, int16_t ey, float n[3]) {
    float fx = float(ex), fy = float(ey);
    float nz = 32767 - fabsf(fx) - fabsf(fy);
    float t = fmaxf(-nz, 0.0f);
    fx += fx >= 0 ? -t : t;
    fy += fy >= 0 ? -t : t;
    float inv_len = 1.0f / sqrtf(fx * fx + fy * fy + nz * nz);
    n[0] = fx * inv_len; n[1] = fy * inv_len; n[2] = nz * inv_len;
}

void oct_encode_precise(const float n[3], int16_t out[2]) {
    auto sgn = [](float v) { return v >= 0.0f ? 1.0f : -1.0f; };
    auto clamp1 = [](float v) { return fminf(fmaxf(v, -1.0f), 1.0f); };
    float denom = fabsf(n[0]) + fabsf(n[1]) + fabsf(n[2]);
    float px = n[0] / denom, py = n[1] / denom;
    float p2x = px, p2y = py;
    if (n[2] < 0) {
        p2x = (1.0f - fabsf(py)) * sgn(px);
        p2y = (1.0f - fabsf(px)) * sgn(py);
    }
    int16_t fx = (int16_t)floorf(clamp1(p2x) * 32767), fy = (int16_t)floorf(clamp1(p2y) * 32767);
    int16_t best[2] = { fx, fy };
    auto error = [&](int16_t ex, int16_t ey) {
        float d[3]; oct_decode(ex, ey, d);
        float dx = d[0] - n[0], dy = d[1] - n[1], dz = d[2] - n[2];
        return dx * dx + dy * dy + dz * dz;
    };
    float lowest = error(fx, fy);
    const int16_t cand[3][2] = { { fx, int16_t(fy + 1) }, { int16_t(fx + 1), fy }, { int16_t(fx + 1), int16_t(fy + 1) } };
    for (auto& c : cand) {
        float m = error(c[0], c[1]);
        if (m < lowest) { best[0] = c[0]; best[1] = c[1]; lowest = m; }
    }
    out[0] = best[0]; out[1] = best[1];
}

int grid_for(Context* ctx, int64_t n, int block) {
    int64_t blocks = (n + block - 1) / block;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

template <typename T>
cudaError_t upload(Context* ctx, DeviceBuffer<T>& buf, const T* host, size_t n) {
    cudaError_t e = buf.resize(n);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(buf.ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
}

} // namespace

namespace {
// Small RAII helper for the host-pointer unit entry points: device scratch that mirrors host arrays.
__global__ void texture_sample_kernel(cudaTextureObject_t texture, int64_t n, const float2* __restrict__ uv, float4* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = tex2D<float4>(texture, uv[i].x, uv[i].y);
}

void destroy_texture(DeviceTexture& t) {
    if (t.object) cudaDestroyTextureObject(t.object);
    if (t.array) cudaFreeArray(t.array);
    t = DeviceTexture();
}

struct Scratch {
    Context* ctx;
    std::vector<void*> allocations;
    bool failed = false; // an allocation failed: the caller returns BPT_ERROR_OUT_OF_MEMORY instead of launching with null pointers
    explicit Scratch(Context* c) : ctx(c) {}
    ~Scratch() { for (void* p : allocations) cudaFreeAsync(p, ctx->stream); cudaStreamSynchronize(ctx->stream); }
    template <typename T> T* in(const T* host, size_t n) {
        if (!host || n == 0) return nullptr;
        T* d = nullptr;
        if (cudaMallocAsync((void**)&d, n * sizeof(T), ctx->stream) != cudaSuccess) { failed = true; cudaGetLastError(); return nullptr; }
        allocations.push_back(d);
        cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
        return d;
    }
    template <typename T> T* out(size_t n) {
        T* d = nullptr;
        if (n == 0) return nullptr;
        if (cudaMallocAsync((void**)&d, n * sizeof(T), ctx->stream) != cudaSuccess) { failed = true; cudaGetLastError(); return nullptr; }
        allocations.push_back(d);
        return d;
    }
    template <typename T> void back(T* host, const T* dev, size_t n) {
        if (host && dev && n) cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
    }
};
} // namespace

namespace bpt {
int sync_texture_table(Context* ctx) {
    if (!ctx->texture_table_dirty) return BPT_OK;
    int max_id = ctx->textures.empty() ? 0 : ctx->textures.rbegin()->first;
    std::vector<unsigned long long> table(max_id + 1, 0ull);
    for (const auto& kv : ctx->textures) table[kv.first] = (unsigned long long)kv.second.object;
    BPT_CUDA_CHECK(ctx, ctx->texture_objects.resize(table.size()));
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->texture_objects.ptr, table.data(), table.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->texture_table_dirty = false;
    return BPT_OK;
}
} // namespace bpt

extern "C" {

int bpt_create(int cuda_device, bpt_ctx** out_ctx) {
    if (!out_ctx) return BPT_ERROR_INVALID_ARGUMENT;
    *out_ctx = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        fprintf(stderr, "bpt_create: no CUDA device available; this library has no CPU fallback.\n");
        return BPT_ERROR_NO_DEVICE;
    }
    if (cuda_device < 0 || cuda_device >= count) return BPT_ERROR_INVALID_ARGUMENT;
    if (cudaSetDevice(cuda_device) != cudaSuccess) return BPT_ERROR_CUDA;
    Context* ctx = new Context();
    ctx->device = cuda_device;
    if (const char* bvh = getenv("BPT_BVH")) ctx->use_ploc = strcmp(bvh, "lbvh") != 0;
    if (const char* wide = getenv("BPT_WIDE")) ctx->use_wide = strcmp(wide, "0") != 0;
    if (const char* cw = getenv("BPT_CW")) ctx->cw_min_triangles = strcmp(cw, "0") != 0 ? 0 : INT64_MAX;
    if (const char* sort_hits = getenv("BPT_SORT_HITS")) ctx->sort_hits_from_iteration = atoi(sort_hits);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cuda_device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BPT_ERROR_CUDA; }
    { // keep stream-ordered scratch allocations cached instead of returning them to the OS at every synchronise
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, cuda_device) == cudaSuccess) {
            uint64_t threshold = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
    }
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    cudaMalloc((void**)&ctx->device_counters, 8 * sizeof(uint64_t));
    cudaMemsetAsync(ctx->device_counters, 0, 8 * sizeof(uint64_t), ctx->stream);

    // g_random_sample_offsets, Renderer.cpp:323-336
    std::vector<float4> offsets(256);
    for (int i = 0; i < 256; ++i) { float v[4]; reverse_halton4(i, v); offsets[i] = make_float4(v[0], v[1], v[2], v[3]); }
    upload(ctx, ctx->nee_offsets, offsets.data(), offsets.size());
    cudaStreamSynchronize(ctx->stream);

    // A single empty material 0 and a black environment so that a context is renderable straight away.
    Material m = {}; m.coverage = 1.0f;
    ctx->host_materials.assign(1, m);
    upload(ctx, ctx->materials, ctx->host_materials.data(), 1);
    cudaStreamSynchronize(ctx->stream);
    *out_ctx = reinterpret_cast<bpt_ctx*>(ctx);
    return BPT_OK;
}

void bpt_destroy(bpt_ctx* c) {
    if (!c) return;
    Context* ctx = as_context(c);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    bpt_comm_destroy(c);
    release_wavefront(ctx);
    ctx->tables.release(); ctx->dielectric_tables.release(); ctx->nee_offsets.release(); ctx->materials.release(); ctx->lights.release();
    ctx->env_texels.release(); ctx->env_pdf.release(); ctx->env_samples.release();
    for (auto& kv : ctx->meshes) kv.second.release();
    for (auto& kv : ctx->textures) destroy_texture(kv.second);
    ctx->textures.clear(); ctx->texture_objects.release(); ctx->accel.shade_uv.release(); ctx->accel.shade_emission.release();
    ctx->accel.nodes.release(); ctx->accel.wide_nodes.release(); ctx->accel.cw_nodes.release(); ctx->accel.triangles.release(); ctx->accel.slot_of_primitive.release(); ctx->accel.world_vertices.release(); ctx->accel.shade.release(); ctx->accel.normal_matrices.release();
    ctx->accumulation.release(); ctx->output_half4.release(); ctx->output_float4.release(); ctx->query_scratch.release();
    for (auto& target : ctx->parked_targets) target.second.buffer.release();
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        for (int i = 0; i < BPT_FRAME_SLOTS; ++i) { cudaEventDestroy(ctx->frame_resolved[i]); cudaEventDestroy(ctx->frame_copied[i]); ctx->frame_staging[i].release(); }
        cudaStreamDestroy(ctx->copy_stream);
    }
    if (ctx->device_counters) cudaFree(ctx->device_counters);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto& e : ctx->stage_events) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* bpt_last_error(const bpt_ctx* c) { return c ? as_context(c)->last_error.c_str() : "null context"; }
void* bpt_stream(bpt_ctx* c) { return c ? (void*)as_context(c)->stream : nullptr; }

int bpt_set_profiling(bpt_ctx* c, int enabled) {
    as_context(c)->profiling = enabled != 0;
    return BPT_OK;
}

int bpt_synchronize(bpt_ctx* c) {
    Context* ctx = as_context(c);
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return BPT_OK;
}

int bpt_set_tables(bpt_ctx* c, const float* ggx_with_fresnel_rho, const float* ggx_rho, const float* estimate_ggx_alpha) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (!ggx_with_fresnel_rho || !ggx_rho || !estimate_ggx_alpha) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_tables: null table");
    cudaSetDevice(ctx->device);
    std::vector<float> all(3 * TABLE_FLOATS);
    memcpy(all.data(), ggx_with_fresnel_rho, TABLE_FLOATS * sizeof(float));
    memcpy(all.data() + TABLE_FLOATS, ggx_rho, TABLE_FLOATS * sizeof(float));
    memcpy(all.data() + 2 * TABLE_FLOATS, estimate_ggx_alpha, TABLE_FLOATS * sizeof(float));
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->tables, all.data(), all.size()));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_tables = true;
    return BPT_OK;
}

int bpt_set_dielectric_tables(bpt_ctx* c, const float* into_light_medium, const float* into_dense_medium) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (!into_light_medium || !into_dense_medium) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_dielectric_tables: null table");
    cudaSetDevice(ctx->device);
    std::vector<float2> all(2 * DIELECTRIC_TABLE_FLOAT2S);
    memcpy(all.data(), into_light_medium, DIELECTRIC_TABLE_FLOAT2S * sizeof(float2));
    memcpy(all.data() + DIELECTRIC_TABLE_FLOAT2S, into_dense_medium, DIELECTRIC_TABLE_FLOAT2S * sizeof(float2));
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->dielectric_tables, all.data(), all.size()));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_dielectric_tables = true;
    return BPT_OK;
}

// Image + sampler creation, Renderer.cpp:650-751.
int bpt_upload_texture(bpt_ctx* c, int texture_id, const bpt_texture_desc* desc, const void* pixels) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (texture_id < 1 || texture_id > (1 << 20) || !desc || !pixels || desc->width <= 0 || desc->height <= 0)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_upload_texture: bad arguments (texture ids start at 1)");
    if (desc->wrap_u < BPT_WRAP_CLAMP || desc->wrap_u > BPT_WRAP_REPEAT || desc->wrap_v < BPT_WRAP_CLAMP || desc->wrap_v > BPT_WRAP_REPEAT)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_upload_texture: unknown wrap mode");
    cudaSetDevice(ctx->device);
    const size_t texels = (size_t)desc->width * desc->height;
    cudaChannelFormatDesc channels;
    std::vector<unsigned char> widened_bytes; std::vector<float> widened_floats;
    const void* source = pixels;
    size_t texel_bytes = 0;
    int channel_count = 4;
    bool eight_bit = true;
    switch (desc->pixel_format) {
    case BPT_PIXEL_ALPHA8: channels = cudaCreateChannelDesc<unsigned char>(); texel_bytes = 1; channel_count = 1; break;
    case BPT_PIXEL_RGB24: { // Renderer.cpp:683-694: ubyte3 cannot back a sampler, widen with alpha 255
        channels = cudaCreateChannelDesc<uchar4>(); texel_bytes = 4;
        widened_bytes.resize(4 * texels);
        const unsigned char* p = static_cast<const unsigned char*>(pixels);
        for (size_t i = 0; i < texels; ++i) { widened_bytes[4 * i] = p[3 * i]; widened_bytes[4 * i + 1] = p[3 * i + 1]; widened_bytes[4 * i + 2] = p[3 * i + 2]; widened_bytes[4 * i + 3] = 255; }
        source = widened_bytes.data();
        break;
    }
    case BPT_PIXEL_RGBA32: channels = cudaCreateChannelDesc<uchar4>(); texel_bytes = 4; break;
    case BPT_PIXEL_RGB_FLOAT: { // CUDA arrays have no three-channel float format: widen with alpha 1
        channels = cudaCreateChannelDesc<float4>(); texel_bytes = 16; eight_bit = false;
        widened_floats.resize(4 * texels);
        const float* p = static_cast<const float*>(pixels);
        for (size_t i = 0; i < texels; ++i) { widened_floats[4 * i] = p[3 * i]; widened_floats[4 * i + 1] = p[3 * i + 1]; widened_floats[4 * i + 2] = p[3 * i + 2]; widened_floats[4 * i + 3] = 1.0f; }
        source = widened_floats.data();
        break;
    }
    case BPT_PIXEL_RGBA_FLOAT: channels = cudaCreateChannelDesc<float4>(); texel_bytes = 16; eight_bit = false; break;
    default: return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_upload_texture: unsupported pixel format (Alpha8, RGB24, RGBA32, RGB_Float, RGBA_Float)");
    }

    DeviceTexture fresh;
    fresh.width = desc->width; fresh.height = desc->height; fresh.channels = channel_count;
    BPT_CUDA_CHECK(ctx, cudaMallocArray(&fresh.array, &channels, desc->width, desc->height));
    cudaError_t e = cudaMemcpy2DToArrayAsync(fresh.array, 0, 0, source, desc->width * texel_bytes, desc->width * texel_bytes, desc->height,
                                             cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // `source` may be a local staging vector
    cudaResourceDesc resource = {};
    resource.resType = cudaResourceTypeArray;
    resource.res.array.array = fresh.array;
    cudaTextureDesc sampler = {};
    sampler.addressMode[0] = desc->wrap_u == BPT_WRAP_REPEAT ? cudaAddressModeWrap : cudaAddressModeClamp;
    sampler.addressMode[1] = desc->wrap_v == BPT_WRAP_REPEAT ? cudaAddressModeWrap : cudaAddressModeClamp;
    sampler.filterMode = desc->linear_filter ? cudaFilterModeLinear : cudaFilterModePoint;
    sampler.readMode = eight_bit ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
    sampler.sRGB = (eight_bit && desc->is_srgb) ? 1 : 0;
    sampler.normalizedCoords = 1;
    if (e == cudaSuccess) e = cudaCreateTextureObject(&fresh.object, &resource, &sampler, nullptr);
    if (e != cudaSuccess) { destroy_texture(fresh); return ctx->cuda_fail(e, "bpt_upload_texture"); }

    DeviceTexture& slot = ctx->textures[texture_id];
    destroy_texture(slot);
    slot = fresh;
    ctx->texture_table_dirty = true;
    return BPT_OK;
}

int bpt_destroy_texture(bpt_ctx* c, int texture_id) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    auto it = ctx->textures.find(texture_id);
    if (it == ctx->textures.end()) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_destroy_texture: unknown texture id");
    for (const Material& m : ctx->host_materials)
        if (m.tint_roughness_texture_id == texture_id || m.roughness_texture_id == texture_id || m.metallic_texture_id == texture_id || m.coverage_texture_id == texture_id)
            return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_destroy_texture: a material still references the texture");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    destroy_texture(it->second);
    ctx->textures.erase(it);
    ctx->texture_table_dirty = true;
    return BPT_OK;
}

int bpt_texture_sample(bpt_ctx* c, int texture_id, int64_t n, const float* uv, float* out_rgba) {
    Context* ctx = as_context(c);
    auto it = ctx->textures.find(texture_id);
    if (it == ctx->textures.end() || n < 0 || !uv || !out_rgba) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_texture_sample: bad arguments or unknown texture id");
    if (n == 0) return BPT_OK;
    cudaSetDevice(ctx->device);
    {
        Scratch s(ctx);
        auto d_uv = s.in(reinterpret_cast<const float2*>(uv), n);
        auto d_out = s.out<float4>(n);
        if (s.failed) return ctx->cuda_fail(cudaErrorMemoryAllocation, "scratch allocation");
        texture_sample_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(it->second.object, n, d_uv, d_out);
        ctx->counters.kernel_launches++;
        s.back(reinterpret_cast<float4*>(out_rgba), d_out, n);
    }
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

int bpt_upload_mesh(bpt_ctx* c, int mesh_id, const uint32_t* indices, int primitive_count, const float* positions, const float* normals,
                    const float* texcoords, const uint8_t* tint_roughness, int vertex_count) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (!indices || !positions || primitive_count < 0 || vertex_count < 0)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_upload_mesh: null indices/positions or negative counts");
    for (int64_t i = 0; i < 3ll * primitive_count; ++i)
        if (indices[i] >= (uint32_t)vertex_count)
            return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_upload_mesh: vertex index out of range");
    cudaSetDevice(ctx->device);
    DeviceMesh& m = ctx->meshes[mesh_id];
    m.primitive_count = primitive_count; m.vertex_count = vertex_count;
    BPT_CUDA_CHECK(ctx, upload(ctx, m.indices, indices, 3ull * primitive_count));
    BPT_CUDA_CHECK(ctx, upload(ctx, m.positions, positions, 3ull * vertex_count));
    m.normals.release(); m.texcoords.release(); m.tints.release(); m.emission.release();
    std::vector<int16_t> encoded_normals;
    if (normals) {
        encoded_normals.resize(2ull * vertex_count);
        for (int v = 0; v < vertex_count; ++v) oct_encode_precise(normals + 3ll * v, encoded_normals.data() + 2ll * v);
        BPT_CUDA_CHECK(ctx, upload(ctx, m.normals, encoded_normals.data(), encoded_normals.size()));
    }
    if (texcoords) BPT_CUDA_CHECK(ctx, upload(ctx, m.texcoords, texcoords, 2ull * vertex_count));
    if (tint_roughness) BPT_CUDA_CHECK(ctx, upload(ctx, m.tints, tint_roughness, 4ull * vertex_count));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); // the caller's arrays (and encoded_normals) may go away
    ctx->accel.valid = false;
    return BPT_OK;
}

int bpt_set_mesh_emission(bpt_ctx* c, int mesh_id, const float* emission, int vertex_count) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    auto it = ctx->meshes.find(mesh_id);
    if (it == ctx->meshes.end()) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_mesh_emission: unknown mesh id");
    if (emission && vertex_count != it->second.vertex_count)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_mesh_emission: vertex count differs from the mesh");
    cudaSetDevice(ctx->device);
    if (emission) {
        BPT_CUDA_CHECK(ctx, upload(ctx, it->second.emission, emission, 3ull * vertex_count));
        BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    } else
        it->second.emission.release();
    ctx->accel.valid = false;
    return BPT_OK;
}

int bpt_remove_mesh(bpt_ctx* c, int mesh_id) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    auto it = ctx->meshes.find(mesh_id);
    if (it == ctx->meshes.end()) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_remove_mesh: unknown mesh id");
    for (const bpt_instance& inst : ctx->instances)
        if (inst.mesh_id == mesh_id) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_remove_mesh: an instance still references the mesh");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    it->second.release();
    ctx->meshes.erase(it);
    return BPT_OK;
}

int bpt_set_instances(bpt_ctx* c, const bpt_instance* instances, int count) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (count < 0 || (count > 0 && !instances)) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_instances: bad arguments");
    for (int i = 0; i < count; ++i)
        if (ctx->meshes.find(instances[i].mesh_id) == ctx->meshes.end())
            return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_instances: unknown mesh id");
    ctx->instances.assign(instances, instances + count);
    ctx->accel.valid = false;
    return BPT_OK;
}

int bpt_set_materials(bpt_ctx* c, const bpt_material* materials, int count) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (count <= 0 || !materials) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_materials: need at least material 0");
    bool any_transmissive = false, any_textured = false;
    for (int i = 0; i < count; ++i) {
        if (materials[i].shading_model > SHADING_TRANSMISSIVE)
            return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_materials: unknown shading model");
        any_transmissive |= materials[i].shading_model == SHADING_TRANSMISSIVE;
        for (int id : { materials[i].tint_roughness_texture_id, materials[i].roughness_texture_id, materials[i].metallic_texture_id, materials[i].coverage_texture_id })
            if (id != 0 && ctx->textures.find(id) == ctx->textures.end())
                return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_materials: a material references a texture id that was not uploaded");
        // Renderer.cpp:763-780,789,803 assert the channel counts; a mismatch would sample garbage
        if (materials[i].tint_roughness_texture_id && ctx->textures[materials[i].tint_roughness_texture_id].channels != 4)
            return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_materials: the tint/roughness texture needs four channels");
        for (int id : { materials[i].roughness_texture_id, materials[i].metallic_texture_id, materials[i].coverage_texture_id })
            if (id != 0 && ctx->textures[id].channels != 1)
                return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_materials: roughness, metallic and coverage textures need one channel");
        any_textured |= material_is_textured(materials[i]);
    }
    cudaSetDevice(ctx->device);
    ctx->host_materials.assign(materials, materials + count);
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->materials, materials, (size_t)count));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_transmissive_materials = any_transmissive;
    ctx->has_textured_materials = any_textured;
    ctx->material_version++;
    // Materials are read at shading time: the acceleration structure stays valid, unless the scene now needs the
    // per-primitive texcoords that a build made without textured materials in view left out. (A build made WITH them in view
    // that found no mesh with texcoords has nothing to add: later material edits keep it.)
    if (any_textured && !ctx->accel.built_for_textures) ctx->accel.valid = false;
    for (const bpt_instance& inst : ctx->instances)
        if (inst.material_id >= count) ctx->accel.valid = false; // bpt_build_accel will report the dangling material id
    return BPT_OK;
}

int bpt_set_lights(bpt_ctx* c, const bpt_light* lights, int count) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (count < 0 || (count > 0 && !lights)) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_lights: bad arguments");
    for (int i = 0; i < count; ++i) {
        uint32_t t = lights[i].flags & BPT_LIGHT_TYPE_MASK;
        if (t != BPT_LIGHT_SPHERE && t != BPT_LIGHT_SPOT && t != BPT_LIGHT_DIRECTIONAL)
            return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_lights: only sphere, spot and directional lights can be set here");
    }
    cudaSetDevice(ctx->device);
    // One extra slot: the environment light is appended last when present (Renderer.cpp:1180-1195).
    std::vector<Light> all(lights, lights + count);
    all.resize(count + 1);
    memset(&all[count], 0, sizeof(Light));
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->lights, all.data(), all.size()));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->light_count = count;
    ctx->env_light_uploaded = false;
    return BPT_OK;
}

int bpt_set_environment(bpt_ctx* c, const float tint[3], const float* texels, int width, int height, const float* per_pixel_pdf,
                        int pdf_width, int pdf_height, const bpt_light_sample* samples, int sample_count) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (!tint) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_environment: null tint");
    cudaSetDevice(ctx->device);
    memcpy(ctx->env_tint, tint, 3 * sizeof(float));
    ctx->env_light_uploaded = false;
    ctx->env_has_cdfs = false; // the CDFs belong to one map: bpt_set_environment_cdfs follows bpt_set_environment
    if (!texels) {
        ctx->env_width = ctx->env_height = ctx->env_pdf_width = ctx->env_pdf_height = ctx->env_sample_count = 0;
        return BPT_OK;
    }
    if (width <= 0 || height <= 0 || !per_pixel_pdf || pdf_width <= 0 || pdf_height <= 0 || !samples || sample_count <= 0)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_environment: an environment map needs texels, per pixel PDF and presampled lights");
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->env_texels, reinterpret_cast<const float4*>(texels), (size_t)width * height));
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->env_pdf, per_pixel_pdf, (size_t)pdf_width * pdf_height));
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->env_samples, samples, (size_t)sample_count));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->env_width = width; ctx->env_height = height; ctx->env_pdf_width = pdf_width; ctx->env_pdf_height = pdf_height;
    ctx->env_sample_count = sample_count;
    return BPT_OK;
}

int bpt_set_environment_cdfs(bpt_ctx* c, const float* marginal_cdf, const float* conditional_cdf, int pdf_width, int pdf_height) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    cudaSetDevice(ctx->device);
    ctx->env_light_uploaded = false;
    if (!marginal_cdf && !conditional_cdf) { ctx->env_has_cdfs = false; return BPT_OK; }
    if (!marginal_cdf || !conditional_cdf) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_environment_cdfs: both CDFs or none");
    if (ctx->env_width <= 0) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_set_environment_cdfs: call bpt_set_environment with a map first");
    if (pdf_width != ctx->env_pdf_width || pdf_height != ctx->env_pdf_height)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_environment_cdfs: the CDFs must have the size of the per pixel PDF");
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->env_marginal_cdf, marginal_cdf, (size_t)pdf_height + 1));
    BPT_CUDA_CHECK(ctx, upload(ctx, ctx->env_conditional_cdf, conditional_cdf, (size_t)(pdf_width + 1) * pdf_height));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->env_has_cdfs = true;
    return BPT_OK;
}

int bpt_set_environment_sampling(bpt_ctx* c, int mode) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    if (mode != BPT_ENVIRONMENT_NEE_PRESAMPLED && mode != BPT_ENVIRONMENT_NEE_CDF)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_set_environment_sampling: unknown mode");
    if (ctx->env_nee_mode != mode) ctx->env_light_uploaded = false;
    ctx->env_nee_mode = mode;
    return BPT_OK;
}

int bpt_set_hit_sorting(bpt_ctx* c, int from_iteration) {
    as_context(c)->scene_epoch++;
    as_context(c)->sort_hits_from_iteration = from_iteration < 0 ? -1 : from_iteration;
    return BPT_OK;
}

int bpt_build_accel(bpt_ctx* c) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    cudaSetDevice(ctx->device);
    return build_accel(ctx);
}

int bpt_accel_info(bpt_ctx* c, int64_t* triangle_count, int64_t* node_count, float* build_ms) {
    Context* ctx = as_context(c);
    if (!ctx->accel.valid) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_accel_info: no acceleration structure built");
    if (triangle_count) *triangle_count = ctx->accel.triangle_count;
    if (node_count) *node_count = ctx->accel.node_count;
    if (build_ms) *build_ms = ctx->accel.build_ms;
    return BPT_OK;
}

int bpt_accel_hierarchy(bpt_ctx* c, int* kind, int64_t* node_count, int* levels) {
    Context* ctx = as_context(c);
    if (!ctx->accel.valid) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_accel_hierarchy: no acceleration structure built");
    const Accel& A = ctx->accel;
    if (kind) *kind = A.cw_levels > 0 ? 8 : (A.wide_levels > 0 ? 4 : 2);
    if (node_count) *node_count = A.cw_levels > 0 ? A.cw_node_count : (A.wide_levels > 0 ? A.wide_node_count : A.node_count);
    if (levels) *levels = A.cw_levels > 0 ? A.cw_levels : (A.wide_levels > 0 ? A.wide_levels : A.ploc_depth);
    return BPT_OK;
}

int bpt_render(bpt_ctx* c, const bpt_camera* camera, const bpt_settings* settings, int width, int height,
               uint32_t first_sample, uint32_t sample_count, int reset_accumulation) {
    Context* ctx = as_context(c);
    cudaSetDevice(ctx->device);
    return render(ctx, camera, settings, width, height, first_sample, sample_count, reset_accumulation);
}

int bpt_render_aov(bpt_ctx* c, const bpt_camera* camera, int aov_kind, int width, int height, uint32_t first_sample, uint32_t sample_count, int reset_accumulation) {
    Context* ctx = as_context(c);
    ctx->scene_epoch++;
    cudaSetDevice(ctx->device);
    return render_aov(ctx, camera, aov_kind, width, height, first_sample, sample_count, reset_accumulation);
}

void* bpt_accumulation_device_ptr(bpt_ctx* c) { return as_context(c)->accumulation.ptr; }

int bpt_select_accumulation(bpt_ctx* c, int slot) {
    Context* ctx = as_context(c);
    if (slot < 0) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_select_accumulation: slot must be >= 0");
    if (slot == ctx->selected_target) return BPT_OK;
    // Work already enqueued on the stream keeps using the buffers it was given; only the host-side selection moves.
    Context::AccumulationTarget& parked = ctx->parked_targets[ctx->selected_target];
    parked.buffer = ctx->accumulation; parked.width = ctx->width; parked.height = ctx->height; parked.half4_scale = ctx->half4_scale;
    auto it = ctx->parked_targets.find(slot);
    if (it != ctx->parked_targets.end()) {
        ctx->accumulation = it->second.buffer; ctx->width = it->second.width; ctx->height = it->second.height; ctx->half4_scale = it->second.half4_scale;
        ctx->parked_targets.erase(it);
    } else {
        ctx->accumulation = DeviceBuffer<double>(); ctx->width = ctx->height = 0; ctx->half4_scale = 1.0f;
    }
    ctx->selected_target = slot;
    return BPT_OK;
}

int bpt_read_accumulation(bpt_ctx* c, double* sums, int* out_width, int* out_height) {
    Context* ctx = as_context(c);
    if (out_width) *out_width = ctx->width;
    if (out_height) *out_height = ctx->height;
    if (!sums) return BPT_OK;
    const size_t count = 4ull * (size_t)ctx->width * (size_t)ctx->height;
    if (count == 0 || !ctx->accumulation.ptr) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_read_accumulation: nothing rendered into the selected target");
    cudaSetDevice(ctx->device);
    // the main stream is ordered behind every sample's accumulation, whichever lane rendered it
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(sums, ctx->accumulation.ptr, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return BPT_OK;
}

int bpt_write_accumulation(bpt_ctx* c, int width, int height, const double* sums) {
    Context* ctx = as_context(c);
    if (width <= 0 || height <= 0 || !sums || (int64_t)width * height > 0x7fffffffll)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_write_accumulation: bad arguments");
    cudaSetDevice(ctx->device);
    const size_t count = 4ull * (size_t)width * (size_t)height;
    if (ctx->accumulation.capacity < count) BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); // the old buffer may still be in use
    BPT_CUDA_CHECK(ctx, ctx->accumulation.resize(count));
    ctx->width = width; ctx->height = height; ctx->half4_scale = 1.0f;
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->accumulation.ptr, sums, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream)); // `sums` is the caller's memory
    return BPT_OK;
}

int bpt_release_accumulation(bpt_ctx* c, int slot) {
    Context* ctx = as_context(c);
    cudaSetDevice(ctx->device);
    if (slot == ctx->selected_target) {
        BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->accumulation.release(); ctx->width = ctx->height = 0; ctx->half4_scale = 1.0f;
        return BPT_OK;
    }
    auto it = ctx->parked_targets.find(slot);
    if (it == ctx->parked_targets.end()) return BPT_OK;
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    it->second.buffer.release();
    ctx->parked_targets.erase(it);
    return BPT_OK;
}
int bpt_resolve_half4(bpt_ctx* c, uint16_t* out, int on_device) { Context* ctx = as_context(c); cudaSetDevice(ctx->device); return resolve_half4(ctx, out, on_device); }
int bpt_resolve_half4_async(bpt_ctx* c, uint16_t* out_host, int slot) { Context* ctx = as_context(c); cudaSetDevice(ctx->device); return resolve_half4_async(ctx, out_host, slot); }
int bpt_wait_frame(bpt_ctx* c, int slot) { Context* ctx = as_context(c); cudaSetDevice(ctx->device); return wait_frame(ctx, slot); }
int bpt_resolve_tonemapped(bpt_ctx* c, const bpt_tonemap_settings* settings, void* out, int output_format) {
    Context* ctx = as_context(c); cudaSetDevice(ctx->device); return resolve_tonemapped(ctx, settings, out, output_format);
}
int bpt_tonemap_colors(bpt_ctx* c, const bpt_tonemap_settings* settings, int64_t n, const float* rgb_in, float* rgb_out) {
    Context* ctx = as_context(c); cudaSetDevice(ctx->device); return tonemap_batch(ctx, settings, n, rgb_in, rgb_out);
}
int bpt_resolve_float4(bpt_ctx* c, float* out) { Context* ctx = as_context(c); cudaSetDevice(ctx->device); return resolve_float4(ctx, out); }

int bpt_get_counters(bpt_ctx* c, bpt_counters* out, int reset) {
    Context* ctx = as_context(c);
    cudaSetDevice(ctx->device);
    uint64_t host[8];
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(host, ctx->device_counters, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->counters.extend_rays = host[0];
    ctx->counters.shadow_rays = host[1];
    ctx->counters.extend_node_visits = host[2];   // only counted by builds with -DBPT_TRAVERSAL_STATS
    ctx->counters.extend_triangle_tests = host[3];
    ctx->counters.traversal_stack_overflows = host[5];
    ctx->counters.nonfinite_samples = host[6];
    ctx->counters.iterations = host[7];
    if (out) {
        *out = ctx->counters;
        out->kernel_launches += host[7] * (uint64_t)ctx->launches_per_iteration; // wavefront iterations, counted by advance_kernel
    }
    if (reset) {
        ctx->counters = {};
        BPT_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->device_counters, 0, sizeof(host), ctx->stream));
    }
    return BPT_OK;
}

// ---- batched unit entry points -------------------------------------------------------------------

int bpt_bsdf_eval_sample_pdf(bpt_ctx* c, int kind, int64_t n, const float* wo, const float* wi, const float* tint, const float* rms,
                             const float* coat, const float* u, float* eval_f, float* eval_pdf, float* sample_f, float* sample_pdf,
                             float* sample_dir, int on_device) {
    Context* ctx = as_context(c);
    if (kind < 0 || kind > BPT_BSDF_GGX || n < 0 || !wo || !wi || !tint || !rms || !u || !eval_f || !eval_pdf || !sample_f || !sample_pdf || !sample_dir)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_bsdf_eval_sample_pdf: bad arguments");
    if (kind == BPT_BSDF_DEFAULT_SHADING && !ctx->has_tables)
        return ctx->fail(BPT_ERROR_NOT_READY, "bpt_bsdf_eval_sample_pdf: call bpt_set_tables first");
    if (kind == BPT_BSDF_TRANSMISSIVE_SHADING && !ctx->has_dielectric_tables)
        return ctx->fail(BPT_ERROR_NOT_READY, "bpt_bsdf_eval_sample_pdf: call bpt_set_dielectric_tables first");
    if (n == 0) return BPT_OK;
    cudaSetDevice(ctx->device);

    BsdfBatchArgs a;
    a.n = n; a.tables = ctx->tables.ptr; a.dielectric_tables = ctx->dielectric_tables.ptr;
    float* staging = nullptr;
    // frees the staging block on every exit, error paths included
    struct StagingGuard { float*& p; cudaStream_t st; ~StagingGuard() { if (p) { cudaFreeAsync(p, st); cudaStreamSynchronize(st); } } } staging_guard = { staging, ctx->stream };
    if (on_device) {
        a.wo = wo; a.wi = wi; a.tint = tint; a.rms = rms; a.coat = coat; a.u = u;
        a.eval_f = eval_f; a.eval_pdf = eval_pdf; a.sample_f = sample_f; a.sample_pdf = sample_pdf; a.sample_dir = sample_dir;
    } else {
        // [wo wi tint rms u | coat | eval_f sample_f sample_dir | eval_pdf sample_pdf], each segment 16 byte aligned.
        size_t n4 = (size_t)((n + 3) & ~int64_t(3));
        size_t total = n4 * (15 + 2 + 9 + 2);
        BPT_CUDA_CHECK(ctx, cudaMallocAsync((void**)&staging, total * sizeof(float), ctx->stream));
        float* p = staging;
        const float* src3[5] = { wo, wi, tint, rms, u };
        const float** dst3[5] = { &a.wo, &a.wi, &a.tint, &a.rms, &a.u };
        for (int k = 0; k < 5; ++k) {
            BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(p, src3[k], 3 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
            *dst3[k] = p; p += 3 * n4;
        }
        a.coat = nullptr;
        if (coat) { BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(p, coat, 2 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream)); a.coat = p; }
        p += 2 * n4;
        a.eval_f = p; p += 3 * n4; a.sample_f = p; p += 3 * n4; a.sample_dir = p; p += 3 * n4;
        a.eval_pdf = p; p += n4; a.sample_pdf = p; p += n4;
    }

    int grid = grid_for(ctx, n, BSDF_BLOCK);
    switch (kind) {
    case BPT_BSDF_DEFAULT_SHADING: bsdf_batch_kernel<BPT_BSDF_DEFAULT_SHADING><<<grid, BSDF_BLOCK, 0, ctx->stream>>>(a); break;
    case BPT_BSDF_GGX_R: bsdf_batch_kernel<BPT_BSDF_GGX_R><<<grid, BSDF_BLOCK, 0, ctx->stream>>>(a); break;
    case BPT_BSDF_OREN_NAYAR: bsdf_batch_kernel<BPT_BSDF_OREN_NAYAR><<<grid, BSDF_BLOCK, 0, ctx->stream>>>(a); break;
    case BPT_BSDF_BURLEY: bsdf_batch_kernel<BPT_BSDF_BURLEY><<<grid, BSDF_BLOCK, 0, ctx->stream>>>(a); break;
    case BPT_BSDF_TRANSMISSIVE_SHADING: bsdf_batch_kernel<BPT_BSDF_TRANSMISSIVE_SHADING><<<grid, BSDF_BLOCK, 0, ctx->stream>>>(a); break;
    default: bsdf_batch_kernel<BPT_BSDF_GGX><<<grid, BSDF_BLOCK, 0, ctx->stream>>>(a); break;
    }
    ctx->counters.kernel_launches++;
    BPT_CUDA_CHECK(ctx, cudaGetLastError());

    if (!on_device) {
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(eval_f, a.eval_f, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(sample_f, a.sample_f, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(sample_dir, a.sample_dir, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(eval_pdf, a.eval_pdf, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(sample_pdf, a.sample_pdf, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return BPT_OK;
}

int bpt_sort_pairs(bpt_ctx* c, int64_t n, uint64_t* keys, uint32_t* values, int begin_bit, int end_bit) {
    Context* ctx = as_context(c);
    if (n < 0 || n > 0x7fffffff || !keys || !values || begin_bit < 0 || end_bit > 64 || begin_bit >= end_bit)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_sort_pairs: bad arguments");
    if (n == 0) return BPT_OK;
    cudaSetDevice(ctx->device);
    {
        Scratch s(ctx);
        uint64_t* d_keys = s.in(keys, n); uint32_t* d_values = s.in(values, n);
        uint64_t* d_keys_alt = s.out<uint64_t>(n); uint32_t* d_values_alt = s.out<uint32_t>(n);
        uint32_t* d_scratch = s.out<uint32_t>(sort::sort_scratch_words((uint32_t)n));
        if (s.failed) return ctx->cuda_fail(cudaErrorMemoryAllocation, "scratch allocation");
        const int in_alt = sort::radix_sort_pairs<uint64_t>(d_keys, d_values, d_keys_alt, d_values_alt, (uint32_t)n, nullptr, begin_bit, end_bit, d_scratch,
                                                            ctx->sm_count, ctx->stream, &ctx->counters.kernel_launches);
        s.back(keys, in_alt ? d_keys_alt : d_keys, n); s.back(values, in_alt ? d_values_alt : d_values, n);
    }
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

int bpt_exclusive_scan(bpt_ctx* c, int64_t n, const uint32_t* in, uint32_t* out, uint32_t* out_total) {
    Context* ctx = as_context(c);
    if (n < 0 || n > 0x7fffffff || !in || !out) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_exclusive_scan: bad arguments");
    if (n == 0) { if (out_total) *out_total = 0; return BPT_OK; }
    cudaSetDevice(ctx->device);
    {
        Scratch s(ctx);
        const uint32_t* d_in = s.in(in, n); uint32_t* d_out = s.out<uint32_t>(n);
        uint32_t* d_scratch = s.out<uint32_t>(sort::scan_scratch_words((uint32_t)n)); uint32_t* d_total = s.out<uint32_t>(1);
        if (s.failed) return ctx->cuda_fail(cudaErrorMemoryAllocation, "scratch allocation");
        sort::exclusive_scan(d_in, d_out, (uint32_t)n, nullptr, d_scratch, d_total, ctx->sm_count, ctx->stream);
        ctx->counters.kernel_launches += sort::SCAN_LAUNCHES;
        s.back(out, d_out, n); s.back(out_total, d_total, 1);
    }
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}


int bpt_default_shading_regularized(bpt_ctx* c, int64_t n, const bpt_material* materials, const float* tint_roughness_scale,
                                    const float* max_pdf_hint, const float* wo, const float* wi, const float* u,
                                    float* eval_f, float* eval_pdf, float* sample_f, float* sample_pdf, float* sample_dir) {
    Context* ctx = as_context(c);
    if (n < 0 || !materials || !max_pdf_hint || !wo || !wi || !u) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_default_shading_regularized: bad arguments");
    if (!ctx->has_tables) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_default_shading_regularized: call bpt_set_tables first");
    if (n == 0) return BPT_OK;
    cudaSetDevice(ctx->device);
    {
        Scratch s(ctx);
        auto d_m = s.in(materials, n); auto d_sc = s.in(tint_roughness_scale, 4 * n); auto d_hint = s.in(max_pdf_hint, n);
        auto d_wo = s.in(wo, 3 * n); auto d_wi = s.in(wi, 3 * n); auto d_u = s.in(u, 3 * n);
        auto o_ef = s.out<float>(3 * n); auto o_ep = s.out<float>(n); auto o_sf = s.out<float>(3 * n); auto o_sp = s.out<float>(n); auto o_sd = s.out<float>(3 * n);
        if (s.failed) return ctx->cuda_fail(cudaErrorMemoryAllocation, "scratch allocation");
        default_shading_regularized_kernel<<<grid_for(ctx, n, 128), 128, 0, ctx->stream>>>(n, d_m, d_sc, d_hint, d_wo, d_wi, d_u, ctx->tables.ptr, o_ef, o_ep, o_sf, o_sp, o_sd);
        ctx->counters.kernel_launches++;
        s.back(eval_f, o_ef, 3 * n); s.back(eval_pdf, o_ep, n); s.back(sample_f, o_sf, 3 * n); s.back(sample_pdf, o_sp, n); s.back(sample_dir, o_sd, 3 * n);
    }
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

int bpt_light_sample_pdf_evaluate(bpt_ctx* c, int64_t n, const bpt_light* lights, int light_stride, const float* position, const float* u2,
                                  const float* query_direction, bpt_light_sample* out_samples, float* out_pdf, float* out_radiance) {
    Context* ctx = as_context(c);
    if (n < 0 || !lights || !position || !u2 || !query_direction || !out_samples || !out_pdf || !out_radiance)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_light_sample_pdf_evaluate: bad arguments");
    if (n == 0) return BPT_OK;
    cudaSetDevice(ctx->device);
    {
        Scratch s(ctx);
        auto d_l = s.in(lights, light_stride ? n : 1); auto d_p = s.in(position, 3 * n); auto d_u = s.in(u2, 2 * n); auto d_q = s.in(query_direction, 3 * n);
        auto o_s = s.out<bpt_light_sample>(n); auto o_p = s.out<float>(n); auto o_r = s.out<float>(3 * n);
        if (s.failed) return ctx->cuda_fail(cudaErrorMemoryAllocation, "scratch allocation");
        light_batch_kernel<<<grid_for(ctx, n, 128), 128, 0, ctx->stream>>>(n, d_l, light_stride, d_p, d_u, d_q, o_s, o_p, o_r, environment_view(ctx, true));
        ctx->counters.kernel_launches++;
        s.back(out_samples, o_s, n); s.back(out_pdf, o_p, n); s.back(out_radiance, o_r, 3 * n);
    }
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

int bpt_rng_sample4(bpt_ctx* c, int64_t n, const uint32_t* accumulation, const uint32_t* pixel_hash, const uint32_t* dimension,
                    uint32_t* out_ui4, float* out_f4) {
    Context* ctx = as_context(c);
    if (n < 0 || !accumulation || !pixel_hash || !dimension) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_rng_sample4: bad arguments");
    if (n == 0) return BPT_OK;
    cudaSetDevice(ctx->device);
    {
        Scratch s(ctx);
        auto d_a = s.in(accumulation, n); auto d_h = s.in(pixel_hash, n); auto d_d = s.in(dimension, n);
        uint4* o_u = out_ui4 ? s.out<uint4>(n) : nullptr; float4* o_f = out_f4 ? s.out<float4>(n) : nullptr;
        if (s.failed) return ctx->cuda_fail(cudaErrorMemoryAllocation, "scratch allocation");
        rng_batch_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(n, d_a, d_h, d_d, o_u, o_f);
        ctx->counters.kernel_launches++;
        s.back(reinterpret_cast<uint4*>(out_ui4), o_u, n); s.back(reinterpret_cast<float4*>(out_f4), o_f, n);
    }
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

int bpt_intersect(bpt_ctx* c, int64_t n, const float* origins, const float* directions, const float* tmin, const float* tmax,
                  int32_t* out_primitive, float* out_t, float* out_uv, uint8_t* out_occluded) {
    Context* ctx = as_context(c);
    cudaSetDevice(ctx->device);
    return intersect_batch(ctx, n, origins, directions, tmin, tmax, out_primitive, out_t, out_uv, out_occluded);
}

} // extern "C"
