// Tonemapped resolve of the accumulation buffer: mean radiance -> exposure -> tonemapping operator -> linear float4 or
// sRGB-encoded RGBA8. The operators are the ones of the core's camera effects (core/Bifrost/Bifrost/Math/CameraEffects.h:
// TonemappingMode :18, filmic :161-224, AgX :236-265, Khronos neutral :272-291), which the reference applies to the
// renderer's output outside OptiXRenderer (DX11Renderer compositor). They are restated here for fp32 on the device; the
// parity test drives the reference header itself on the same colours (tests/test_tonemap.py).
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#include "bpt_context.h"
#include "bpt_math.cuh"

namespace bpt {

namespace {

struct Mat3 { float m[9]; };
BPT_HD float3 mul(const Mat3& a, float3 v) {
    return f3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z, a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}

struct TonemapConstants {
    Mat3 srgb_to_ap1, ap1_to_srgb; // CameraEffects.h:134-157
    float3 ap1_rgb2y;
};

// sRGB -> ACEScg (AP1) and back, composed on the host in double precision from the published matrices.
TonemapConstants make_constants() {
    const double d65_to_d60[9] = { 1.01303, 0.00610531, -0.014971, 0.00769823, 0.998165, -0.00503203, -0.00284131, 0.00468516, 0.924507 };
    const double srgb_to_xyz[9] = { 0.4124564, 0.3575761, 0.1804375, 0.2126729, 0.7151522, 0.0721750, 0.0193339, 0.1191920, 0.9503041 };
    const double xyz_to_ap1[9] = { 1.6410233797, -0.3248032942, -0.2364246952, -0.6636628587, 1.6153315917, 0.0167563477, 0.0117218943, -0.0082844420, 0.9883948585 };
    auto matmul = [](const double* a, const double* b, double* c) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { c[3 * i + j] = 0; for (int k = 0; k < 3; ++k) c[3 * i + j] += a[3 * i + k] * b[3 * k + j]; }
    };
    double t[9], a[9], inv[9];
    matmul(xyz_to_ap1, d65_to_d60, t);
    matmul(t, srgb_to_xyz, a);
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    inv[0] = (a[4] * a[8] - a[5] * a[7]) / det; inv[1] = (a[2] * a[7] - a[1] * a[8]) / det; inv[2] = (a[1] * a[5] - a[2] * a[4]) / det;
    inv[3] = (a[5] * a[6] - a[3] * a[8]) / det; inv[4] = (a[0] * a[8] - a[2] * a[6]) / det; inv[5] = (a[2] * a[3] - a[0] * a[5]) / det;
    inv[6] = (a[3] * a[7] - a[4] * a[6]) / det; inv[7] = (a[1] * a[6] - a[0] * a[7]) / det; inv[8] = (a[0] * a[4] - a[1] * a[3]) / det;
    TonemapConstants c;
    for (int i = 0; i < 9; ++i) { c.srgb_to_ap1.m[i] = float(a[i]); c.ap1_to_srgb.m[i] = float(inv[i]); }
    c.ap1_rgb2y = f3(0.2722287168f, 0.6740817658f, 0.0536895174f); // row 1 of AP1_to_XYZ
    return c;
}

// CameraEffects.h:161-224 (the Unreal Engine 4 film curve in ACEScg).
__device__ float3 filmic(const TonemapConstants& k, float3 color, const bpt_tonemap_settings& s) {
    float3 working = max3(f3(0.0f), mul(k.srgb_to_ap1, color));
    working = lerp(f3(dot(working, k.ap1_rgb2y)), working, 0.96f); // pre desaturate

    const float toe_scale = 1.0f + s.black_clip - s.toe;
    const float shoulder_scale = 1.0f + s.white_clip - s.shoulder;
    const float in_match = 0.18f, out_match = 0.18f;
    float toe_match;
    if (s.toe > 0.8f)
        toe_match = (1.0f - s.toe - out_match) / s.slope + log10f(in_match);
    else {
        const float bt = (out_match + s.black_clip) / toe_scale - 1.0f;
        toe_match = log10f(in_match) - 0.5f * logf((1.0f + bt) / (1.0f - bt)) * (toe_scale / s.slope);
    }
    const float straight_match = (1.0f - s.toe) / s.slope - toe_match;
    const float shoulder_match = s.shoulder / s.slope - straight_match;

    auto curve = [&](float c) {
        float log_c = log10f(c);
        float straight = (log_c + straight_match) * s.slope;
        float toe = (-s.black_clip) + (2.0f * toe_scale) / (1.0f + expf((log_c - toe_match) * (-2.0f * s.slope / toe_scale)));
        toe = log_c < toe_match ? toe : straight;
        float shoulder = (1.0f + s.white_clip) - (2.0f * shoulder_scale) / (1.0f + expf((log_c - shoulder_match) * (2.0f * s.slope / shoulder_scale)));
        shoulder = log_c > shoulder_match ? shoulder : straight;
        float t = clampf((log_c - toe_match) / (shoulder_match - toe_match), 0.0f, 1.0f);
        t = shoulder_match < toe_match ? 1.0f - t : t;
        t = (3.0f - t * 2.0f) * t * t;
        return toe + t * (shoulder - toe);
    };
    float3 tone = f3(curve(working.x), curve(working.y), curve(working.z));
    tone = lerp(f3(dot(tone, k.ap1_rgb2y)), tone, 0.93f); // post desaturate
    return mul(k.ap1_to_srgb, max3(f3(0.0f), tone));
}

// CameraEffects.h:231-265
__device__ float3 agx(float3 color) {
    const Mat3 linear_to_agx = { { 0.842479062253094f, 0.0784335999999992f, 0.0792237451477643f, 0.0423282422610123f, 0.878468636469772f,
                                   0.0791661274605434f, 0.0423756549057051f, 0.0784336f, 0.879142973793104f } };
    const Mat3 agx_to_tonemapped = { { 1.19687900512017f, -0.0980208811401368f, -0.0990297440797205f, -0.0528968517574562f, 1.15190312990417f,
                                       -0.0989611768448433f, -0.0529716355144438f, -0.0980434501171241f, 1.15107367264116f } };
    float3 c = mul(linear_to_agx, color);
    const float min_ev = -12.47393f, max_ev = 4.026069f;
    auto encode = [&](float v) {
        float x = saturate((log2f(v) - min_ev) / (max_ev - min_ev));
        return -0.00232f + x * (0.1191f + x * (0.4298f + x * (-6.868f + x * (31.96f + x * (-40.14f + x * 15.5f))))); // sigmoid fit
    };
    c = mul(agx_to_tonemapped, f3(encode(c.x), encode(c.y), encode(c.z)));
    return f3(powf(c.x, 2.2f), powf(c.y, 2.2f), powf(c.z, 2.2f));
}

// CameraEffects.h:272-291
__device__ float3 khronos_neutral(float3 c) {
    const float start_compression = 0.8f - 0.04f;
    const float desaturation = 0.15f;
    float x = fminf(c.x, fminf(c.y, c.z));
    float offset = x < 0.08f ? x - 6.25f * x * x : 0.04f;
    c = c - offset;
    float peak = fmaxf(c.x, fmaxf(c.y, c.z));
    if (peak < start_compression) return c;
    float d = 1.0f - start_compression;
    float new_peak = 1.0f - d * d / (peak + d - start_compression);
    c *= new_peak / peak;
    float g = 1.0f - 1.0f / (desaturation * (peak - new_peak) + 1.0f);
    return lerp(c, f3(new_peak), g);
}

__device__ float3 tonemap(const TonemapConstants& k, const bpt_tonemap_settings& s, float3 color) {
    color = color * s.exposure;
    switch (s.mode) {
    case BPT_TONEMAP_FILMIC: return filmic(k, color, s);
    case BPT_TONEMAP_AGX: return agx(color);
    case BPT_TONEMAP_KHRONOS_NEUTRAL: return khronos_neutral(color);
    default: return color;
    }
}

// Color.h:372-377
__device__ float linear_to_srgb(float v) { return v < 0.0031308f ? v * 12.92f : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f; }
__device__ unsigned char to_byte(float v) { return (unsigned char)(saturate(v) * 255.0f + 0.5f); }

__global__ void tonemap_batch_kernel(TonemapConstants k, bpt_tonemap_settings s, int64_t n, const float* __restrict__ in, float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float3 c = tonemap(k, s, f3(in[3 * i], in[3 * i + 1], in[3 * i + 2]));
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
}

template <bool RGBA8>
__global__ void resolve_tonemapped_kernel(TonemapConstants k, bpt_tonemap_settings s, const double* __restrict__ accum, void* __restrict__ out, int64_t pixel_count) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pixel_count; p += (int64_t)gridDim.x * blockDim.x) {
        const double2* a = reinterpret_cast<const double2*>(accum + 4 * p);
        double2 rg = a[0], bw = a[1];
        double inv = bw.y > 0.0 ? 1.0 / bw.y : 0.0;
        float3 c = tonemap(k, s, f3(float(rg.x * inv), float(rg.y * inv), float(bw.x * inv)));
        if (RGBA8)
            static_cast<uchar4*>(out)[p] = make_uchar4(to_byte(linear_to_srgb(c.x)), to_byte(linear_to_srgb(c.y)), to_byte(linear_to_srgb(c.z)), 255);
        else
            static_cast<float4*>(out)[p] = make_float4(c.x, c.y, c.z, 1.0f);
    }
}

int check_settings(Context* ctx, const bpt_tonemap_settings* s, const char* who) {
    if (!s || s->mode < BPT_TONEMAP_LINEAR || s->mode > BPT_TONEMAP_KHRONOS_NEUTRAL || !(s->exposure > 0.0f))
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, std::string(who) + ": bad tonemapping settings");
    return BPT_OK;
}

} // namespace

int tonemap_batch(Context* ctx, const bpt_tonemap_settings* settings, int64_t n, const float* rgb_in, float* rgb_out) {
    if (int status = check_settings(ctx, settings, "bpt_tonemap_colors")) return status;
    if (n < 0 || !rgb_in || !rgb_out) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_tonemap_colors: bad arguments");
    if (n == 0) return BPT_OK;
    cudaStream_t st = ctx->stream;
    float *d_in = nullptr, *d_out = nullptr;
    BPT_CUDA_CHECK(ctx, cudaMallocAsync((void**)&d_in, 3 * n * sizeof(float), st));
    BPT_CUDA_CHECK(ctx, cudaMallocAsync((void**)&d_out, 3 * n * sizeof(float), st));
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(d_in, rgb_in, 3 * n * sizeof(float), cudaMemcpyHostToDevice, st));
    int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
    tonemap_batch_kernel<<<grid, 256, 0, st>>>(make_constants(), *settings, n, d_in, d_out);
    ctx->counters.kernel_launches++;
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(rgb_out, d_out, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, st));
    BPT_CUDA_CHECK(ctx, cudaFreeAsync(d_in, st)); BPT_CUDA_CHECK(ctx, cudaFreeAsync(d_out, st));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

int resolve_tonemapped(Context* ctx, const bpt_tonemap_settings* settings, void* out, int output_format) {
    if (int status = check_settings(ctx, settings, "bpt_resolve_tonemapped")) return status;
    if (!out || !ctx->accumulation.ptr) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_resolve_tonemapped: nothing rendered");
    if (output_format != BPT_OUTPUT_FLOAT4 && output_format != BPT_OUTPUT_SRGB_RGBA8)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_resolve_tonemapped: unknown output format");
    const int64_t pixels = (int64_t)ctx->width * ctx->height;
    const size_t bytes = pixels * (output_format == BPT_OUTPUT_FLOAT4 ? sizeof(float4) : sizeof(uchar4));
    cudaStream_t st = ctx->stream;
    void* d = nullptr;
    BPT_CUDA_CHECK(ctx, cudaMallocAsync(&d, bytes, st));
    if (output_format == BPT_OUTPUT_FLOAT4)
        resolve_tonemapped_kernel<false><<<ctx->sm_count * 8, 256, 0, st>>>(make_constants(), *settings, ctx->accumulation.ptr, d, pixels);
    else
        resolve_tonemapped_kernel<true><<<ctx->sm_count * 8, 256, 0, st>>>(make_constants(), *settings, ctx->accumulation.ptr, d, pixels);
    ctx->counters.kernel_launches++;
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, st));
    BPT_CUDA_CHECK(ctx, cudaFreeAsync(d, st));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

} // namespace bpt
