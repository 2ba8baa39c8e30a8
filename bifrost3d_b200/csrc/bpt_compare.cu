// Image comparison on the device: the three metrics of the reference's ImageOperations extension
// (extensions/ImageOperations/ImageOperations/Compare.h): rms (:23-44), ssim (:88-118) and mssim (:123-180), restated as CUDA
// kernels over float4 images. SURVEY.md 8(f) row 4 ("tonemap + image output + compare"): used to compare renders against
// reference images and as the second opinion beside the relMSE of the parity tests.
//
// Arithmetic follows the header: per pixel differences and the luminance weights (Math/Color.h:391-393) in fp32, every sum in
// fp64 (Compare.h:31, Statistics :46-83). Sums are reduced in a different order than the reference's sequential loops, so
// results agree to fp64 rounding (tests: 1e-6 relative), not bit for bit. Two properties of the header are kept as they are:
// the window of mssim is [p - support, p + support) - exclusive at the upper end - and its Gaussian weight is
// exp(+d^2 / (2 sigma^2)) / sqrt(2 pi sigma^2) with a POSITIVE exponent (:150-152).
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#include "bpt_context.h"
#include "bpt_math.cuh"
#include "../../include/bpt_c_api.h"

namespace bpt {
namespace {

constexpr int COMPARE_BLOCK = 256;
constexpr int STAT_COUNT = 16; // 15 sums + the squared luminance error

__device__ __forceinline__ float luminance(float r, float g, float b) { return 0.2126f * r + 0.7152f * g + 0.0722f * b; }

// Statistics::add with weight 1 for every pixel plus the rms term; one partial result per block, summed by finish_kernel.
__global__ void __launch_bounds__(COMPARE_BLOCK) global_statistics_kernel(const float4* __restrict__ reference, const float4* __restrict__ target, int64_t pixel_count,
                                                                         double* __restrict__ partial, float4* __restrict__ diff) {
    double s[STAT_COUNT];
#pragma unroll
    for (int k = 0; k < STAT_COUNT; ++k) s[k] = 0.0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pixel_count; p += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = reference[p], b = target[p];
        const float er = fabsf(a.x - b.x), eg = fabsf(a.y - b.y), eb = fabsf(a.z - b.z);
        const float l1 = luminance(er, eg, eb);
        if (diff) diff[p] = make_float4(er, eg, eb, 1.0f);
        const double ra[3] = { a.x, a.y, a.z }, tb[3] = { b.x, b.y, b.z };
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s[c] += ra[c]; s[3 + c] += ra[c] * ra[c]; s[6 + c] += tb[c]; s[9 + c] += tb[c] * tb[c]; s[12 + c] += ra[c] * tb[c];
        }
        s[15] += (double)(l1 * l1);
    }
    __shared__ double warp_sums[COMPARE_BLOCK / 32][STAT_COUNT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < STAT_COUNT; ++k) {
        double v = s[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) warp_sums[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < STAT_COUNT) {
        double v = 0.0;
        for (int w = 0; w < COMPARE_BLOCK / 32; ++w) v += warp_sums[w][threadIdx.x];
        partial[(int64_t)blockIdx.x * STAT_COUNT + threadIdx.x] = v;
    }
}

// SSIM of one set of statistics, Compare.h:104-117 (algorithm (13) of Wang et al. with C1 = 0.01, C2 = 0.03).
__device__ __forceinline__ void ssim_of(const double* s, double weight, float out_rgb[3]) {
    const double C1 = 0.01, C2 = 0.03;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double reference_mean = s[c] / weight, target_mean = s[6 + c] / weight;
        const double reference_variance = s[3 + c] / weight - reference_mean * reference_mean;
        const double target_variance = s[9 + c] / weight - target_mean * target_mean;
        const double covariance = s[12 + c] / weight - s[c] * s[6 + c] / (weight * weight);
        const double ssim = (2.0 * reference_mean * target_mean + C1) * (2.0 * covariance + C2) /
                            ((reference_mean * reference_mean + target_mean * target_mean + C1) * (reference_variance + target_variance + C2));
        out_rgb[c] = float(ssim);
    }
}

// mssim, Compare.h:123-180: the SSIM of the weighted window around every pixel; one thread per pixel, block partial sums.
__global__ void __launch_bounds__(COMPARE_BLOCK) windowed_ssim_kernel(const float4* __restrict__ reference, const float4* __restrict__ target, int width, int height,
                                                                     int support, double* __restrict__ partial, float4* __restrict__ diff) {
    double local = 0.0;
    const int64_t pixel_count = (int64_t)width * height;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixel_count; i += (int64_t)gridDim.x * blockDim.x) {
        const int xx = int(i % width), yy = int(i / width);
        const int y_start = max(yy - support, 0), y_end = min(yy + support, height);
        const int x_start = max(xx - support, 0), x_end = min(xx + support, width);
        double s[15];
#pragma unroll
        for (int k = 0; k < 15; ++k) s[k] = 0.0;
        double summed_weight = 0.0;
        for (int y = y_start; y < y_end; ++y)
            for (int x = x_start; x < x_end; ++x) {
                const float dx = float(x - xx) / float(support), dy = float(y - yy) / float(support);
                const float distance_squared = dx * dx + dy * dy;
                const float weight_variance = 1.5f * 1.5f;
                const double weight = expf(distance_squared / (2.0f * weight_variance)) / sqrtf(2.0f * PI_F * weight_variance);
                const float4 a = reference[(int64_t)y * width + x], b = target[(int64_t)y * width + x];
                const double ra[3] = { a.x, a.y, a.z }, tb[3] = { b.x, b.y, b.z };
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    s[c] += weight * ra[c]; s[3 + c] += weight * ra[c] * ra[c]; s[6 + c] += weight * tb[c]; s[9 + c] += weight * tb[c] * tb[c];
                    s[12 + c] += weight * ra[c] * tb[c];
                }
                summed_weight += weight;
            }
        float ssim_rgb[3];
        ssim_of(s, summed_weight, ssim_rgb);
        local += (double)luminance(ssim_rgb[0], ssim_rgb[1], ssim_rgb[2]);
        if (diff) diff[i] = make_float4(1.0f - ssim_rgb[0], 1.0f - ssim_rgb[1], 1.0f - ssim_rgb[2], 1.0f);
    }
    __shared__ double warp_sums[COMPARE_BLOCK / 32];
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < COMPARE_BLOCK / 32; ++w) v += warp_sums[w];
        partial[blockIdx.x] = v;
    }
}

// results: [0] rms, [1] ssim, [2] mssim
__global__ void finish_kernel(const double* __restrict__ global_partial, int global_blocks, const double* __restrict__ window_partial, int window_blocks,
                              int64_t pixel_count, float* __restrict__ results) {
    double s[STAT_COUNT];
    for (int k = 0; k < STAT_COUNT; ++k) { s[k] = 0.0; for (int b = 0; b < global_blocks; ++b) s[k] += global_partial[(int64_t)b * STAT_COUNT + k]; }
    results[0] = sqrtf(float(s[15] / double(pixel_count)));
    float ssim_rgb[3];
    ssim_of(s, double(pixel_count), ssim_rgb);
    results[1] = luminance(ssim_rgb[0], ssim_rgb[1], ssim_rgb[2]);
    double m = 0.0;
    for (int b = 0; b < window_blocks; ++b) m += window_partial[b];
    results[2] = window_blocks ? float(m / double(pixel_count)) : 0.0f;
}

} // namespace
} // namespace bpt

using namespace bpt;

extern "C" int bpt_compare_images(bpt_ctx* c, int width, int height, const float* reference_rgba, const float* target_rgba, int mssim_support,
                                  float* out_rms, float* out_ssim, float* out_mssim, float* out_rms_diff_rgba, float* out_mssim_diff_rgba) {
    Context* ctx = as_context(c);
    if (width <= 0 || height <= 0 || !reference_rgba || !target_rgba || mssim_support < 0)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_compare_images: bad arguments");
    if (mssim_support == 0 && (out_mssim || out_mssim_diff_rgba)) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_compare_images: mssim needs a support > 0");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    const int64_t pixels = (int64_t)width * height;
    const int blocks = ctx->sm_count * 4;
    float4 *d_reference = nullptr, *d_target = nullptr, *d_diff = nullptr, *d_window_diff = nullptr;
    double *d_partial = nullptr, *d_window_partial = nullptr;
    float* d_results = nullptr;
    auto release = [&]() {
        for (void* p : { (void*)d_reference, (void*)d_target, (void*)d_diff, (void*)d_window_diff, (void*)d_partial, (void*)d_window_partial, (void*)d_results })
            if (p) cudaFreeAsync(p, st);
        cudaStreamSynchronize(st);
    };
#define COMPARE_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { release(); return ctx->cuda_fail(_e, #expr); } } while (0)
    COMPARE_CHECK(cudaMallocAsync((void**)&d_reference, pixels * sizeof(float4), st));
    COMPARE_CHECK(cudaMallocAsync((void**)&d_target, pixels * sizeof(float4), st));
    COMPARE_CHECK(cudaMallocAsync((void**)&d_partial, (size_t)blocks * STAT_COUNT * sizeof(double), st));
    COMPARE_CHECK(cudaMallocAsync((void**)&d_window_partial, (size_t)blocks * sizeof(double), st));
    COMPARE_CHECK(cudaMallocAsync((void**)&d_results, 3 * sizeof(float), st));
    if (out_rms_diff_rgba) COMPARE_CHECK(cudaMallocAsync((void**)&d_diff, pixels * sizeof(float4), st));
    if (out_mssim_diff_rgba) COMPARE_CHECK(cudaMallocAsync((void**)&d_window_diff, pixels * sizeof(float4), st));
    COMPARE_CHECK(cudaMemcpyAsync(d_reference, reference_rgba, pixels * sizeof(float4), cudaMemcpyHostToDevice, st));
    COMPARE_CHECK(cudaMemcpyAsync(d_target, target_rgba, pixels * sizeof(float4), cudaMemcpyHostToDevice, st));
    global_statistics_kernel<<<blocks, COMPARE_BLOCK, 0, st>>>(d_reference, d_target, pixels, d_partial, d_diff);
    const bool windowed = mssim_support > 0 && (out_mssim || out_mssim_diff_rgba);
    if (windowed) windowed_ssim_kernel<<<blocks, COMPARE_BLOCK, 0, st>>>(d_reference, d_target, width, height, mssim_support, d_window_partial, d_window_diff);
    finish_kernel<<<1, 1, 0, st>>>(d_partial, blocks, d_window_partial, windowed ? blocks : 0, pixels, d_results);
    ctx->counters.kernel_launches += windowed ? 3 : 2;
    float results[3] = { 0, 0, 0 };
    COMPARE_CHECK(cudaMemcpyAsync(results, d_results, sizeof(results), cudaMemcpyDeviceToHost, st));
    if (out_rms_diff_rgba) COMPARE_CHECK(cudaMemcpyAsync(out_rms_diff_rgba, d_diff, pixels * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (out_mssim_diff_rgba) COMPARE_CHECK(cudaMemcpyAsync(out_mssim_diff_rgba, d_window_diff, pixels * sizeof(float4), cudaMemcpyDeviceToHost, st));
    COMPARE_CHECK(cudaStreamSynchronize(st));
    COMPARE_CHECK(cudaGetLastError());
#undef COMPARE_CHECK
    release();
    if (out_rms) *out_rms = results[0];
    if (out_ssim) *out_ssim = results[1];
    if (out_mssim) *out_mssim = results[2];
    return BPT_OK;
}
