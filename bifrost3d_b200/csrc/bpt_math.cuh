// Small float3/float4 vocabulary for the device code. IEEE semantics on purpose: the parity
// target is the reference's host-compiled headers (no fast-math, no FMA contraction), so the
// compound operations below fix the same evaluation order those headers get from the OptiX
// math vocabulary: normalize = v * (1/sqrt(dot)), v / s = v * (1/s), lerp = a + t*(b-a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define BPT_HD __host__ __device__ __forceinline__
#define BPT_D __device__ __forceinline__
// Large shading routines that are called from several places: one out-of-line copy keeps the shade kernel's code inside
// the instruction cache (ncu: 'stalled_no_instruction' was the top stall reason with everything inlined).
#define BPT_CALL static __device__ __noinline__
// Routines with a single call site in the render kernels (light sampling, the Sobol sampler) are inlined instead: their
// call frames were half of the surface shade kernel's 512-byte stack, and that local memory misses L1 (measured: shade
// -7 %). Inlining the BSDF sample / evaluate routines as well was measured too: no further gain, cornell -2 %.
#define BPT_CALL1 static __device__ __forceinline__

namespace bpt {

// Correctly rounded division / square root behind one name each (the vector operators and the shading routines use them).
// An out-of-line variant (one call per site instead of ~9 inlined instructions) was measured on B200 in round 2: the shade
// kernel got 6 % SLOWER (materials 1.179 -> 1.252 ms per sample), so the operations stay inlined.
BPT_HD float fdiv(float a, float b) { return a / b; }
BPT_HD float rcp(float s) { return fdiv(1.0f, s); }
BPT_HD float fsqrt(float x) { return sqrtf(x); }

constexpr float PI_F = 3.14159265358979323846f;
constexpr float TWO_PI_F = 6.283185307f;
constexpr float RECIP_PI_F = 0.31830988618379067153776752674503f;

BPT_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
BPT_HD float3 f3(float s) { return make_float3(s, s, s); }
BPT_HD float3 f3(float4 v) { return make_float3(v.x, v.y, v.z); }
BPT_HD float2 f2(float x, float y) { return make_float2(x, y); }
BPT_HD float4 f4(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }

BPT_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
BPT_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
BPT_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
BPT_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
BPT_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
BPT_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
BPT_HD float3 operator+(float3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
BPT_HD float3 operator-(float3 a, float s) { return f3(a.x - s, a.y - s, a.z - s); }
BPT_HD float3 operator-(float s, float3 a) { return f3(s - a.x, s - a.y, s - a.z); }
BPT_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
BPT_HD float3 operator/(float3 a, float s) { float inv = rcp(s); return a * inv; }
BPT_HD float3 operator/(float s, float3 a) { return f3(s / a.x, s / a.y, s / a.z); }
BPT_HD void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
BPT_HD void operator*=(float3& a, float3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
BPT_HD void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
BPT_HD void operator/=(float3& a, float s) { float inv = rcp(s); a *= inv; }

BPT_HD float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
BPT_HD float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
BPT_HD float2 operator*(float2 a, float s) { return f2(a.x * s, a.y * s); }
BPT_HD float2 operator*(float s, float2 a) { return f2(a.x * s, a.y * s); }
BPT_HD float2 operator/(float2 a, float s) { float inv = rcp(s); return a * inv; }

BPT_HD float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
BPT_HD float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
BPT_HD float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

BPT_HD float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
BPT_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
BPT_HD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
BPT_HD float length(float3 v) { return fsqrt(dot(v, v)); }
BPT_HD float length(float2 v) { return fsqrt(dot(v, v)); }
BPT_HD float3 normalize(float3 v) { float inv_len = rcp(fsqrt(dot(v, v))); return v * inv_len; }
BPT_HD float lerp(float a, float b, float t) { return a + t * (b - a); }
BPT_HD float3 lerp(float3 a, float3 b, float t) { return a + t * (b - a); }
BPT_HD float clampf(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }
BPT_HD float saturate(float v) { return clampf(v, 0.0f, 1.0f); }
BPT_HD float3 reflect(float3 i, float3 n) { return i - 2.0f * n * dot(n, i); }
BPT_HD float3 min3(float3 a, float3 b) { return f3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
BPT_HD float3 max3(float3 a, float3 b) { return f3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
BPT_HD float sum(float3 v) { return v.x + v.y + v.z; }
BPT_HD float pow2(float x) { return x * x; }
BPT_HD float pow4(float x) { float xx = x * x; return xx * xx; }
BPT_HD float pow5(float x) { float xx = x * x; return xx * xx * x; }
BPT_HD bool is_black(float3 c) { return c.x <= 0.0f && c.y <= 0.0f && c.z <= 0.0f; }

} // namespace bpt
