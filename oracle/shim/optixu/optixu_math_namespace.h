// Stand-in for the OptiX 6.5 SDK header <optixu/optixu_math_namespace.h>.
//
// TEST INFRASTRUCTURE ONLY. The OptiX SDK is not available in this image; the reference's
// shading / light / RNG headers only need its small vector-math vocabulary to be compiled
// for the host (as the reference's own OptiXRendererTests does). This file is our own
// implementation of that vocabulary on top of CUDA's <vector_types.h>; it is not OptiX and
// calls no OptiX. Semantics that matter for floating point parity (documented OptiX behaviour):
//   normalize(v)   = v * (1 / sqrtf(dot(v, v)))
//   v / s          = v * (1 / s)
//   lerp(a, b, t)  = a + t * (b - a)
//   reflect(i, n)  = i - 2 * n * dot(n, i)
//   refract        = see below; pinned by the reference's MiscTest.h:292-325.
#ifndef BPT_ORACLE_OPTIXU_MATH_NAMESPACE_H
#define BPT_ORACLE_OPTIXU_MATH_NAMESPACE_H

#include <cuda_runtime.h>
#include <vector_functions.h>
#include <vector_types.h>

#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <cstdlib>

#define OPTIXU_INLINE inline

// ------------------------------------------------------------------------------------------------
// Global-namespace operators for the CUDA vector types.
// ------------------------------------------------------------------------------------------------

// float2
OPTIXU_INLINE float2 operator-(float2 a) { return make_float2(-a.x, -a.y); }
OPTIXU_INLINE float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
OPTIXU_INLINE float2 operator+(float2 a, float b) { return make_float2(a.x + b, a.y + b); }
OPTIXU_INLINE float2 operator+(float a, float2 b) { return make_float2(a + b.x, a + b.y); }
OPTIXU_INLINE void operator+=(float2& a, float2 b) { a.x += b.x; a.y += b.y; }
OPTIXU_INLINE float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
OPTIXU_INLINE float2 operator-(float2 a, float b) { return make_float2(a.x - b, a.y - b); }
OPTIXU_INLINE float2 operator-(float a, float2 b) { return make_float2(a - b.x, a - b.y); }
OPTIXU_INLINE void operator-=(float2& a, float2 b) { a.x -= b.x; a.y -= b.y; }
OPTIXU_INLINE float2 operator*(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
OPTIXU_INLINE float2 operator*(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
OPTIXU_INLINE float2 operator*(float s, float2 a) { return make_float2(a.x * s, a.y * s); }
OPTIXU_INLINE void operator*=(float2& a, float2 s) { a.x *= s.x; a.y *= s.y; }
OPTIXU_INLINE void operator*=(float2& a, float s) { a.x *= s; a.y *= s; }
OPTIXU_INLINE float2 operator/(float2 a, float2 b) { return make_float2(a.x / b.x, a.y / b.y); }
OPTIXU_INLINE float2 operator/(float2 a, float s) { float inv = 1.0f / s; return a * inv; }
OPTIXU_INLINE float2 operator/(float s, float2 a) { return make_float2(s / a.x, s / a.y); }
OPTIXU_INLINE void operator/=(float2& a, float s) { float inv = 1.0f / s; a *= inv; }

// float3
OPTIXU_INLINE float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
OPTIXU_INLINE float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
OPTIXU_INLINE float3 operator+(float3 a, float b) { return make_float3(a.x + b, a.y + b, a.z + b); }
OPTIXU_INLINE float3 operator+(float a, float3 b) { return make_float3(a + b.x, a + b.y, a + b.z); }
OPTIXU_INLINE void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
OPTIXU_INLINE float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
OPTIXU_INLINE float3 operator-(float3 a, float b) { return make_float3(a.x - b, a.y - b, a.z - b); }
OPTIXU_INLINE float3 operator-(float a, float3 b) { return make_float3(a - b.x, a - b.y, a - b.z); }
OPTIXU_INLINE void operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
OPTIXU_INLINE float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
OPTIXU_INLINE float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
OPTIXU_INLINE float3 operator*(float s, float3 a) { return make_float3(a.x * s, a.y * s, a.z * s); }
OPTIXU_INLINE void operator*=(float3& a, float3 s) { a.x *= s.x; a.y *= s.y; a.z *= s.z; }
OPTIXU_INLINE void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
OPTIXU_INLINE float3 operator/(float3 a, float3 b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
OPTIXU_INLINE float3 operator/(float3 a, float s) { float inv = 1.0f / s; return a * inv; }
OPTIXU_INLINE float3 operator/(float s, float3 a) { return make_float3(s / a.x, s / a.y, s / a.z); }
OPTIXU_INLINE void operator/=(float3& a, float s) { float inv = 1.0f / s; a *= inv; }

// float4
OPTIXU_INLINE float4 operator-(float4 a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }
OPTIXU_INLINE float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
OPTIXU_INLINE float4 operator+(float4 a, float b) { return make_float4(a.x + b, a.y + b, a.z + b, a.w + b); }
OPTIXU_INLINE float4 operator+(float a, float4 b) { return make_float4(a + b.x, a + b.y, a + b.z, a + b.w); }
OPTIXU_INLINE void operator+=(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
OPTIXU_INLINE float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
OPTIXU_INLINE float4 operator-(float4 a, float b) { return make_float4(a.x - b, a.y - b, a.z - b, a.w - b); }
OPTIXU_INLINE float4 operator-(float a, float4 b) { return make_float4(a - b.x, a - b.y, a - b.z, a - b.w); }
OPTIXU_INLINE void operator-=(float4& a, float4 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }
OPTIXU_INLINE float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
OPTIXU_INLINE float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
OPTIXU_INLINE float4 operator*(float s, float4 a) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
OPTIXU_INLINE void operator*=(float4& a, float4 s) { a.x *= s.x; a.y *= s.y; a.z *= s.z; a.w *= s.w; }
OPTIXU_INLINE void operator*=(float4& a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }
OPTIXU_INLINE float4 operator/(float4 a, float4 b) { return make_float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
OPTIXU_INLINE float4 operator/(float4 a, float s) { float inv = 1.0f / s; return a * inv; }
OPTIXU_INLINE float4 operator/(float s, float4 a) { return make_float4(s / a.x, s / a.y, s / a.z, s / a.w); }
OPTIXU_INLINE void operator/=(float4& a, float s) { float inv = 1.0f / s; a *= inv; }

// int3 / uint3 (only what the reference headers touch)
OPTIXU_INLINE int3 operator-(int3 a) { return make_int3(-a.x, -a.y, -a.z); }

// ---- functions (global namespace, re-exported into namespace optix below) ----
// Scalar helpers (OptiX exposes these in its namespace as well).
OPTIXU_INLINE float lerp(float a, float b, float t) { return a + t * (b - a); }
OPTIXU_INLINE float clamp(float f, float a, float b) { return ::fmaxf(a, ::fminf(f, b)); }
OPTIXU_INLINE int float_as_int(float f) { union { float f; int i; } u; u.f = f; return u.i; }
OPTIXU_INLINE float int_as_float(int i) { union { float f; int i; } u; u.i = i; return u.f; }

// make_* conversions.
OPTIXU_INLINE float2 make_float2(float s) { return ::make_float2(s, s); }
OPTIXU_INLINE float2 make_float2(int2 v) { return ::make_float2(float(v.x), float(v.y)); }
OPTIXU_INLINE float2 make_float2(uint2 v) { return ::make_float2(float(v.x), float(v.y)); }
OPTIXU_INLINE float2 make_float2(float3 v) { return ::make_float2(v.x, v.y); }
OPTIXU_INLINE float2 make_float2(float4 v) { return ::make_float2(v.x, v.y); }
OPTIXU_INLINE float3 make_float3(float s) { return ::make_float3(s, s, s); }
OPTIXU_INLINE float3 make_float3(float2 v, float z) { return ::make_float3(v.x, v.y, z); }
OPTIXU_INLINE float3 make_float3(float x, float2 v) { return ::make_float3(x, v.x, v.y); }
OPTIXU_INLINE float3 make_float3(float4 v) { return ::make_float3(v.x, v.y, v.z); }
OPTIXU_INLINE float3 make_float3(int3 v) { return ::make_float3(float(v.x), float(v.y), float(v.z)); }
OPTIXU_INLINE float3 make_float3(uint3 v) { return ::make_float3(float(v.x), float(v.y), float(v.z)); }
OPTIXU_INLINE float4 make_float4(float s) { return ::make_float4(s, s, s, s); }
OPTIXU_INLINE float4 make_float4(float3 v, float w) { return ::make_float4(v.x, v.y, v.z, w); }
OPTIXU_INLINE float4 make_float4(float2 a, float2 b) { return ::make_float4(a.x, a.y, b.x, b.y); }
OPTIXU_INLINE float4 make_float4(float2 a, float z, float w) { return ::make_float4(a.x, a.y, z, w); }
OPTIXU_INLINE float4 make_float4(int4 v) { return ::make_float4(float(v.x), float(v.y), float(v.z), float(v.w)); }
OPTIXU_INLINE float4 make_float4(uint4 v) { return ::make_float4(float(v.x), float(v.y), float(v.z), float(v.w)); }
OPTIXU_INLINE int3 make_int3(float3 v) { return ::make_int3(int(v.x), int(v.y), int(v.z)); }
OPTIXU_INLINE int3 make_int3(int s) { return ::make_int3(s, s, s); }
OPTIXU_INLINE uint3 make_uint3(float3 v) { return ::make_uint3((unsigned int)v.x, (unsigned int)v.y, (unsigned int)v.z); }

// float2 functions
OPTIXU_INLINE float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
OPTIXU_INLINE float length(float2 v) { return sqrtf(dot(v, v)); }
OPTIXU_INLINE float2 normalize(float2 v) { float inv_len = 1.0f / sqrtf(dot(v, v)); return v * inv_len; }
OPTIXU_INLINE float2 lerp(float2 a, float2 b, float t) { return a + t * (b - a); }
OPTIXU_INLINE float2 fminf(float2 a, float2 b) { return ::make_float2(::fminf(a.x, b.x), ::fminf(a.y, b.y)); }
OPTIXU_INLINE float2 fmaxf(float2 a, float2 b) { return ::make_float2(::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y)); }
OPTIXU_INLINE float2 floor(float2 v) { return ::make_float2(::floorf(v.x), ::floorf(v.y)); }
OPTIXU_INLINE float2 clamp(float2 v, float a, float b) { return ::make_float2(clamp(v.x, a, b), clamp(v.y, a, b)); }

// float3 functions
OPTIXU_INLINE float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
OPTIXU_INLINE float3 cross(float3 a, float3 b) {
    return ::make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
OPTIXU_INLINE float length(float3 v) { return sqrtf(dot(v, v)); }
OPTIXU_INLINE float3 normalize(float3 v) { float inv_len = 1.0f / sqrtf(dot(v, v)); return v * inv_len; }
OPTIXU_INLINE float3 lerp(float3 a, float3 b, float t) { return a + t * (b - a); }
OPTIXU_INLINE float3 fminf(float3 a, float3 b) { return ::make_float3(::fminf(a.x, b.x), ::fminf(a.y, b.y), ::fminf(a.z, b.z)); }
OPTIXU_INLINE float3 fmaxf(float3 a, float3 b) { return ::make_float3(::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y), ::fmaxf(a.z, b.z)); }
OPTIXU_INLINE float fminf(float3 a) { return ::fminf(::fminf(a.x, a.y), a.z); }
OPTIXU_INLINE float fmaxf(float3 a) { return ::fmaxf(::fmaxf(a.x, a.y), a.z); }
OPTIXU_INLINE float3 floor(float3 v) { return ::make_float3(::floorf(v.x), ::floorf(v.y), ::floorf(v.z)); }
OPTIXU_INLINE float3 clamp(float3 v, float a, float b) { return ::make_float3(clamp(v.x, a, b), clamp(v.y, a, b), clamp(v.z, a, b)); }
OPTIXU_INLINE float3 reflect(float3 i, float3 n) { return i - 2.0f * n * dot(n, i); }
OPTIXU_INLINE float3 faceforward(float3 n, float3 i, float3 nref) { return n * copysignf(1.0f, dot(i, nref)); }

// Refraction of the incident direction i about the normal n; ior is n_inside / n_outside.
// Returns false (and a zero vector) on total internal reflection.
OPTIXU_INLINE bool refract(float3& r, float3 i, float3 n, float ior) {
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
        r = ::make_float3(0.0f, 0.0f, 0.0f);
        return false;
    }
    r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
    return true;
}

// float4 functions
OPTIXU_INLINE float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
OPTIXU_INLINE float length(float4 v) { return sqrtf(dot(v, v)); }
OPTIXU_INLINE float4 normalize(float4 v) { float inv_len = 1.0f / sqrtf(dot(v, v)); return v * inv_len; }
OPTIXU_INLINE float4 lerp(float4 a, float4 b, float t) { return a + t * (b - a); }
OPTIXU_INLINE float4 fminf(float4 a, float4 b) { return ::make_float4(::fminf(a.x, b.x), ::fminf(a.y, b.y), ::fminf(a.z, b.z), ::fminf(a.w, b.w)); }
OPTIXU_INLINE float4 fmaxf(float4 a, float4 b) { return ::make_float4(::fmaxf(a.x, b.x), ::fmaxf(a.y, b.y), ::fmaxf(a.z, b.z), ::fmaxf(a.w, b.w)); }
OPTIXU_INLINE float4 floor(float4 v) { return ::make_float4(::floorf(v.x), ::floorf(v.y), ::floorf(v.z), ::floorf(v.w)); }

namespace optix {

using ::float2; using ::float3; using ::float4;
using ::double2; using ::double3; using ::double4;
using ::int2; using ::int3; using ::int4;
using ::uint2; using ::uint3; using ::uint4;
using ::short2; using ::ushort2; using ::ushort4; using ::uchar4; using ::char4;
typedef unsigned int uint;
struct size_t2 { size_t x, y; };

using ::make_float2; using ::make_float3; using ::make_float4;
using ::make_int2; using ::make_int3; using ::make_uint2; using ::make_uint3; using ::make_uint4;
using ::make_short2; using ::make_uchar4; using ::make_ushort4;
using ::make_double3; using ::make_double4;

using ::fminf;
using ::fmaxf;
using ::lerp;
using ::clamp;
using ::float_as_int;
using ::int_as_float;
using ::make_float2;
using ::make_float3;
using ::make_float4;
using ::make_int3;
using ::make_uint3;
using ::dot;
using ::length;
using ::normalize;
using ::floor;
using ::cross;
using ::reflect;
using ::faceforward;
using ::refract;

// Ray, as declared by the OptiX device headers.
struct Ray {
    float3 origin;
    float3 direction;
    unsigned int ray_type;
    float tmin;
    float tmax;

    Ray() {}
    Ray(float3 origin, float3 direction, unsigned int ray_type, float tmin, float tmax = 1e16f)
        : origin(origin), direction(direction), ray_type(ray_type), tmin(tmin), tmax(tmax) {}
};

} // namespace optix

// Dummy bindless buffer id so the POD structs in Types.h keep their size on the host.
template <typename T, int Dim = 1>
struct rtBufferId {
    int m_id;
    rtBufferId() = default;
    rtBufferId(int id) : m_id(id) {}
};

#ifndef RT_EXCEPTION_USER
#define RT_EXCEPTION_USER 0x400
#endif

#include <optixu/optixpp_namespace.h>

#endif // BPT_ORACLE_OPTIXU_MATH_NAMESPACE_H
