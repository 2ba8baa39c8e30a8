// CPU oracle for traversal and the path integrator.
//
// TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load this. The reference's traversal and integrator are OptiX 6.5 programs and cannot be compiled here
// ("parity unpinned by the reference" for these two parts); this file restates them:
//   * integrator: extensions/OptiXRenderer/OptiXRenderer/Shading/MonteCarlo.cu:61-233,278-302,
//                 Shading/SimpleRGPs.cu:44-140,349-362, Shading/TriangleAttributes.cu:35-84,
//                 Shading/LightSources/LightSources.cu:31-70 - followed statement by statement, and CALLING the
//                 staged reference headers for every BSDF, light, RNG, TBN, MIS and ray-offset evaluation.
//   * traversal:  OptiX' closed source Trbvh/RTX traversal is replaced by (a) a brute force loop over all
//                 triangles and (b) a median split BVH, both using the watertight ray/triangle test of
//                 Woop, Benthin, Wald (JCGT 2013) in plain IEEE fp32 and resolving the closest hit as
//                 min (t, global primitive index).
// Scenes are flattened to world space (instance-major primitive order) with the same rounding as the product.
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#include <OptiXRenderer/MonteCarlo.h>
#include <OptiXRenderer/RNG.h>
#include <OptiXRenderer/Intersect.h>
#include <OptiXRenderer/Shading/ShadingModels/DiffuseShading.h>
#include <OptiXRenderer/Shading/BSDFs/GGX.h>
#include <OptiXRenderer/Shading/ShadingModels/Utils.h> // DielectricRho, which TransmissiveShading.h uses without including
#include <OptiXRenderer/Shading/ShadingModels/TransmissiveShading.h>
#include <OptiXRenderer/Shading/LightSources/DirectionalLightImpl.h>
#include <OptiXRenderer/Shading/LightSources/SphereLightImpl.h>
#include <OptiXRenderer/Shading/LightSources/SpotLightImpl.h>
#include <OptiXRenderer/TBN.h>
#include <OptiXRenderer/Types.h>
#include <Bifrost/Assets/Shading/Fittings.h>
#include <Bifrost/Math/OctahedralNormal.h>
#include <sstream>
#include <string>
#include <vector>
#include <iostream>
#include <functional>
#include <map>
#include <algorithm>
#define private public
#include <OptiXRenderer/Shading/ShadingModels/DefaultShading.h>
#undef private

#include <omp.h>
#include <cstdint>
#include <cstring>
#include <cfloat>
#include <cmath>

using namespace OptiXRenderer;
using namespace optix;

namespace {

struct InstanceIn { int32_t mesh_id; int32_t material_id; float to_world[12]; };
struct CameraIn { float view_to_world_rotation[9]; float inverse_projection[16]; float inverse_view_projection[16]; };
struct SettingsIn { uint32_t max_bounce_count; int32_t next_event_sample_count; float path_regularization_pdf_scale; uint32_t russian_roulette_start_bounce; };

struct Mesh {
    std::vector<uint32_t> indices;
    std::vector<float3> positions;
    std::vector<float2> texcoords;         // empty when the mesh has none
    std::vector<float3> emission;          // per-vertex emission scale, empty when the mesh has none
    std::vector<OctahedralNormal> normals; // empty when the mesh has none
    std::vector<uchar4> tints;
};

struct Triangle {
    float3 p0, p1, p2;
    OctahedralNormal n0, n1, n2;
    uchar4 t0, t1, t2;
    float2 uv0, uv1, uv2;
    float3 e0, e1, e2;
    int material;
    int instance;
    bool has_normals, has_tints, has_texcoords, has_emission;
};

// Restatement of the sampler the reference creates for a Bifrost texture (Renderer.cpp:650-751) and reads with rtTex2D
// (Types.h:388-414). OptiX 6.5 hands this to the CUDA texture unit, whose arithmetic is documented in the CUDA C++
// Programming Guide, appendix "Texture Fetching": normalized coordinates x = u * N; wrap: x = frac(u) * N, clamp: x stays
// in [0, N); nearest: T[floor(x)]; linear: xB = x - 0.5, i = floor(xB), alpha = frac(xB) held in 1.8 fixed point,
// (1 - alpha) T[i] + alpha T[i + 1] with i and i + 1 wrapped or clamped. 8-bit texels read as v / 255, sRGB images are
// decoded per texel before filtering (alpha excluded). PARITY UNPINNED by the reference: no reference test samples a texture.
struct Texture {
    int width = 0, height = 0, channels = 0;
    bool linear = true;
    int wrap_u = 1, wrap_v = 1; // 0 clamp, 1 repeat
    std::vector<float4> texels;
};

inline float srgb_to_linear(float c) { return c <= 0.04045f ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f); }

inline int wrap_index(int i, int n, int mode) {
    if (mode == 1) { i %= n; return i < 0 ? i + n : i; }
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

inline float4 texture_sample(const Texture& tex, float u, float v) {
    auto coord = [](float s, int n, int mode) {
        if (mode == 1) s = s - floorf(s);
        float x = s * n;
        if (mode == 0) x = fminf(fmaxf(x, 0.0f), (float)n); // texel indices are clamped below
        return x;
    };
    float x = coord(u, tex.width, tex.wrap_u), y = coord(v, tex.height, tex.wrap_v);
    auto texel = [&](int i, int j) { return tex.texels[(size_t)wrap_index(j, tex.height, tex.wrap_v) * tex.width + wrap_index(i, tex.width, tex.wrap_u)]; };
    if (!tex.linear)
        return texel((int)floorf(x), (int)floorf(y));
    float xb = x - 0.5f, yb = y - 0.5f;
    int i = (int)floorf(xb), j = (int)floorf(yb);
    // Rounded to nearest: measured against B200's texture unit this halves the mean deviation compared with truncation
    // (4.6e-4 against 1.1e-3 of the texel range); what remains (<= 2/256) is the unit's own fixed-point coordinate handling.
    float alpha = floorf((xb - floorf(xb)) * 256.0f + 0.5f) / 256.0f;
    float beta = floorf((yb - floorf(yb)) * 256.0f + 0.5f) / 256.0f;
    float4 t00 = texel(i, j), t10 = texel(i + 1, j), t01 = texel(i, j + 1), t11 = texel(i + 1, j + 1);
    return (1.0f - alpha) * (1.0f - beta) * t00 + alpha * (1.0f - beta) * t10 + (1.0f - alpha) * beta * t01 + alpha * beta * t11;
}

struct Box { float3 lo, hi; };

struct Node {
    Box box[2];
    int child[2]; // >= 0 inner node, < 0: ~first triangle slot
    int count[2]; // 0 inner, > 0 leaf size, -1 absent
};

struct EnvironmentIn {
    std::vector<float4> texels; int width = 0, height = 0;
    std::vector<float> pdf; int pdf_width = 0, pdf_height = 0;
    std::vector<LightSample> samples;
    float3 tint = { 0, 0, 0 };
    std::vector<float> marginal_cdf, conditional_cdf; // Distribution2D CDFs for Light::Environment (CDF inversion per sample)
};

struct Scene {
    std::map<int, Mesh> meshes;
    std::vector<InstanceIn> instances;
    std::vector<Material> materials;
    std::map<int, Texture> textures;
    std::vector<Light> lights; // analytical lights, then (optionally) the environment
    int light_count = 0;
    EnvironmentIn env;
    bool env_in_light_list = false;

    std::vector<Triangle> triangles;          // global (instance-major) order
    std::vector<float> normal_matrices;       // 9 per instance
    std::vector<int> order;                   // BVH slot -> global primitive
    std::vector<Node> nodes;
    float4 nee_offsets[256];
};

// Material::get_coverage / get_tint_roughness / get_metallic (Types.h:388-414, GPU_DEVICE only) with rtTex2D restated by
// texture_sample above.
inline float material_coverage(const Scene& sc, const Material& m, float2 texcoord) {
    float coverage_tex_sample = 1.0f;
    if (m.coverage_texture_ID)
        coverage_tex_sample = texture_sample(sc.textures.at(m.coverage_texture_ID), texcoord.x, texcoord.y).x;
    if (m.is_cutout()) return coverage_tex_sample < m.coverage ? 0.0f : 1.0f;
    return m.coverage * coverage_tex_sample;
}

inline float2 interpolate_texcoord(const Triangle& tri, float bx, float by) {
    // TriangleAttributes.cu:57-64; `texcoord - make_float2(0)` is a no-op there, an untextured mesh reads (0, 0) here.
    if (!tri.has_texcoords) return make_float2(0.0f, 0.0f);
    float bz = 1.0f - bx - by;
    return tri.uv1 * bx + tri.uv2 * by + tri.uv0 * bz;
}

// The material as the shading models see it through get_tint_roughness(texcoord) and get_metallic(texcoord).
inline Material material_at(const Scene& sc, const Material& m, float2 texcoord) {
    Material r = m;
    float4 tint_roughness = make_float4(m.tint, m.roughness);
    if (m.tint_roughness_texture_ID)
        tint_roughness *= texture_sample(sc.textures.at(m.tint_roughness_texture_ID), texcoord.x, texcoord.y);
    if (m.roughness_texture_ID)
        tint_roughness.w *= texture_sample(sc.textures.at(m.roughness_texture_ID), texcoord.x, texcoord.y).x;
    r.tint = make_float3(tint_roughness); r.roughness = tint_roughness.w;
    if (m.metallic_texture_ID)
        r.metallic = m.metallic * texture_sample(sc.textures.at(m.metallic_texture_ID), texcoord.x, texcoord.y).x;
    return r;
}

// ---- watertight ray / triangle (Woop et al. 2013), plain fp32 -----------------------------------
struct Shear { int kx, ky, kz; float Sx, Sy, Sz; };

inline float comp(float3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

inline Shear make_shear(float3 d) {
    Shear s;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    s.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    s.kx = s.kz + 1; if (s.kx == 3) s.kx = 0;
    s.ky = s.kx + 1; if (s.ky == 3) s.ky = 0;
    float dz = comp(d, s.kz);
    if (dz < 0.0f) std::swap(s.kx, s.ky);
    s.Sx = comp(d, s.kx) / dz;
    s.Sy = comp(d, s.ky) / dz;
    s.Sz = 1.0f / dz;
    return s;
}

// volatile stores keep every intermediate rounded to fp32 and rule out contraction whatever the flags.
inline bool watertight(const Shear& s, float3 o, float3 p0, float3 p1, float3 p2, float& t, float& u, float& v) {
    const float3 A = make_float3(p0.x - o.x, p0.y - o.y, p0.z - o.z);
    const float3 B = make_float3(p1.x - o.x, p1.y - o.y, p1.z - o.z);
    const float3 C = make_float3(p2.x - o.x, p2.y - o.y, p2.z - o.z);
    const float Akz = comp(A, s.kz), Bkz = comp(B, s.kz), Ckz = comp(C, s.kz);
    volatile float m;
    m = s.Sx * Akz; const float Ax = comp(A, s.kx) - m;
    m = s.Sy * Akz; const float Ay = comp(A, s.ky) - m;
    m = s.Sx * Bkz; const float Bx = comp(B, s.kx) - m;
    m = s.Sy * Bkz; const float By = comp(B, s.ky) - m;
    m = s.Sx * Ckz; const float Cx = comp(C, s.kx) - m;
    m = s.Sy * Ckz; const float Cy = comp(C, s.ky) - m;

    volatile float a, b;
    a = Cx * By; b = Cy * Bx; float U = a - b;
    a = Ax * Cy; b = Ay * Cx; float V = a - b;
    a = Bx * Ay; b = By * Ax; float W = a - b;

    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        volatile double da, db;
        da = (double)Cx * (double)By; db = (double)Cy * (double)Bx; U = (float)(da - db);
        da = (double)Ax * (double)Cy; db = (double)Ay * (double)Cx; V = (float)(da - db);
        da = (double)Bx * (double)Ay; db = (double)By * (double)Ax; W = (float)(da - db);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
        return false;
    volatile float uv = U + V;
    const float det = uv + W;
    if (det == 0.0f)
        return false;
    const float Az = s.Sz * Akz, Bz = s.Sz * Bkz, Cz = s.Sz * Ckz;
    volatile float ta = U * Az, tb = V * Bz, tc = W * Cz;
    volatile float tab = ta + tb;
    const float T = tab + tc;
    const float rcp_det = 1.0f / det;
    t = T * rcp_det;
    u = V * rcp_det;
    v = W * rcp_det;
    return true;
}

struct HitRecord { float t; int primitive; float u, v; };

inline void test_triangle(const Scene& sc, int gp, const Shear& sh, float3 o, float tmin, HitRecord& hit) {
    const Triangle& tri = sc.triangles[gp];
    float t, u, v;
    if (!watertight(sh, o, tri.p0, tri.p1, tri.p2, t, u, v)) return;
    if (t > tmin && (t < hit.t || (t == hit.t && gp < hit.primitive))) { hit.t = t; hit.primitive = gp; hit.u = u; hit.v = v; }
}

HitRecord closest_brute(const Scene& sc, float3 o, float3 d, float tmin, float tmax) {
    HitRecord hit = { tmax, 0x7fffffff, 0, 0 };
    Shear sh = make_shear(d);
    for (int gp = 0; gp < (int)sc.triangles.size(); ++gp) test_triangle(sc, gp, sh, o, tmin, hit);
    if (hit.primitive == 0x7fffffff) hit.primitive = -1;
    return hit;
}

inline bool slab(const Box& b, float3 o, float3 inv_d, float tmin, float tmax, float& tn_out) {
    float t0x = (b.lo.x - o.x) * inv_d.x, t1x = (b.hi.x - o.x) * inv_d.x;
    float t0y = (b.lo.y - o.y) * inv_d.y, t1y = (b.hi.y - o.y) * inv_d.y;
    float t0z = (b.lo.z - o.z) * inv_d.z, t1z = (b.hi.z - o.z) * inv_d.z;
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    tn = fmaxf(tmin, tn * 0.999999f);
    tf = fminf(tmax, tf * 1.000001f);
    tn_out = tn;
    return tn <= tf;
}

HitRecord closest_bvh(const Scene& sc, float3 o, float3 d, float tmin, float tmax) {
    HitRecord hit = { tmax, 0x7fffffff, 0, 0 };
    Shear sh = make_shear(d);
    float3 inv_d = make_float3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128]; int sp = 0;
    int node = 0;
    while (true) {
        const Node& n = sc.nodes[node];
        int next[2]; float tn[2]; int nn = 0;
        for (int k = 0; k < 2; ++k) {
            if (n.count[k] < 0) continue;
            float tnear;
            if (!slab(n.box[k], o, inv_d, tmin, hit.t, tnear)) continue;
            if (n.count[k] > 0) {
                for (int i = 0; i < n.count[k]; ++i) test_triangle(sc, sc.order[~n.child[k] + i], sh, o, tmin, hit);
            } else { next[nn] = n.child[k]; tn[nn] = tnear; ++nn; }
        }
        if (nn == 2) {
            int first = tn[0] <= tn[1] ? 0 : 1;
            stack[sp++] = next[1 - first];
            node = next[first];
        } else if (nn == 1) node = next[0];
        else { if (sp == 0) break; node = stack[--sp]; }
    }
    if (hit.primitive == 0x7fffffff) hit.primitive = -1;
    return hit;
}

// Product of (1 - coverage) over every triangle hit in (tmin, tmax). shadow_any_hit (MonteCarlo.cu:278-285) multiplies the
// light sample's radiance by (1 - coverage) per surface and terminates the ray once every channel is below 1e-7;
// `largest_radiance_channel` is max(r, g, b) of that radiance, so the rule here is the same up to rounding.
float transmission_bvh(const Scene& sc, float3 o, float3 d, float tmin, float tmax, bool brute, float largest_radiance_channel = 1.0f) {
    Shear sh = make_shear(d);
    float transmission = 1.0f;
    auto visit = [&](int gp) -> bool {
        const Triangle& tri = sc.triangles[gp];
        float t, u, v;
        if (!watertight(sh, o, tri.p0, tri.p1, tri.p2, t, u, v)) return false;
        if (!(t > tmin && t < tmax)) return false;
        const Material& m = sc.materials[tri.material];
        float coverage = material_coverage(sc, m, interpolate_texcoord(tri, u, v));
        transmission *= 1.0f - coverage;
        if (transmission * largest_radiance_channel < 0.0000001f) { transmission = 0.0f; return true; }
        return false;
    };
    if (brute) {
        for (int gp = 0; gp < (int)sc.triangles.size(); ++gp) if (visit(gp)) break;
        return transmission;
    }
    float3 inv_d = make_float3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[128]; int sp = 0; int node = 0;
    while (true) {
        const Node& n = sc.nodes[node];
        int next[2]; int nn = 0;
        for (int k = 0; k < 2; ++k) {
            if (n.count[k] < 0) continue;
            float tnear;
            if (!slab(n.box[k], o, inv_d, tmin, tmax, tnear)) continue;
            if (n.count[k] > 0) {
                for (int i = 0; i < n.count[k]; ++i) if (visit(sc.order[~n.child[k] + i])) return 0.0f;
            } else next[nn++] = n.child[k];
        }
        if (nn == 2) { stack[sp++] = next[1]; node = next[0]; }
        else if (nn == 1) node = next[0];
        else { if (sp == 0) break; node = stack[--sp]; }
    }
    return transmission;
}

// ---- scene flattening and BVH --------------------------------------------------------------------
inline float3 transform_point(const float* m, float3 p) {
    volatile float a, b, c;
    float3 r;
    a = m[0] * p.x; b = m[1] * p.y; c = m[2] * p.z; { volatile float ab = a + b; volatile float abc = ab + c; r.x = abc + m[3]; }
    a = m[4] * p.x; b = m[5] * p.y; c = m[6] * p.z; { volatile float ab = a + b; volatile float abc = ab + c; r.y = abc + m[7]; }
    a = m[8] * p.x; b = m[9] * p.y; c = m[10] * p.z; { volatile float ab = a + b; volatile float abc = ab + c; r.z = abc + m[11]; }
    return r;
}

void flatten(Scene& sc) {
    sc.triangles.clear(); sc.normal_matrices.clear();
    int instance_index = 0;
    for (const InstanceIn& inst : sc.instances) {
        const Mesh& mesh = sc.meshes.at(inst.mesh_id);
        int prims = (int)mesh.indices.size() / 3;
        if (prims == 0) continue;
        for (int p = 0; p < prims; ++p) {
            uint32_t i0 = mesh.indices[3 * p], i1 = mesh.indices[3 * p + 1], i2 = mesh.indices[3 * p + 2];
            Triangle t = {};
            t.p0 = transform_point(inst.to_world, mesh.positions[i0]);
            t.p1 = transform_point(inst.to_world, mesh.positions[i1]);
            t.p2 = transform_point(inst.to_world, mesh.positions[i2]);
            t.has_normals = !mesh.normals.empty(); t.has_tints = !mesh.tints.empty();
            if (t.has_normals) { t.n0 = mesh.normals[i0]; t.n1 = mesh.normals[i1]; t.n2 = mesh.normals[i2]; }
            if (t.has_tints) { t.t0 = mesh.tints[i0]; t.t1 = mesh.tints[i1]; t.t2 = mesh.tints[i2]; }
            t.has_emission = !mesh.emission.empty();
            if (t.has_emission) { t.e0 = mesh.emission[i0]; t.e1 = mesh.emission[i1]; t.e2 = mesh.emission[i2]; }
            t.has_texcoords = !mesh.texcoords.empty();
            if (t.has_texcoords) { t.uv0 = mesh.texcoords[i0]; t.uv1 = mesh.texcoords[i1]; t.uv2 = mesh.texcoords[i2]; }
            t.material = inst.material_id; t.instance = instance_index;
            sc.triangles.push_back(t);
        }
        const float* m = inst.to_world;
        double a[9] = { m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10] };
        double c[9] = { a[4] * a[8] - a[5] * a[7], a[5] * a[6] - a[3] * a[8], a[3] * a[7] - a[4] * a[6],
                        a[2] * a[7] - a[1] * a[8], a[0] * a[8] - a[2] * a[6], a[1] * a[6] - a[0] * a[7],
                        a[1] * a[5] - a[2] * a[4], a[2] * a[3] - a[0] * a[5], a[0] * a[4] - a[1] * a[3] };
        double det = a[0] * c[0] + a[1] * c[1] + a[2] * c[2];
        for (int k = 0; k < 9; ++k) sc.normal_matrices.push_back(float(det != 0.0 ? c[k] / det : (k % 4 == 0 ? 1.0 : 0.0)));
        ++instance_index;
    }
}

inline Box tri_box(const Triangle& t) {
    Box b;
    b.lo = make_float3(fminf(fminf(t.p0.x, t.p1.x), t.p2.x), fminf(fminf(t.p0.y, t.p1.y), t.p2.y), fminf(fminf(t.p0.z, t.p1.z), t.p2.z));
    b.hi = make_float3(fmaxf(fmaxf(t.p0.x, t.p1.x), t.p2.x), fmaxf(fmaxf(t.p0.y, t.p1.y), t.p2.y), fmaxf(fmaxf(t.p0.z, t.p1.z), t.p2.z));
    return b;
}
inline Box merge(const Box& a, const Box& b) {
    return { make_float3(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z)),
             make_float3(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z)) };
}

// Median split on the longest centroid axis; leaves of <= 4 triangles. Returns the subtree's box.
Box build_range(Scene& sc, std::vector<float3>& centroids, int first, int last, int& out_child, int& out_count) {
    int size = last - first;
    Box box = tri_box(sc.triangles[sc.order[first]]);
    for (int i = first + 1; i < last; ++i) box = merge(box, tri_box(sc.triangles[sc.order[i]]));
    if (size <= 4) { out_child = ~first; out_count = size; return box; }
    float3 clo = centroids[sc.order[first]], chi = clo;
    for (int i = first + 1; i < last; ++i) {
        float3 c = centroids[sc.order[i]];
        clo = make_float3(fminf(clo.x, c.x), fminf(clo.y, c.y), fminf(clo.z, c.z));
        chi = make_float3(fmaxf(chi.x, c.x), fmaxf(chi.y, c.y), fmaxf(chi.z, c.z));
    }
    float3 e = chi - clo;
    int axis = e.x > e.y ? (e.x > e.z ? 0 : 2) : (e.y > e.z ? 1 : 2);
    int mid = first + size / 2;
    std::nth_element(sc.order.begin() + first, sc.order.begin() + mid, sc.order.begin() + last,
                     [&](int a, int b) { float ca = comp(centroids[a], axis), cb = comp(centroids[b], axis); return ca < cb || (ca == cb && a < b); });
    int index = (int)sc.nodes.size();
    sc.nodes.push_back(Node());
    Node n;
    n.box[0] = build_range(sc, centroids, first, mid, n.child[0], n.count[0]);
    n.box[1] = build_range(sc, centroids, mid, last, n.child[1], n.count[1]);
    sc.nodes[index] = n;
    out_child = index; out_count = 0;
    return box;
}

void build_bvh(Scene& sc) {
    int n = (int)sc.triangles.size();
    sc.order.resize(n);
    for (int i = 0; i < n; ++i) sc.order[i] = i;
    std::vector<float3> centroids(n);
    for (int i = 0; i < n; ++i) { Box b = tri_box(sc.triangles[i]); centroids[i] = (b.lo + b.hi) * 0.5f; }
    sc.nodes.clear();
    Node root = {};
    root.child[0] = root.child[1] = -1; root.count[0] = root.count[1] = -1;
    if (n == 0) { sc.nodes.push_back(root); return; }
    if (n <= 4) {
        root.box[0] = tri_box(sc.triangles[0]);
        for (int i = 1; i < n; ++i) root.box[0] = merge(root.box[0], tri_box(sc.triangles[i]));
        root.child[0] = ~0; root.count[0] = n;
        sc.nodes.push_back(root);
        return;
    }
    int child, count;
    build_range(sc, centroids, 0, n, child, count); // the root of a range > 4 is always node 0
}

// ---- integrator ----------------------------------------------------------------------------------

struct Counters { uint64_t extend_rays = 0, shadow_rays = 0; };



inline float3 transform_normal(const float* nm, float3 n) {
    return make_float3(nm[0] * n.x + nm[1] * n.y + nm[2] * n.z, nm[3] * n.x + nm[4] * n.y + nm[5] * n.z, nm[6] * n.x + nm[7] * n.y + nm[8] * n.z);
}

// Environment lookups that the reference performs with texture hardware (PresampledEnvironmentLightImpl.h:17-55):
// bilinear RGBA fetch with wrap in u / clamp in v, nearest PDF fetch with clamp.
inline float3 env_fetch_bilinear(const EnvironmentIn& e, float2 uv) {
    float x = uv.x * e.width - 0.5f, y = uv.y * e.height - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float tx = x - fx, ty = y - fy;
    int x0 = int(fx), y0 = int(fy), x1 = x0 + 1, y1 = y0 + 1;
    x0 = ((x0 % e.width) + e.width) % e.width; x1 = ((x1 % e.width) + e.width) % e.width;
    y0 = std::max(0, std::min(e.height - 1, y0)); y1 = std::max(0, std::min(e.height - 1, y1));
    float3 p00 = make_float3(e.texels[y0 * e.width + x0]), p10 = make_float3(e.texels[y0 * e.width + x1]);
    float3 p01 = make_float3(e.texels[y1 * e.width + x0]), p11 = make_float3(e.texels[y1 * e.width + x1]);
    float3 a = lerp(p00, p10, tx), b = lerp(p01, p11, tx);
    return lerp(a, b, ty);
}
inline float env_fetch_pdf(const EnvironmentIn& e, float2 uv) {
    int x = int(floorf(uv.x * e.pdf_width)), y = int(floorf(uv.y * e.pdf_height));
    x = std::max(0, std::min(e.pdf_width - 1, x)); // RT_WRAP_CLAMP_TO_EDGE, PresampledEnvironmentMap.cpp:46-47
    y = std::max(0, std::min(e.pdf_height - 1, y));
    return e.pdf[y * e.pdf_width + x];
}
inline LightSample env_sample_radiance(const EnvironmentIn& e, float2 u) {
    int index = u.x * (int)e.samples.size();
    LightSample s = e.samples[index];
    s.radiance *= e.tint;
    return s;
}
// Deviation from Utils.h:288-292 (shared with the product, bpt_lights.cuh): direction.y is clamped to [-1, 1] before asinf.
// A normalised direction can be y = 1 + 1 ulp; the reference then passes a NaN coordinate to the texture unit, which
// tolerates it, while this software fetch would return NaN radiance.
inline float2 latlong_texcoord_clamped(float3 direction) {
    direction.y = fminf(fmaxf(direction.y, -1.0f), 1.0f);
    return direction_to_latlong_texcoord(direction);
}
// sample_CDFs_for_uv + sample_radiance(EnvironmentLight), EnvironmentLightImpl.h:22-83. The reference reads the CDFs through
// unfiltered textures at integer coordinates (plain array reads here), the map bilinearly and the per pixel PDF nearest.
inline LightSample env_sample_radiance_cdf(const EnvironmentIn& e, float2 random_sample) {
    if (e.marginal_cdf.empty()) return LightSample::none();
    float2 uv;
    int conditional_row = 0;
    {
        int lowerbound = 0, upperbound = e.pdf_height;
        while (lowerbound + 1 != upperbound) {
            int middlebound = (lowerbound + upperbound) / 2;
            float cdf = e.marginal_cdf[middlebound];
            if (random_sample.y < cdf) upperbound = middlebound; else lowerbound = middlebound;
        }
        conditional_row = lowerbound;
        float cdf_at_lowerbound = e.marginal_cdf[lowerbound];
        float dv = random_sample.y - cdf_at_lowerbound;
        dv /= e.marginal_cdf[lowerbound + 1] - cdf_at_lowerbound;
        uv.y = (lowerbound + dv) / float(e.pdf_height);
    }
    {
        const float* row = e.conditional_cdf.data() + (size_t)conditional_row * (e.pdf_width + 1);
        int lowerbound = 0, upperbound = e.pdf_width;
        while (lowerbound + 1 != upperbound) {
            int middlebound = (lowerbound + upperbound) / 2;
            float cdf = row[middlebound];
            if (random_sample.x < cdf) upperbound = middlebound; else lowerbound = middlebound;
        }
        float cdf_at_lowerbound = row[lowerbound];
        float du = random_sample.x - cdf_at_lowerbound;
        du /= row[lowerbound + 1] - cdf_at_lowerbound;
        uv.x = (lowerbound + du) / float(e.pdf_width);
    }
    LightSample sample;
    sample.direction_to_light = latlong_texcoord_to_direction(uv);
    sample.distance = 1e30f;
    sample.radiance = env_fetch_bilinear(e, uv);
    sample.radiance *= e.tint;
    float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - sample.direction_to_light.y * sample.direction_to_light.y));
    float PDF = env_fetch_pdf(e, uv) / sin_theta;
    sample.PDF = sin_theta == 0.0f ? 0.0f : PDF;
    return sample;
}
inline PDF env_pdf(const EnvironmentIn& e, float3 direction_to_light) {
    float2 uv = latlong_texcoord_clamped(direction_to_light);
    float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - direction_to_light.y * direction_to_light.y));
    float p = env_fetch_pdf(e, uv) / sin_theta;
    return sin_theta == 0.0f ? PDF::delta_dirac(0) : PDF(p);
}
inline float3 env_evaluate(const EnvironmentIn& e, float3 direction_to_light) {
    float2 uv = latlong_texcoord_clamped(direction_to_light);
    return e.tint * env_fetch_bilinear(e, uv);
}

// LightImpl.h:38-108 dispatch, restated for the host (the original is __inline_dev__ and reads OptiX buffers).
LightSample light_sample_radiance(const Scene& sc, const Light& light, float3 position, float2 u) {
    switch (light.get_type()) {
    case Light::Sphere: return LightSources::sample_radiance(light.sphere, position, u);
    case Light::Directional: return LightSources::sample_radiance(light.directional, u);
    case Light::Environment: return env_sample_radiance_cdf(sc.env, u);
    case Light::PresampledEnvironment: return env_sample_radiance(sc.env, u);
    case Light::Spot: return LightSources::sample_radiance(light.spot, position, u);
    default: return LightSample::none();
    }
}

using Shading::ShadingModels::DefaultShading;

// DefaultMaterialCreator::create, MonteCarlo.cu:239-244 -> DefaultShading::initialize_with_max_PDF_hint
// (DefaultShading.h:155-179, GPU_DEVICE only) restated through the class' own setup functions.
DefaultShading create_default_shading(const Material& m, float4 tint_and_roughness_scale, float abs_cos_theta_o, PDF max_PDF_hint) {
    float min_roughness = Shading::ShadingModels::GGXMinimumRoughness::from_PDF(abs_cos_theta_o, max_PDF_hint);
    float coat_roughness = fmaxf(float(m.coat_roughness), min_roughness);
    float metallic = m.metallic;
    float4 tint_roughness = make_float4(m.tint, m.roughness) * tint_and_roughness_scale;
    float3 tint = make_float3(tint_roughness);
    float roughness = fmaxf(tint_roughness.w, min_roughness);
    DefaultShading shading(m, abs_cos_theta_o);
    float coat_rho;
    shading.setup_shading(tint, roughness, m.specularity, metallic, m.coat, coat_roughness, abs_cos_theta_o, coat_rho);
    shading.setup_sampling_probabilities(abs_cos_theta_o, coat_rho);
    return shading;
}

struct PathState {
    const Scene* scene;
    const CameraIn* camera;
    const SettingsIn* settings;
    MonteCarloPayload payload;
    Counters* counters;
};

inline float4 rng_sample4f(const MonteCarloPayload& p, unsigned int sampling_dimension) {
    unsigned int dimension = RngSamplingDimension::MAX_DIMENSIONS * p.bounces + sampling_dimension; // Types.h:452-459
    return RNG::PracticalScrambledSobol::sample4f(p.accumulation_count, p.pixel_hash, dimension);
}

// sample_single_light, MonteCarlo.cu:61-87
template <class ShadingModel>
LightSample sample_single_light(const Scene& sc, const ShadingModel& material, float3 intersection_point, float3 wo, const TBN& world_shading_tbn, float3 random_sample) {
    int light_index = std::min(sc.light_count - 1, int(random_sample.z * sc.light_count));
    const Light& light = sc.lights[light_index];
    LightSample light_sample = light_sample_radiance(sc, light, intersection_point, make_float2(random_sample));
    light_sample.radiance *= sc.light_count;

    float N_dot_L = dot(world_shading_tbn.get_normal(), light_sample.direction_to_light);
    light_sample.radiance *= abs(N_dot_L) / light_sample.PDF.value();

    const float3 shading_light_direction = world_shading_tbn * light_sample.direction_to_light;
    BSDFResponse bsdf_response = material.evaluate_with_PDF(wo, shading_light_direction);
    bool apply_MIS = !light_sample.PDF.is_delta_dirac();
    if (apply_MIS)
        light_sample.radiance *= MonteCarlo::MIS_weight(light_sample.PDF, bsdf_response.PDF);
    else
        bsdf_response.reflectance = fminf(bsdf_response.reflectance, make_float3(32.0f));
    light_sample.radiance *= bsdf_response.reflectance;
    return light_sample;
}

// reestimated_light_samples, MonteCarlo.cu:91-123
template <class ShadingModel>
LightSample reestimated_light_samples(const Scene& sc, const SettingsIn& settings, const MonteCarloPayload& payload, const ShadingModel& material,
                                      float3 intersection_point, float3 wo, const TBN& world_shading_tbn) {
    if (sc.light_count == 0)
        return LightSample::none();
    float4 light_random_base = rng_sample4f(payload, RngSamplingDimension::NEXT_EVENT_ESTIMATION);
    LightSample light_sample = LightSample::none();
    int light_sample_count = settings.next_event_sample_count;
    for (int s = 0; s < light_sample_count; ++s) {
        float4 light_random_4f = toroidal_shift(light_random_base, sc.nee_offsets[s]);
        float3 light_random_number = make_float3(light_random_4f);
        float use_new_light_decision = light_random_4f.w;
        LightSample new_light_sample = sample_single_light(sc, material, intersection_point, wo, world_shading_tbn, light_random_number);
        float light_weight = sum(light_sample.radiance);
        float new_light_weight = sum(new_light_sample.radiance);
        float new_light_probability = new_light_weight / (light_weight + new_light_weight);
        if (use_new_light_decision < new_light_probability) {
            light_sample = new_light_sample;
            light_sample.radiance /= new_light_probability;
        } else
            light_sample.radiance /= 1.0f - new_light_probability;
    }
    light_sample.radiance /= light_sample_count;
    return light_sample;
}

// interpolate_attributes (TriangleAttributes.cu:35-84) + path_tracing_closest_hit (MonteCarlo.cu:129-233)
void triangle_closest_hit(PathState& st, const HitRecord& hit, float3 ray_origin, float3 ray_direction) {
    const Scene& sc = *st.scene;
    MonteCarloPayload& payload = st.payload;
    const Triangle& tri = sc.triangles[hit.primitive];
    const float t_hit = hit.t;

    // -- attribute program --
    float3 geometric_normal = normalize(cross(tri.p1 - tri.p0, tri.p2 - tri.p0));
    const float2 barycentrics = make_float2(hit.u, hit.v);
    float barycentrics_z = 1.0f - barycentrics.x - barycentrics.y;
    float3 intersection_point = tri.p1 * barycentrics.x + tri.p2 * barycentrics.y + tri.p0 * barycentrics_z;
    const float* normal_matrix = sc.normal_matrices.data() + 9 * tri.instance;
    float3 shading_normal;
    if (tri.has_normals) {
        shading_normal = tri.n1.decode() * barycentrics.x + tri.n2.decode() * barycentrics.y + tri.n0.decode() * barycentrics_z;
        shading_normal = normalize(shading_normal);
    } else
        shading_normal = geometric_normal;
    float2 texcoord = interpolate_texcoord(tri, barycentrics.x, barycentrics.y);
    float4 tint_and_roughness_scale;
    if (tri.has_tints) {
        const float byte_to_float_normalizer = 1.0f / 255.0f;
        uchar4 tint0 = tri.t0, tint1 = tri.t1, tint2 = tri.t2;
        tint_and_roughness_scale.x = (tint1.x * barycentrics.x + tint2.x * barycentrics.y + tint0.x * barycentrics_z) * byte_to_float_normalizer;
        tint_and_roughness_scale.y = (tint1.y * barycentrics.x + tint2.y * barycentrics.y + tint0.y * barycentrics_z) * byte_to_float_normalizer;
        tint_and_roughness_scale.z = (tint1.z * barycentrics.x + tint2.z * barycentrics.y + tint0.z * barycentrics_z) * byte_to_float_normalizer;
        tint_and_roughness_scale.w = (tint1.w * barycentrics.x + tint2.w * barycentrics.y + tint0.w * barycentrics_z) * byte_to_float_normalizer;
    } else
        tint_and_roughness_scale = make_float4(1.0f);
    float3 emission = make_float3(1.0f); // TriangleAttributes.cu:78-83
    if (tri.has_emission)
        emission = tri.e1 * barycentrics.x + tri.e2 * barycentrics.y + tri.e0 * barycentrics_z;

    // -- closest hit --
    payload.light_sample = LightSample::none();
    InstanceID instance_id = InstanceID::make(InstanceID::Type::MeshModel, 0);
    PrimitiveID primitive_id = PrimitiveID::make(instance_id, hit.primitive); // global primitive index: unique across instances
    if (primitive_id == payload.primitive_id) {
        payload.ray_min_t = nextafterf(t_hit, INFINITY);
        return;
    }
    const Material material_parameter = material_at(sc, sc.materials[tri.material], texcoord);

    // The geometry is already in world space: the geometric normal needs no transform, the shading normal
    // (object space, per vertex) goes through the instance's inverse transpose like rtTransformNormal.
    float3 world_geometric_normal = geometric_normal;
    bool hit_from_front = dot(world_geometric_normal, ray_direction) < 0.0f;
    bool backside_cull = !hit_from_front && !material_parameter.is_thin_walled();
    backside_cull &= !material_parameter.is_transmissive();

    float4 bsdf_coverage_random_4f = rng_sample4f(payload, RngSamplingDimension::BSDF);
    float coverage_cutoff = bsdf_coverage_random_4f.w;
    float3 bsdf_random_uvs = make_float3(bsdf_coverage_random_4f);
    float coverage = material_coverage(sc, material_parameter, texcoord);
    bool discard_from_coverage = coverage < coverage_cutoff;
    if (backside_cull || discard_from_coverage) {
        payload.ray_min_t = nextafterf(t_hit, INFINITY);
        return;
    }

    payload.material_index = tri.material;
    payload.texcoord = texcoord;
    payload.tint_and_roughness_scale = float_to_unorm8(tint_and_roughness_scale);
    payload.primitive_id = primitive_id;

    world_geometric_normal = hit_from_front ? world_geometric_normal : -world_geometric_normal;
    float3 world_shading_normal = tri.has_normals ? normalize(transform_normal(normal_matrix, shading_normal)) : shading_normal;
    world_shading_normal = hit_from_front ? world_shading_normal : -world_shading_normal;
    world_shading_normal = fix_backfacing_shading_normal(-ray_direction, world_shading_normal, 0.002f);
    payload.shading_normal = world_shading_normal;
    const TBN world_shading_tbn = TBN(world_shading_normal);

    float3 world_intersection_point = intersection_point;
    float3 world_wo = -ray_direction;
    float3 wo = world_shading_tbn * world_wo;

    float cos_theta = hit_from_front || material_parameter.is_thin_walled() ? wo.z : -wo.z;
    payload.radiance += payload.throughput * emission * material_parameter.emission;

    // The shading-model specific part of path_tracing_closest_hit<MaterialCreator>, MonteCarlo.cu:191-212.
    BSDFSample bsdf_sample;
    auto light_and_bsdf = [&](const auto& material) {
        payload.light_sample = reestimated_light_samples(sc, *st.settings, payload, material, world_intersection_point, wo, world_shading_tbn);
        payload.light_sample_origin = offset_ray_origin(world_intersection_point, payload.light_sample.direction_to_light, world_geometric_normal);
        payload.light_sample.radiance *= payload.throughput;
        bsdf_sample = material.sample(wo, bsdf_random_uvs);
    };
    if (material_parameter.shading_model == Material::ShadingModel::Diffuse) {
        // DiffuseMaterialCreator::create, MonteCarlo.cu:250-255
        float4 tint_roughness = make_float4(material_parameter.tint, material_parameter.roughness) * tint_and_roughness_scale;
        light_and_bsdf(Shading::ShadingModels::DiffuseShading(make_float3(tint_roughness), tint_roughness.w));
    } else if (material_parameter.shading_model == Material::ShadingModel::Transmissive) {
        // TransmissiveMaterialCreator::create, MonteCarlo.cu:261-266 -> TransmissiveShading::initialize_with_max_PDF_hint
        // (TransmissiveShading.h:51-66, GPU_DEVICE only) restated through the class' public setup_shading.
        PDF max_PDF_hint = payload.bsdf_PDF * st.settings->path_regularization_pdf_scale;
        float min_roughness = Shading::ShadingModels::GGXMinimumRoughness::from_PDF(abs(cos_theta), max_PDF_hint);
        float4 tint_roughness = make_float4(material_parameter.tint, material_parameter.roughness) * tint_and_roughness_scale;
        Shading::ShadingModels::TransmissiveShading transmissive(material_parameter, cos_theta);
        transmissive.setup_shading(make_float3(tint_roughness), fmaxf(tint_roughness.w, min_roughness), material_parameter.specularity, cos_theta);
        light_and_bsdf(transmissive);
    } else {
        PDF max_PDF_hint = payload.bsdf_PDF * st.settings->path_regularization_pdf_scale;
        light_and_bsdf(create_default_shading(material_parameter, tint_and_roughness_scale, cos_theta, max_PDF_hint));
    }

    bool is_reflection = bsdf_sample.direction.z >= 0;
    payload.direction = bsdf_sample.direction * world_shading_tbn;
    payload.bsdf_PDF = bsdf_sample.PDF;
    if (bsdf_sample.PDF.is_valid())
        payload.throughput *= bsdf_sample.reflectance * abs(bsdf_sample.direction.z) / bsdf_sample.PDF.value();
    else
        payload.throughput = make_float3(0.0f);

    float cos_geometric_theta_i = dot(payload.direction, world_geometric_normal);
    if (is_reflection ? cos_geometric_theta_i < 0.0f : cos_geometric_theta_i >= 0.0f)
        payload.direction = reflect(payload.direction, world_geometric_normal);

    payload.position = offset_ray_origin(world_intersection_point, payload.direction, world_geometric_normal);
    payload.ray_min_t = 0.0f;
    // Russian roulette: NOT in the reference. An opt-in extension of the product (bpt_settings.russian_roulette_start_bounce)
    // restated here so that it can be checked sample for sample; it draws from RNG dimension 3, which the reference leaves free.
    if (st.settings->russian_roulette_start_bounce != 0u && payload.bounces + 1u >= st.settings->russian_roulette_start_bounce &&
        (payload.throughput.x > 0.0f || payload.throughput.y > 0.0f || payload.throughput.z > 0.0f)) {
        float survival = fminf(fmaxf(fmaxf(fmaxf(payload.throughput.x, payload.throughput.y), payload.throughput.z), 0.05f), 1.0f);
        float u = rng_sample4f(payload, RngSamplingDimension(3)).x;
        if (u < survival) payload.throughput = payload.throughput / survival;
        else payload.throughput = make_float3(0.0f);
    }
    payload.bounces += 1u;
    if (!payload.light_sample.PDF.is_valid())
        payload.bsdf_PDF.disable_MIS();
}

// light_closest_hit (MonteCarlo.cu:291-302) with evaluate_intersection (LightImpl.h:86-108), which reads the
// current ray's origin and direction.
void light_closest_hit(PathState& st, int light_index, float t_hit, float3 ray_origin, float3 ray_direction, float3 light_shading_normal) {
    MonteCarloPayload& payload = st.payload;
    const Light& light = st.scene->lights[light_index];
    float3 light_radiance;
    PDF bsdf_PDF = payload.bsdf_PDF;
    if (light.get_type() == Light::Sphere) {
        light_radiance = LightSources::evaluate(light.sphere, ray_origin, ray_direction);
        if (bsdf_PDF.use_for_MIS())
            light_radiance *= MonteCarlo::MIS_weight(bsdf_PDF, LightSources::pdf(light.sphere, ray_origin, ray_direction));
    } else if (light.get_type() == Light::Spot) {
        light_radiance = LightSources::evaluate(light.spot, ray_origin, ray_direction);
        if (bsdf_PDF.use_for_MIS())
            light_radiance *= MonteCarlo::MIS_weight(bsdf_PDF, LightSources::pdf(light.spot, ray_origin, ray_direction));
    } else
        light_radiance = make_float3(1000.0f, 0, 1000);

    payload.throughput = fminf(payload.throughput, make_float3(4));
    payload.radiance += payload.throughput * light_radiance;
    payload.throughput = make_float3(0.0f);
    payload.position = ray_direction * t_hit + ray_origin;
    payload.shading_normal = light_shading_normal;
    payload.primitive_id = PrimitiveID::make(InstanceID::analytical_light_sources(), light_index);
}

// miss, SimpleRGPs.cu:349-362
void miss(PathState& st, float3 ray_direction) {
    MonteCarloPayload& payload = st.payload;
    const EnvironmentIn& env = st.scene->env;
    float3 environment_radiance = env.tint;
    if (!env.texels.empty()) {
        environment_radiance = env_evaluate(env, ray_direction);
        if (payload.bsdf_PDF.use_for_MIS())
            environment_radiance *= MonteCarlo::MIS_weight(payload.bsdf_PDF, env_pdf(env, ray_direction));
    }
    payload.radiance += payload.throughput * environment_radiance;
    payload.throughput = make_float3(0.0f);
    payload.position = 1e30f * payload.direction;
    payload.shading_normal = -ray_direction;
    payload.primitive_id = PrimitiveID::make(InstanceID::analytical_light_sources(), 0xFFFFFFFF);
}

// One rtTrace of a MonteCarlo ray: triangles via the BVH, analytical lights via LightSources.cu:31-70.
void trace_monte_carlo(PathState& st) {
    const Scene& sc = *st.scene;
    MonteCarloPayload& payload = st.payload;
    float3 origin = payload.position, direction = payload.direction;
    float tmin = payload.ray_min_t;
    const float tmax = 1e27f; // RT_DEFAULT_MAX
    st.counters->extend_rays++;
    HitRecord hit = closest_bvh(sc, origin, direction, tmin, tmax);
    float t_closest = hit.primitive >= 0 ? hit.t : tmax;
    int light_hit = -1; float light_t = 0;
    for (int i = 0; i < sc.light_count; ++i) {
        const Light& light = sc.lights[i];
        float t = -1e30f;
        float radius = 0.0f;
        if (light.get_type() == Light::Sphere) {
            radius = light.sphere.radius;
            t = Intersect::ray_sphere(origin, direction, light.sphere.position, light.sphere.radius);
        } else if (light.get_type() == Light::Spot) {
            radius = light.spot.radius;
            t = Intersect::ray_disk(origin, direction, light.spot.position, light.spot.direction, light.spot.radius);
        }
        if (!(radius > 0.0f)) continue; // bounds program invalidates the box: never intersected
        if (t > tmin && t < t_closest) { t_closest = t; light_hit = i; light_t = t; } // rtPotentialIntersection
    }
    if (light_hit >= 0) {
        const Light& light = sc.lights[light_hit];
        float3 coarse = light_t * direction + origin;
        float3 n = light.get_type() == Light::Sphere ? normalize(coarse - light.sphere.position) : light.spot.direction;
        light_closest_hit(st, light_hit, light_t, origin, direction, n);
    } else if (hit.primitive >= 0)
        triangle_closest_hit(st, hit, origin, direction);
    else
        miss(st, direction);
}

// initialize_monte_carlo_payload + fill_ray_info, SimpleRGPs.cu:44-72
MonteCarloPayload initialize_payload(int x, int y, int image_width, int image_height, int accumulation_count, const CameraIn& cam) {
    MonteCarloPayload payload = {};
    payload.pixel_hash = RNG::pcg2d(x, y).x;
    payload.accumulation_count = accumulation_count;
    payload.throughput = make_float3(1.0f);
    payload.light_sample = LightSample::none();
    payload.bsdf_PDF = PDF::delta_dirac();

    float2 screen_pos = make_float2(x, y) + (accumulation_count == 0 ? make_float2(0.5f) : make_float2(rng_sample4f(payload, RngSamplingDimension::CAMERA_PARAMETERS)));
    float2 viewport_pos = make_float2(screen_pos.x / float(image_width), screen_pos.y / float(image_height));

    Matrix4x4 inverse_view_projection(cam.inverse_view_projection), inverse_projection(cam.inverse_projection);
    Matrix3x3 view_to_world_rotation(cam.view_to_world_rotation);
    float4 NDC_near_pos = make_float4(viewport_pos.x * 2.0f - 1.0f, viewport_pos.y * 2.0f - 1.0f, -1.0f, 1.0f);
    float4 scaled_near_world_pos = inverse_view_projection * NDC_near_pos;
    payload.position = make_float3(scaled_near_world_pos) / scaled_near_world_pos.w;
    float4 NDC_far_pos = make_float4(NDC_near_pos.x, NDC_near_pos.y, 1.0f, 1.0f);
    float4 scaled_near_view_pos = inverse_projection * NDC_far_pos;
    payload.direction = normalize(view_to_world_rotation * make_float3(scaled_near_view_pos));
    return payload;
}

// path_tracing_RPG + path_trace_single_bounce, SimpleRGPs.cu:112-140
float3 trace_path(const Scene& sc, const CameraIn& cam, const SettingsIn& settings, int x, int y, int w, int h, int accumulation_count, Counters& counters) {
    PathState st = { &sc, &cam, &settings, initialize_payload(x, y, w, h, accumulation_count, cam), &counters };
    MonteCarloPayload& payload = st.payload;
    payload.primitive_id = PrimitiveID::make(InstanceID::make(InstanceID::Type::MeshModel, 0), -1);
    do {
        payload.material_index = 0;
        trace_monte_carlo(st);
        const LightSample& light_sample = payload.light_sample;
        if (light_sample.radiance.x > 0 || light_sample.radiance.y > 0 || light_sample.radiance.z > 0) {
            counters.shadow_rays++;
            float transmission = transmission_bvh(sc, payload.light_sample_origin, light_sample.direction_to_light, 0.0f, light_sample.distance, false,
                                                  fmaxf(fmaxf(light_sample.radiance.x, light_sample.radiance.y), light_sample.radiance.z));
            payload.radiance += light_sample.radiance * transmission;
        }
        payload.light_sample = LightSample::none();
    } while (payload.bounces <= settings.max_bounce_count && !is_black(payload.throughput));
    return payload.radiance;
}

} // namespace

extern "C" {

void* pto_scene_create() {
    Scene* sc = new Scene();
    for (int i = 0; i < 256; ++i) sc->nee_offsets[i] = RNG::ReverseHalton(i).sample4f(); // Renderer.cpp:323-336
    Material m = {}; m.coverage = 1.0f;
    sc->materials.assign(1, m);
    return sc;
}
void pto_scene_destroy(void* s) { delete (Scene*)s; }

void pto_scene_add_mesh(void* s, int mesh_id, const uint32_t* indices, int primitive_count, const float* positions, const float* normals,
                        const uint8_t* tint_roughness, int vertex_count) {
    Mesh& m = ((Scene*)s)->meshes[mesh_id];
    m.indices.assign(indices, indices + 3ll * primitive_count);
    m.positions.resize(vertex_count);
    for (int i = 0; i < vertex_count; ++i) m.positions[i] = make_float3(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
    m.normals.clear(); m.tints.clear();
    if (normals) {
        m.normals.resize(vertex_count);
        for (int i = 0; i < vertex_count; ++i) { // load_mesh, Renderer.cpp:104-108
            auto e = Bifrost::Math::OctahedralNormal::encode_precise(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]);
            m.normals[i].encoding = make_short2(e.encoding.x, e.encoding.y);
        }
    }
    if (tint_roughness) {
        m.tints.resize(vertex_count);
        for (int i = 0; i < vertex_count; ++i) m.tints[i] = make_uchar4(tint_roughness[4 * i], tint_roughness[4 * i + 1], tint_roughness[4 * i + 2], tint_roughness[4 * i + 3]);
    }
}
void pto_scene_set_mesh_emission(void* s, int mesh_id, const float* emission, int vertex_count) {
    Mesh& m = ((Scene*)s)->meshes[mesh_id];
    m.emission.resize(vertex_count);
    for (int i = 0; i < vertex_count; ++i) m.emission[i] = make_float3(emission[3 * i], emission[3 * i + 1], emission[3 * i + 2]);
}
void pto_scene_set_mesh_texcoords(void* s, int mesh_id, const float* texcoords, int vertex_count) {
    Mesh& m = ((Scene*)s)->meshes[mesh_id];
    m.texcoords.resize(vertex_count);
    for (int i = 0; i < vertex_count; ++i) m.texcoords[i] = make_float2(texcoords[2 * i], texcoords[2 * i + 1]);
}
// pixel_format: Bifrost::Assets::PixelFormat values (Alpha8 1, RGB24 3, RGBA32 4, RGB_Float 6, RGBA_Float 7).
void pto_scene_add_texture(void* s, int texture_id, int width, int height, int pixel_format, int is_srgb, int wrap_u, int wrap_v, int linear_filter,
                           const void* pixels) {
    Texture& tex = ((Scene*)s)->textures[texture_id];
    tex.width = width; tex.height = height; tex.linear = linear_filter != 0; tex.wrap_u = wrap_u; tex.wrap_v = wrap_v;
    tex.texels.resize((size_t)width * height);
    const unsigned char* bytes = (const unsigned char*)pixels;
    const float* floats = (const float*)pixels;
    auto decode = [&](unsigned char b, bool color) { float c = b / 255.0f; return (is_srgb && color) ? srgb_to_linear(c) : c; };
    for (size_t i = 0; i < tex.texels.size(); ++i) {
        switch (pixel_format) {
        case 1: tex.channels = 1; tex.texels[i] = make_float4(decode(bytes[i], true), 0.0f, 0.0f, 1.0f); break;
        case 3: tex.channels = 4; tex.texels[i] = make_float4(decode(bytes[3 * i], true), decode(bytes[3 * i + 1], true), decode(bytes[3 * i + 2], true), 1.0f); break;
        case 4: tex.channels = 4; tex.texels[i] = make_float4(decode(bytes[4 * i], true), decode(bytes[4 * i + 1], true), decode(bytes[4 * i + 2], true), decode(bytes[4 * i + 3], false)); break;
        case 6: tex.channels = 4; tex.texels[i] = make_float4(floats[3 * i], floats[3 * i + 1], floats[3 * i + 2], 1.0f); break;
        default: tex.channels = 4; tex.texels[i] = make_float4(floats[4 * i], floats[4 * i + 1], floats[4 * i + 2], floats[4 * i + 3]); break;
        }
    }
}
void pto_texture_sample(void* s, int texture_id, int64_t n, const float* uv, float* out_rgba) {
    const Texture& tex = ((Scene*)s)->textures.at(texture_id);
    for (int64_t i = 0; i < n; ++i) {
        float4 r = texture_sample(tex, uv[2 * i], uv[2 * i + 1]);
        out_rgba[4 * i] = r.x; out_rgba[4 * i + 1] = r.y; out_rgba[4 * i + 2] = r.z; out_rgba[4 * i + 3] = r.w;
    }
}
void pto_scene_set_instances(void* s, const void* instances, int count) { ((Scene*)s)->instances.assign((const InstanceIn*)instances, (const InstanceIn*)instances + count); }
void pto_scene_set_materials(void* s, const void* materials, int count) { ((Scene*)s)->materials.assign((const Material*)materials, (const Material*)materials + count); }
void pto_scene_set_lights(void* s, const void* lights, int count) {
    Scene* sc = (Scene*)s;
    sc->lights.assign((const Light*)lights, (const Light*)lights + count);
    sc->light_count = count;
    sc->env_in_light_list = false;
}
void pto_scene_set_environment(void* s, const float* tint, const float* texels, int width, int height, const float* pdf, int pdf_width, int pdf_height,
                               const void* samples, int sample_count) {
    Scene* sc = (Scene*)s;
    EnvironmentIn& e = sc->env;
    e = EnvironmentIn();
    e.tint = make_float3(tint[0], tint[1], tint[2]);
    if (sc->env_in_light_list) { sc->lights.pop_back(); sc->light_count--; sc->env_in_light_list = false; }
    if (!texels) return;
    e.width = width; e.height = height; e.pdf_width = pdf_width; e.pdf_height = pdf_height;
    e.texels.assign((const float4*)texels, (const float4*)texels + (size_t)width * height);
    e.pdf.assign(pdf, pdf + (size_t)pdf_width * pdf_height);
    e.samples.assign((const LightSample*)samples, (const LightSample*)samples + sample_count);
    if (sample_count > 1) { // next_event_estimation_possible, PresampledEnvironmentMap.h:64; appended last, Renderer.cpp:1180-1195
        Light l = {};
        l.flags = Light::PresampledEnvironment;
        sc->lights.push_back(l);
        sc->light_count++;
        sc->env_in_light_list = true;
    }
}
// Light::Environment instead of Light::PresampledEnvironment for next event estimation (bpt_set_environment_cdfs +
// bpt_set_environment_sampling(BPT_ENVIRONMENT_NEE_CDF) on the product side). Call after pto_scene_set_environment.
void pto_scene_set_environment_cdfs(void* s, const float* marginal, const float* conditional, int pdf_width, int pdf_height, int sample_by_cdf) {
    Scene* sc = (Scene*)s;
    EnvironmentIn& e = sc->env;
    e.marginal_cdf.clear(); e.conditional_cdf.clear();
    if (marginal && conditional && pdf_width == e.pdf_width && pdf_height == e.pdf_height) {
        e.marginal_cdf.assign(marginal, marginal + pdf_height + 1);
        e.conditional_cdf.assign(conditional, conditional + (size_t)(pdf_width + 1) * pdf_height);
    }
    if (sc->env_in_light_list) sc->lights.back().flags = sample_by_cdf ? Light::Environment : Light::PresampledEnvironment;
}
void pto_scene_build(void* s) { Scene* sc = (Scene*)s; flatten(*sc); build_bvh(*sc); }
int64_t pto_triangle_count(void* s) { return (int64_t)((Scene*)s)->triangles.size(); }
void pto_world_vertices(void* s, float* out9) {
    Scene* sc = (Scene*)s;
    for (size_t i = 0; i < sc->triangles.size(); ++i) {
        const Triangle& t = sc->triangles[i];
        float v[9] = { t.p0.x, t.p0.y, t.p0.z, t.p1.x, t.p1.y, t.p1.z, t.p2.x, t.p2.y, t.p2.z };
        memcpy(out9 + 9 * i, v, sizeof(v));
    }
}

// brute != 0: loop over all triangles; otherwise the median split BVH. Both must agree bit for bit.
void pto_intersect(void* s, int64_t n, const float* origins, const float* directions, const float* tmin, const float* tmax, int brute,
                   int32_t* out_primitive, float* out_t, float* out_uv, uint8_t* out_occluded) {
    const Scene& sc = *(Scene*)s;
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        float3 o = make_float3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
        float3 d = make_float3(directions[3 * i], directions[3 * i + 1], directions[3 * i + 2]);
        HitRecord h = brute ? closest_brute(sc, o, d, tmin[i], tmax[i]) : closest_bvh(sc, o, d, tmin[i], tmax[i]);
        if (out_primitive) out_primitive[i] = h.primitive;
        if (out_t) out_t[i] = h.primitive >= 0 ? h.t : INFINITY;
        if (out_uv) { out_uv[2 * i] = h.u; out_uv[2 * i + 1] = h.v; }
        if (out_occluded) out_occluded[i] = transmission_bvh(sc, o, d, tmin[i], tmax[i], brute != 0) < 1.0f ? 1 : 0;
    }
}

// Adds `sample_count` samples (accumulation indices first_sample ...) per pixel to accum_sum (double4 per pixel:
// radiance sum and sample count), for the pixel rows [row_begin, row_end). threads <= 0: all OpenMP threads.
void pto_render(void* s, const void* camera, const void* settings_in, int width, int height, uint32_t first_sample, uint32_t sample_count,
                int row_begin, int row_end, double* accum_sum, uint64_t* out_counters /*[2]*/, int threads) {
    const Scene& sc = *(Scene*)s;
    const CameraIn& cam = *(const CameraIn*)camera;
    const SettingsIn& settings = *(const SettingsIn*)settings_in;
    if (threads <= 0) threads = omp_get_max_threads();
    uint64_t extend = 0, shadow = 0;
    #pragma omp parallel for schedule(dynamic, 16) num_threads(threads) reduction(+ : extend, shadow) collapse(2)
    for (int y = row_begin; y < row_end; ++y)
        for (int x = 0; x < width; ++x) {
            Counters counters;
            double* px = accum_sum + 4ll * (y * (int64_t)width + x);
            for (uint32_t k = 0; k < sample_count; ++k) {
                float3 radiance = trace_path(sc, cam, settings, x, y, width, height, int(first_sample + k), counters);
                px[0] += radiance.x; px[1] += radiance.y; px[2] += radiance.z; px[3] += 1.0;
            }
            extend += counters.extend_rays; shadow += counters.shadow_rays;
        }
    if (out_counters) { out_counters[0] += extend; out_counters[1] += shadow; }
}

} // extern "C"
