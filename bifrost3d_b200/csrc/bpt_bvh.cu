// Acceleration structure build on the device (replaces OptiX "Trbvh", Renderer.cpp:161-182,470-477) and the
// batched intersection entry point.
//
// Build = flatten instances to world space -> centroid bounds -> 63-bit Morton codes -> radix sort ->
// Karras' parallel radix tree (HPG 2012) -> bottom-up box fit; its subtrees of at most two triangles (one on scenes above
// 8 M triangles) are the leaf clusters -> PLOC (Meister and Bittner 2018) rebuilds the hierarchy above them, its passes
// inside one cooperative launch -> 64-byte two-child nodes -> collapse to compressed eight-wide nodes (bpt_cw.cuh; the
// triangle array is regrouped per node) for scenes from 131 072 triangles on, to 128-byte four-wide nodes below that.
// Fallbacks: the plain Morton hierarchy when PLOC gives up, narrower nodes when a tree is too deep for a traversal stack.
#include "bpt_context.h"
#include "bpt_trace.cuh"
#include <type_traits>

#include "bpt_sort.cuh"
#include <cooperative_groups.h>

#include <algorithm>
#include <float.h>

namespace bpt {

namespace {

#ifndef BPT_LEAF_MAX
#define BPT_LEAF_MAX 2
#else
#define BPT_LEAF_MAX_FORCED 1
#endif
constexpr int LEAF_MAX = BPT_LEAF_MAX;

struct InstanceRecord {
    int prim_offset;   // first global primitive index
    int prim_count;
    int material;
    uint32_t flags;    // bit0: has normals, bit1: has tints, bit2: has texcoords, bit3: has emission
    // the instance's mesh, resident on the device since bpt_upload_mesh
    const uint32_t* indices;
    const float* positions;
    const int16_t* normals;
    const uint8_t* tints;
    const float2* texcoords;
    const float* emission;
    float m[12];       // object -> world, row-major 3x4
};

struct Aabb { float3 lo, hi; };

__device__ __forceinline__ float3 transform_point(const float* m, float3 p) {
    // ((m0*x + m1*y) + m2*z) + m3, no contraction: the oracle flattens with the same rounding.
    return f3(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], p.x), __fmul_rn(m[1], p.y)), __fmul_rn(m[2], p.z)), m[3]),
              __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], p.x), __fmul_rn(m[5], p.y)), __fmul_rn(m[6], p.z)), m[7]),
              __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], p.x), __fmul_rn(m[9], p.y)), __fmul_rn(m[10], p.z)), m[11]));
}

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// One thread per global primitive: world-space vertices, shading record, centroid bounds.
__global__ void flatten_kernel(int64_t prim_total, int instance_count, const InstanceRecord* __restrict__ instances,
                               float4* __restrict__ world_vertices, ShadeTriangle* __restrict__ shade, float2* __restrict__ shade_uv,
                               float* __restrict__ shade_emission, float* scene_bounds /*[6]*/) {
    float3 lo = f3(FLT_MAX), hi = f3(-FLT_MAX);
    for (int64_t gp = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gp < prim_total; gp += (int64_t)gridDim.x * blockDim.x) {
        // binary search: last instance with prim_offset <= gp
        int a = 0, b = instance_count - 1;
        while (a < b) {
            int mid = (a + b + 1) >> 1;
            if (instances[mid].prim_offset <= gp) a = mid; else b = mid - 1;
        }
        const InstanceRecord& inst = instances[a];
        int lp = int(gp - inst.prim_offset);
        const uint32_t* tri = inst.indices + 3ll * lp;
        const uint32_t i0 = tri[0], i1 = tri[1], i2 = tri[2];
        const float* __restrict__ positions = inst.positions;
        const int16_t* __restrict__ normals = inst.normals;
        const uint8_t* __restrict__ tints = inst.tints;
        const float2* __restrict__ texcoords = inst.texcoords;
        float3 p0 = transform_point(inst.m, f3(positions[3ll * i0], positions[3ll * i0 + 1], positions[3ll * i0 + 2]));
        float3 p1 = transform_point(inst.m, f3(positions[3ll * i1], positions[3ll * i1 + 1], positions[3ll * i1 + 2]));
        float3 p2 = transform_point(inst.m, f3(positions[3ll * i2], positions[3ll * i2 + 1], positions[3ll * i2 + 2]));
        world_vertices[3 * gp] = f4(p0, 0.0f);
        world_vertices[3 * gp + 1] = f4(p1, 0.0f);
        world_vertices[3 * gp + 2] = f4(p2, 0.0f);

        ShadeTriangle s = {};
        if (inst.flags & 1u) {
            s.n0[0] = normals[2ll * i0]; s.n0[1] = normals[2ll * i0 + 1];
            s.n1[0] = normals[2ll * i1]; s.n1[1] = normals[2ll * i1 + 1];
            s.n2[0] = normals[2ll * i2]; s.n2[1] = normals[2ll * i2 + 1];
        }
        if (inst.flags & 2u) {
            for (int k = 0; k < 4; ++k) { s.t0[k] = tints[4ll * i0 + k]; s.t1[k] = tints[4ll * i1 + k]; s.t2[k] = tints[4ll * i2 + k]; }
        }
        s.material_index = inst.material;
        s.flags = (uint32_t(a) << 2) | (inst.flags & 3u);
        shade[gp] = s;
        if (shade_uv != nullptr) { // TriangleAttributes.cu:57-64; meshes without texcoords read (0, 0)
            const bool has_uv = inst.flags & 4u;
            const float2 zero = make_float2(0.0f, 0.0f);
            shade_uv[3 * gp] = has_uv ? texcoords[i0] : zero;
            shade_uv[3 * gp + 1] = has_uv ? texcoords[i1] : zero;
            shade_uv[3 * gp + 2] = has_uv ? texcoords[i2] : zero;
        }

        if (shade_emission != nullptr) { // TriangleAttributes.cu:78-83; meshes without the buffer scale by 1
            const bool has_emission = inst.flags & 8u;
            const uint32_t vertex[3] = { i0, i1, i2 };
            for (int k = 0; k < 3; ++k)
                for (int ch = 0; ch < 3; ++ch)
                    shade_emission[9 * gp + 3 * k + ch] = has_emission ? inst.emission[3ll * vertex[k] + ch] : 1.0f;
        }

        float3 c = (min3(min3(p0, p1), p2) + max3(max3(p0, p1), p2)) * 0.5f;
        lo = min3(lo, c); hi = max3(hi, c);
    }
    // warp reduce, then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o)); lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o)); lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if ((threadIdx.x & 31) == 0 && lo.x <= hi.x) {
        atomic_min_float(scene_bounds + 0, lo.x); atomic_min_float(scene_bounds + 1, lo.y); atomic_min_float(scene_bounds + 2, lo.z);
        atomic_max_float(scene_bounds + 3, hi.x); atomic_max_float(scene_bounds + 4, hi.y); atomic_max_float(scene_bounds + 5, hi.z);
    }
}

__device__ __forceinline__ uint64_t expand_bits_21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void morton_kernel(int64_t n, const float4* __restrict__ world_vertices, const float* __restrict__ scene_bounds,
                              uint64_t* __restrict__ keys, uint32_t* __restrict__ values) {
    float3 lo = f3(scene_bounds[0], scene_bounds[1], scene_bounds[2]);
    float3 hi = f3(scene_bounds[3], scene_bounds[4], scene_bounds[5]);
    float3 extent = max3(hi - lo, f3(1e-30f));
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float3 p0 = f3(world_vertices[3 * i]), p1 = f3(world_vertices[3 * i + 1]), p2 = f3(world_vertices[3 * i + 2]);
        float3 c = (min3(min3(p0, p1), p2) + max3(max3(p0, p1), p2)) * 0.5f;
        float3 q = (c - lo) / extent;
        const float scale = 2097151.0f; // 2^21 - 1
        uint64_t x = (uint64_t)fminf(fmaxf(q.x * scale, 0.0f), scale);
        uint64_t y = (uint64_t)fminf(fmaxf(q.y * scale, 0.0f), scale);
        uint64_t z = (uint64_t)fminf(fmaxf(q.z * scale, 0.0f), scale);
        keys[i] = (expand_bits_21(x) << 2) | (expand_bits_21(y) << 1) | expand_bits_21(z);
        values[i] = (uint32_t)i;
    }
}

// Karras 2012. Internal node i in [0, n-2]; leaves are sorted positions [0, n-1].
struct TreeNode {
    int left, right;   // >= 0: internal node, < 0: ~leaf position
    int first, last;   // covered range of sorted positions
};

__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void hierarchy_kernel(int n, const uint64_t* __restrict__ keys, TreeNode* __restrict__ tree, int* __restrict__ parent_of_internal,
                                 int* __restrict__ parent_of_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int delta_min = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > delta_min) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > delta_min) l += t;
    int j = i + l * d;
    int delta_node = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > delta_node) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    TreeNode node;
    node.first = first; node.last = last;
    if (first == gamma) { node.left = ~gamma; parent_of_leaf[gamma] = i; } else { node.left = gamma; parent_of_internal[gamma] = i; }
    if (last == gamma + 1) { node.right = ~(gamma + 1); parent_of_leaf[gamma + 1] = i; } else { node.right = gamma + 1; parent_of_internal[gamma + 1] = i; }
    tree[i] = node;
    if (i == 0) parent_of_internal[0] = -1;
}

__device__ __forceinline__ Aabb load_box_cg(const Aabb* p) {
    const float* f = reinterpret_cast<const float*>(p);
    Aabb b;
    b.lo = f3(__ldcg(f), __ldcg(f + 1), __ldcg(f + 2));
    b.hi = f3(__ldcg(f + 3), __ldcg(f + 4), __ldcg(f + 5));
    return b;
}

// Writes the Morton-ordered triangle array and the leaf boxes, then climbs: the second thread to reach an
// internal node merges its children's boxes (classic atomic-counter refit).
__global__ void fit_kernel(int n, const uint32_t* __restrict__ sorted_prims, const float4* __restrict__ world_vertices,
                           const ShadeTriangle* __restrict__ shade,
                           TraceTriangle* __restrict__ triangles, uint32_t* __restrict__ slot_of_primitive, Aabb* __restrict__ leaf_boxes,
                           Aabb* __restrict__ node_boxes,
                           const TreeNode* __restrict__ tree, const int* __restrict__ parent_of_internal, const int* __restrict__ parent_of_leaf,
                           int* __restrict__ arrival) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t gp = sorted_prims[p];
    slot_of_primitive[gp] = (uint32_t)p;
    float4 v0 = world_vertices[3ll * gp], v1 = world_vertices[3ll * gp + 1], v2 = world_vertices[3ll * gp + 2];
    int material = shade[gp].material_index;
    TraceTriangle t;
    t.v0 = make_float4(v0.x, v0.y, v0.z, __int_as_float((int)gp));
    t.v1 = make_float4(v1.x, v1.y, v1.z, __int_as_float(material));
    t.v2 = make_float4(v2.x, v2.y, v2.z, 0.0f);
    triangles[p] = t;
    Aabb box;
    box.lo = min3(min3(f3(v0), f3(v1)), f3(v2));
    box.hi = max3(max3(f3(v0), f3(v1)), f3(v2));
    leaf_boxes[p] = box;
    if (n == 1) return;

    int node = parent_of_leaf[p];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(arrival + node, 1) == 0)
            return; // first arrival: the sibling subtree is not finished yet
        TreeNode tn = tree[node];
        // L1 is not coherent across SMs: read the children's boxes from L2 (ld.global.cg). The sibling published
        // its box before its atomicAdd (threadfence above).
        Aabb l = load_box_cg(tn.left < 0 ? leaf_boxes + ~tn.left : node_boxes + tn.left);
        Aabb r = load_box_cg(tn.right < 0 ? leaf_boxes + ~tn.right : node_boxes + tn.right);
        Aabb merged;
        merged.lo = min3(l.lo, r.lo); merged.hi = max3(l.hi, r.hi);
        node_boxes[node] = merged;
        node = parent_of_internal[node];
    }
}

__global__ void emit_kernel(int n, int leaf_max, const TreeNode* __restrict__ tree, const Aabb* __restrict__ leaf_boxes, const Aabb* __restrict__ node_boxes,
                            BvhNode* __restrict__ nodes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    TreeNode tn = tree[i];
    int link[2];
    Aabb box[2];
    int child[2] = { tn.left, tn.right };
    for (int k = 0; k < 2; ++k) {
        if (child[k] < 0) {
            link[k] = pack_leaf(~child[k], 1); box[k] = leaf_boxes[~child[k]];
        } else {
            TreeNode c = tree[child[k]];
            int size = c.last - c.first + 1;
            box[k] = node_boxes[child[k]];
            link[k] = size <= leaf_max ? pack_leaf(c.first, size) : child[k];
        }
    }
    BvhNode out;
    out.lo_l_hi_l_x = make_float4(box[0].lo.x, box[0].lo.y, box[0].lo.z, box[0].hi.x);
    out.hi_l_lo_r = make_float4(box[0].hi.y, box[0].hi.z, box[1].lo.x, box[1].lo.y);
    out.lo_r_hi_r = make_float4(box[1].lo.z, box[1].hi.x, box[1].hi.y, box[1].hi.z);
    out.left = link[0]; out.right = link[1]; out.pad0 = 0; out.pad1 = 0;
    nodes[i] = out;
}

// ---- PLOC: parallel locally-ordered clustering (Meister and Bittner, TVCG 2018) over the LBVH's leaf clusters ----------
// The Morton-ordered LBVH above is kept for its bottom: its subtrees of at most LEAF_MAX triangles become the leaf
// clusters (contiguous ranges of the sorted triangle array). The hierarchy ABOVE them is rebuilt bottom-up: every
// cluster looks for the neighbour within PLOC_RADIUS positions whose union with it has the smallest surface area,
// mutual nearest neighbours merge into a node, the cluster list is compacted (order preserving) and the search repeats.
// This follows the surface area heuristic locally instead of the Morton code's bit pattern: fewer node visits per ray.
#ifndef BPT_PLOC_RADIUS
#define BPT_PLOC_RADIUS 16
#else
#define BPT_PLOC_RADIUS_FORCED 1
#endif
constexpr int PLOC_RADIUS = BPT_PLOC_RADIUS;
constexpr int PLOC_MAX_DEPTH = 96; // the traversal stack holds STACK_SMEM + STACK_LOCAL = 104 entries
constexpr int PLOC_TAIL = 1024;    // the last clusters finish inside one block (ploc_tail_kernel)

__device__ __forceinline__ bool is_leaf_cluster_child(const TreeNode* __restrict__ tree, int child, int leaf_max, int& first, int& size) {
    if (child < 0) { first = ~child; size = 1; return true; }
    TreeNode c = tree[child];
    first = c.first; size = c.last - c.first + 1;
    return size <= leaf_max;
}

// flag[p] = 1 where a leaf cluster starts (p = position in the sorted triangle array).
__global__ void ploc_mark_kernel(int n, int leaf_max, const TreeNode* __restrict__ tree, uint32_t* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    TreeNode tn = tree[i];
    if (i != 0 && tn.last - tn.first + 1 <= leaf_max) return; // inside a leaf cluster
    int first, size;
    if (is_leaf_cluster_child(tree, tn.left, leaf_max, first, size)) flag[first] = 1u;
    if (is_leaf_cluster_child(tree, tn.right, leaf_max, first, size)) flag[first] = 1u;
}

__global__ void ploc_gather_kernel(int n, int leaf_max, const TreeNode* __restrict__ tree, const Aabb* __restrict__ leaf_boxes, const Aabb* __restrict__ node_boxes,
                                   const uint32_t* __restrict__ position, int* __restrict__ cl_link, Aabb* __restrict__ cl_box, int* __restrict__ cl_depth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    TreeNode tn = tree[i];
    if (i != 0 && tn.last - tn.first + 1 <= leaf_max) return;
    int child[2] = { tn.left, tn.right };
    for (int k = 0; k < 2; ++k) {
        int first, size;
        if (!is_leaf_cluster_child(tree, child[k], leaf_max, first, size)) continue;
        uint32_t c = position[first];
        cl_link[c] = pack_leaf(first, size);
        cl_box[c] = child[k] < 0 ? leaf_boxes[~child[k]] : node_boxes[child[k]];
        cl_depth[c] = 0;
    }
}

__device__ __forceinline__ float union_area(const Aabb& a, const Aabb& b) {
    float3 d = max3(a.hi, b.hi) - min3(a.lo, b.lo);
    return d.x * d.y + d.y * d.z + d.z * d.x;
}

// What a PLOC pass needs to know, kept in device memory so that the passes can follow each other without the host: the
// number of clusters, which of the two cluster lists is current, and how it went.
struct PlocState { uint32_t m; int cur; int passes; int failed; };
struct PlocLists { int* link[2]; Aabb* box[2]; int* depth[2]; int radius; /* positions searched to either side */ };

// The three steps of a pass for one cluster; the stream-launch kernels and the persistent kernel below share them.
__device__ __forceinline__ int ploc_nearest_of(int i, int m, int radius, const Aabb* __restrict__ cl_box) {
    const Aabb mine = cl_box[i];
    float best = FLT_MAX; int best_j = -1, best_rank = 0x7fffffff;
    const int lo = max(0, i - radius), hi = min(m - 1, i + radius);
    for (int j = lo; j <= hi; ++j) {
        if (j == i) continue;
        float a = union_area(mine, cl_box[j]);
        // exact ties (regular or coincident geometry): prefer the aligned partner i ^ 1, then the closer position, then the
        // lower one, so that equal boxes still pair up instead of forming a chain that merges one pair per pass
        int rank = (j == (i ^ 1)) ? 0 : 2 * abs(j - i) + (j > i ? 1 : 0);
        if (a < best || (a == best && rank < best_rank)) { best = a; best_j = j; best_rank = rank; }
    }
    return best_j;
}

// Mutual nearest neighbours merge: the lower position keeps the new node, the higher one is dropped. Returns the keep flag.
__device__ __forceinline__ uint32_t ploc_merge_at(int i, const int* __restrict__ nearest, int* __restrict__ cl_link, Aabb* __restrict__ cl_box, int* __restrict__ cl_depth,
                                                  BvhNode* __restrict__ nodes, int* __restrict__ node_counter, int* __restrict__ max_depth) {
    int j = nearest[i];
    bool mutual = j >= 0 && nearest[j] == i;
    if (!mutual) return 1u;
    if (i > j) return 0u;
    Aabb a = cl_box[i], b = cl_box[j];
    int index = atomicAdd(node_counter, 1);
    BvhNode out;
    out.lo_l_hi_l_x = make_float4(a.lo.x, a.lo.y, a.lo.z, a.hi.x);
    out.hi_l_lo_r = make_float4(a.hi.y, a.hi.z, b.lo.x, b.lo.y);
    out.lo_r_hi_r = make_float4(b.lo.z, b.hi.x, b.hi.y, b.hi.z);
    out.left = cl_link[i]; out.right = cl_link[j]; out.pad0 = 0; out.pad1 = 0;
    nodes[index] = out;
    Aabb merged; merged.lo = min3(a.lo, b.lo); merged.hi = max3(a.hi, b.hi);
    int depth = max(cl_depth[i], cl_depth[j]) + 1;
    // cluster j is only read by this thread (its own thread returned above), so updating slot i in place is race free: no
    // other cluster has i or j as a MUTUAL partner, and non-mutual clusters only read `nearest`.
    cl_box[i] = merged; cl_link[i] = index; cl_depth[i] = depth;
    atomicMax(max_depth, depth);
    return 1u;
}

__global__ void ploc_nearest_kernel(const PlocState* __restrict__ state, PlocLists lists, int* __restrict__ nearest) {
    const int m = (int)state->m;
    const Aabb* __restrict__ cl_box = lists.box[state->cur];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) nearest[i] = ploc_nearest_of(i, m, lists.radius, cl_box);
}

__global__ void ploc_merge_kernel(const PlocState* __restrict__ state, PlocLists lists, const int* __restrict__ nearest, uint32_t* __restrict__ keep,
                                  BvhNode* __restrict__ nodes, int* __restrict__ node_counter, int* __restrict__ max_depth) {
    const int m = (int)state->m, cur = state->cur;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
        keep[i] = ploc_merge_at(i, nearest, lists.link[cur], lists.box[cur], lists.depth[cur], nodes, node_counter, max_depth);
}

__global__ void ploc_compact_kernel(const PlocState* __restrict__ state, PlocLists lists, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ position) {
    const int m = (int)state->m, cur = state->cur;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        if (!keep[i]) continue;
        uint32_t c = position[i];
        lists.link[cur ^ 1][c] = lists.link[cur][i]; lists.box[cur ^ 1][c] = lists.box[cur][i]; lists.depth[cur ^ 1][c] = lists.depth[cur][i];
    }
}

// Ends a pass of the stream-launch form: the compacted list becomes the current one (the host reads the state back).
__global__ void ploc_advance_kernel(PlocState* __restrict__ state, const uint32_t* __restrict__ kept) {
    const uint32_t next_m = *kept;
    if (next_m >= state->m || state->passes >= 4096) state->failed = 1; // cannot happen: the globally closest pair is always mutual
    else { state->m = next_m; state->cur ^= 1; state->passes += 1; }
}

__global__ void ploc_begin_kernel(PlocState* __restrict__ state, uint32_t m) {
    state->m = m; state->cur = 0; state->passes = 0; state->failed = 0;
}

// All passes in ONE cooperative launch: the grid stays resident and three grid-wide barriers separate the steps of a pass
// (nearest | merge + count | compact). Every block owns a contiguous chunk of the cluster list, so the order-preserving
// compaction is a block-local scan plus the sum of the preceding blocks' counts, which every block adds up for itself -
// as it does the grand total, so all blocks know the next pass's cluster count without reading it back from anywhere.
// Measured on B200 at 1 M triangles (33 passes): 1.26 ms; the five stream kernels of a pass as one graph with a conditional
// WHILE node 2.05 ms (62 us per pass, mostly node-to-node latency), as stream launches with a host read per pass 2.4 ms.
constexpr int PLOC_BLOCK = 512;

__device__ __forceinline__ uint32_t ploc_block_exclusive_scan(uint32_t v, uint32_t* warp_sums /*[PLOC_BLOCK / 32]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inclusive = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inclusive, o); if (lane >= o) inclusive += t; }
    if (lane == 31) warp_sums[warp] = inclusive;
    __syncthreads();
    uint32_t before = 0; total = 0;
#pragma unroll
    for (int w = 0; w < PLOC_BLOCK / 32; ++w) { const uint32_t s = warp_sums[w]; if (w < warp) before += s; total += s; }
    __syncthreads();
    return before + inclusive - v;
}

__global__ void __launch_bounds__(PLOC_BLOCK) ploc_persistent_kernel(PlocState* __restrict__ state, PlocLists lists, int* __restrict__ nearest, uint32_t* __restrict__ keep,
                                                                     uint32_t* __restrict__ block_counts, BvhNode* __restrict__ nodes, int* __restrict__ node_counter,
                                                                     int* __restrict__ max_depth) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ uint32_t s_warp_sums[PLOC_BLOCK / 32];
    __shared__ uint32_t s_reduce[2];
    int m = (int)state->m, cur = state->cur, passes = state->passes;
    bool failed = state->failed != 0;
    grid.sync(); // every block has read the state that block 0 rewrites at the end
    while (m > PLOC_TAIL && !failed) {
        const Aabb* __restrict__ cl_box = lists.box[cur];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) nearest[i] = ploc_nearest_of(i, m, lists.radius, cl_box);
        grid.sync();
        // this block's chunk of the list
        const int chunk = (m + (int)gridDim.x - 1) / (int)gridDim.x;
        const int first = min((int)blockIdx.x * chunk, m), last = min(first + chunk, m);
        uint32_t kept_here = 0;
        for (int base = first; base < last; base += PLOC_BLOCK) {
            const int i = base + (int)threadIdx.x;
            uint32_t k = 0;
            if (i < last) { k = ploc_merge_at(i, nearest, lists.link[cur], lists.box[cur], lists.depth[cur], nodes, node_counter, max_depth); keep[i] = k; }
            kept_here += (uint32_t)__syncthreads_count((int)k);
        }
        if (threadIdx.x == 0) block_counts[blockIdx.x] = kept_here;
        grid.sync();
        // clusters kept by the blocks in front of this one, and by all of them
        uint32_t before = 0, all = 0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += PLOC_BLOCK) { const uint32_t c = block_counts[b]; all += c; if (b < (int)blockIdx.x) before += c; }
        {
            uint32_t total_before, total_all;
            ploc_block_exclusive_scan(before, s_warp_sums, total_before);
            ploc_block_exclusive_scan(all, s_warp_sums, total_all);
            if (threadIdx.x == 0) { s_reduce[0] = total_before; s_reduce[1] = total_all; }
            __syncthreads();
        }
        uint32_t running = s_reduce[0];
        const int next_m = (int)s_reduce[1];
        for (int base = first; base < last; base += PLOC_BLOCK) {
            const int i = base + (int)threadIdx.x;
            const uint32_t k = i < last ? keep[i] : 0u;
            uint32_t tile_total;
            const uint32_t c = running + ploc_block_exclusive_scan(k, s_warp_sums, tile_total);
            if (k) { lists.link[cur ^ 1][c] = lists.link[cur][i]; lists.box[cur ^ 1][c] = lists.box[cur][i]; lists.depth[cur ^ 1][c] = lists.depth[cur][i]; }
            running += tile_total;
        }
        if (next_m >= m || passes >= 4096) { failed = true; break; } // cannot happen; every block takes the same branch
        m = next_m; cur ^= 1; ++passes;
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { state->m = (uint32_t)m; state->cur = cur; state->passes = passes; state->failed = failed ? 1 : 0; }
}

// The last PLOC_TAIL clusters finish inside one block: the same search / merge / compact passes as above on shared memory,
// without a kernel launch per pass (the top of the tree takes ~50 passes that merge a few pairs each).
__global__ void __launch_bounds__(PLOC_TAIL) ploc_tail_kernel(const PlocState* __restrict__ state, PlocLists lists,
                                                              BvhNode* __restrict__ nodes, int* __restrict__ node_counter, int* __restrict__ max_depth) {
    const int m = (int)state->m;
    if (state->failed || m > PLOC_TAIL) return;
    const int* __restrict__ cl_link = lists.link[state->cur]; const Aabb* __restrict__ cl_box = lists.box[state->cur]; const int* __restrict__ cl_depth = lists.depth[state->cur];
    __shared__ Aabb s_box[PLOC_TAIL];
    __shared__ int s_link[PLOC_TAIL], s_depth[PLOC_TAIL], s_nearest[PLOC_TAIL];
    __shared__ int s_warp_sums[32];
    const int i = threadIdx.x;
    if (i < m) { s_box[i] = cl_box[i]; s_link[i] = cl_link[i]; s_depth[i] = cl_depth[i]; }
    int count = m;
    __syncthreads();
    while (count > 1) {
        if (i < count) {
            const Aabb mine = s_box[i];
            float best = FLT_MAX; int best_j = -1, best_rank = 0x7fffffff;
            const int lo = max(0, i - lists.radius), hi = min(count - 1, i + lists.radius);
            for (int j = lo; j <= hi; ++j) {
                if (j == i) continue;
                float a = union_area(mine, s_box[j]);
                int rank = (j == (i ^ 1)) ? 0 : 2 * abs(j - i) + (j > i ? 1 : 0);
                if (a < best || (a == best && rank < best_rank)) { best = a; best_j = j; best_rank = rank; }
            }
            s_nearest[i] = best_j;
        }
        __syncthreads();
        int keep = 0, link = 0, depth = 0;
        Aabb box = {};
        if (i < count) {
            const int j = s_nearest[i];
            const bool mutual = s_nearest[j] == i;
            box = s_box[i]; link = s_link[i]; depth = s_depth[i];
            keep = (!mutual || i < j) ? 1 : 0;
            if (mutual && i < j) {
                const Aabb other = s_box[j];
                const int index = atomicAdd(node_counter, 1);
                BvhNode out;
                out.lo_l_hi_l_x = make_float4(box.lo.x, box.lo.y, box.lo.z, box.hi.x);
                out.hi_l_lo_r = make_float4(box.hi.y, box.hi.z, other.lo.x, other.lo.y);
                out.lo_r_hi_r = make_float4(other.lo.z, other.hi.x, other.hi.y, other.hi.z);
                out.left = link; out.right = s_link[j]; out.pad0 = 0; out.pad1 = 0;
                nodes[index] = out;
                box.lo = min3(box.lo, other.lo); box.hi = max3(box.hi, other.hi);
                depth = max(depth, s_depth[j]) + 1;
                link = index;
                atomicMax(max_depth, depth);
            }
        }
        // exclusive scan of `keep` over the 32 warps of the block: shuffle scan inside each warp, then over the warp totals
        int position, total;
        {
            const int lane = i & 31, warp = i >> 5;
            int inclusive = keep;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inclusive, o); if (lane >= o) inclusive += t; }
            if (lane == 31) s_warp_sums[warp] = inclusive;
            __syncthreads();
            if (warp == 0) {
                int w = s_warp_sums[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
                s_warp_sums[lane] = w;
            }
            __syncthreads();
            position = (warp ? s_warp_sums[warp - 1] : 0) + inclusive - keep;
            total = s_warp_sums[31];
        }
        __syncthreads(); // every read of the old cluster list (and of the warp totals) is done
        if (keep) { s_box[position] = box; s_link[position] = link; s_depth[position] = depth; }
        count = total;
        __syncthreads();
    }
    if (i == 0) {
        __threadfence();
        nodes[0] = nodes[s_link[0]]; // the traversal enters at node 0
    }
}


// ---- four-wide collapse ---------------------------------------------------------------------------------------------
// Top down over the finished binary hierarchy: a wide node starts from a binary node's two children and twice replaces the
// inner child with the largest surface area by that child's own two children. One kernel launch per level of the wide
// tree; the tasks of the next level are appended with an atomic counter.
struct WideTask { int wide_index, binary_index; };

__device__ __forceinline__ float box_area(float3 lo, float3 hi) {
    float3 d = hi - lo;
    return d.x * d.y + d.y * d.z + d.z * d.x;
}

__global__ void collapse_kernel(int count, const WideTask* __restrict__ tasks, WideTask* __restrict__ next_tasks, int* __restrict__ counters /*[0] next, [1] wide nodes*/,
                                const BvhNode* __restrict__ nodes, WideNode* __restrict__ wide) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const WideTask task = tasks[i];
    const float far = 3.0e38f;
    int link[4] = { NODE_EMPTY, NODE_EMPTY, NODE_EMPTY, NODE_EMPTY };
    float3 lo[4] = { f3(far), f3(far), f3(far), f3(far) }, hi[4] = { f3(far), f3(far), f3(far), f3(far) };
    auto children = [&](int binary_index, int a, int b) {
        BvhNode n = nodes[binary_index];
        lo[a] = f3(n.lo_l_hi_l_x.x, n.lo_l_hi_l_x.y, n.lo_l_hi_l_x.z); hi[a] = f3(n.lo_l_hi_l_x.w, n.hi_l_lo_r.x, n.hi_l_lo_r.y); link[a] = n.left;
        lo[b] = f3(n.hi_l_lo_r.z, n.hi_l_lo_r.w, n.lo_r_hi_r.x); hi[b] = f3(n.lo_r_hi_r.y, n.lo_r_hi_r.z, n.lo_r_hi_r.w); link[b] = n.right;
    };
    children(task.binary_index, 0, 1);
    for (int used = 2; used < 4; ++used) {
        int best = -1; float best_area = -1.0f;
        for (int k = 0; k < used; ++k)
            if (link[k] >= 0) { float area = box_area(lo[k], hi[k]); if (area > best_area) { best_area = area; best = k; } }
        if (best < 0) break;
        children(link[best], best, used);
    }
    for (int k = 0; k < 4; ++k)
        if (link[k] >= 0) {
            int w = atomicAdd(counters + 1, 1);
            next_tasks[atomicAdd(counters, 1)] = { w, link[k] };
            link[k] = w;
        }
    WideNode out;
    out.lo_x = make_float4(lo[0].x, lo[1].x, lo[2].x, lo[3].x); out.lo_y = make_float4(lo[0].y, lo[1].y, lo[2].y, lo[3].y);
    out.lo_z = make_float4(lo[0].z, lo[1].z, lo[2].z, lo[3].z); out.hi_x = make_float4(hi[0].x, hi[1].x, hi[2].x, hi[3].x);
    out.hi_y = make_float4(hi[0].y, hi[1].y, hi[2].y, hi[3].y); out.hi_z = make_float4(hi[0].z, hi[1].z, hi[2].z, hi[3].z);
    out.link = make_int4(link[0], link[1], link[2], link[3]);
    out.pad = make_int4(0, 0, 0, 0);
    wide[task.wide_index] = out;
}

__global__ void tiny_root_kernel(int n, const Aabb* __restrict__ leaf_boxes, BvhNode* __restrict__ nodes) {
    // n == 0: both children absent. n == 1: the left child is the only triangle. An absent child is a point box at
    // (3e38, 3e38, 3e38): for a normalised direction its slab distances are >= 3e38 in magnitude, outside any [tmin, tmax].
    const float far = 3.0e38f;
    BvhNode out;
    out.lo_l_hi_l_x = make_float4(far, far, far, far);
    out.hi_l_lo_r = make_float4(far, far, far, far);
    out.lo_r_hi_r = make_float4(far, far, far, far);
    out.left = out.right = NODE_EMPTY; out.pad0 = out.pad1 = 0;
    if (n == 1) {
        Aabb b = leaf_boxes[0];
        out.lo_l_hi_l_x = make_float4(b.lo.x, b.lo.y, b.lo.z, b.hi.x);
        out.hi_l_lo_r = make_float4(b.hi.y, b.hi.z, far, far);
        out.left = pack_leaf(0, 1);
    }
    nodes[0] = out;
}

// ---- compressed eight-wide collapse ------------------------------------------------------------------------------------
// Top down like the four-wide collapse, but a node keeps opening its inner child with the largest surface area until it has
// eight children (or only leaves). bpt_cw.cuh encodes the node (octant slots, quantised boxes); here the node's inner
// children get consecutive node indices and the triangles of its leaf children consecutive places in a second triangle
// array, both in slot order, which is what lets the node address them with two base indices and a few bits per child.
// The tasks of the next level sit at (node index - first node index of that level).
__global__ void cw_collapse_kernel(int count, int next_level_base, const WideTask* __restrict__ tasks, WideTask* __restrict__ next_tasks,
                                   int* __restrict__ counters /*[0] nodes allocated, [1] triangles placed*/, const BvhNode* __restrict__ nodes,
                                   CwNode* __restrict__ cw, const TraceTriangle* __restrict__ triangles_in, TraceTriangle* __restrict__ triangles_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const WideTask task = tasks[i];
    int link[CW_WIDTH]; float3 lo[CW_WIDTH], hi[CW_WIDTH];
    int k = 0;
    auto open = [&](int binary_index, int a, int b) {
        const BvhNode n = nodes[binary_index];
        lo[a] = f3(n.lo_l_hi_l_x.x, n.lo_l_hi_l_x.y, n.lo_l_hi_l_x.z); hi[a] = f3(n.lo_l_hi_l_x.w, n.hi_l_lo_r.x, n.hi_l_lo_r.y); link[a] = n.left;
        lo[b] = f3(n.hi_l_lo_r.z, n.hi_l_lo_r.w, n.lo_r_hi_r.x); hi[b] = f3(n.lo_r_hi_r.y, n.lo_r_hi_r.z, n.lo_r_hi_r.w); link[b] = n.right;
    };
    open(task.binary_index, 0, 1); k = 2;
    while (k < CW_WIDTH) {
        int best = -1; float best_area = -1.0f;
        for (int c = 0; c < k; ++c)
            if (link[c] >= 0) { float area = box_area(lo[c], hi[c]); if (area > best_area) { best_area = area; best = c; } }
        if (best < 0) break;
        open(link[best], best, k); ++k;
    }
    int triangle_counts[CW_WIDTH];
    for (int c = 0; c < k; ++c) triangle_counts[c] = link[c] >= 0 ? 0 : leaf_count(link[c]);
    CwNode node; CwPlacement place;
    cw_encode(k, lo, hi, triangle_counts, node, place);
    node.child_base = place.inner_count ? (uint32_t)atomicAdd(counters, place.inner_count) : 0u;
    node.triangle_base = place.triangle_count ? (uint32_t)atomicAdd(counters + 1, place.triangle_count) : 0u;
    for (int c = 0; c < k; ++c) {
        if (link[c] >= 0) {
            const int child = (int)node.child_base + place.offset[c];
            next_tasks[child - next_level_base] = { child, link[c] };
        } else {
            const int first = leaf_first(link[c]);
            for (int t = 0; t < triangle_counts[c]; ++t) triangles_out[node.triangle_base + place.offset[c] + t] = triangles_in[first + t];
        }
    }
    cw[task.wide_index] = node;
}

// ---- batched queries -----------------------------------------------------------------------------

struct BatchSource {
    const float* origins; const float* directions; const float* tmin; const float* tmax;
    int32_t* out_primitive; float* out_t; float* out_uv; uint8_t* out_occluded;
    __device__ void load(unsigned int i, Ray& ray, int& skip) const {
        ray.origin = f3(origins[3ll * i], origins[3ll * i + 1], origins[3ll * i + 2]);
        ray.direction = f3(directions[3ll * i], directions[3ll * i + 1], directions[3ll * i + 2]);
        ray.tmin = tmin[i]; ray.tmax = tmax[i];
        skip = -1;
    }
    __device__ float termination_weight(unsigned int) const { return 1.0f; }
    template <class Trav>
    __device__ void store_closest(unsigned int i, const Trav& tr) const {
        Hit h = tr.result();
        if (out_primitive) out_primitive[i] = h.primitive;
        if (out_t) out_t[i] = h.primitive >= 0 ? h.t : INFINITY;
        if (out_uv) { out_uv[2ll * i] = h.u; out_uv[2ll * i + 1] = h.v; }
    }
    __device__ void store(unsigned int i, const Traversal<false>& tr) const { store_closest(i, tr); }
    __device__ void store(unsigned int i, const TraversalCW<false>& tr) const { store_closest(i, tr); }
    __device__ void store(unsigned int i, const Traversal<true>& tr) const { out_occluded[i] = tr.transmission < 1.0f ? 1 : 0; }
    __device__ void store(unsigned int i, const TraversalCW<true>& tr) const { out_occluded[i] = tr.transmission < 1.0f ? 1 : 0; }
};

template <bool ANY_HIT, bool COMPRESSED>
__global__ void __launch_bounds__(TRACE_BLOCK) intersect_kernel(AccelView accel, unsigned int n, BatchSource source, const float* __restrict__ coverage,
                                                                unsigned int* fetch_counter) {
    __shared__ __align__(16) int s_stack[STACK_SMEM * TRACE_BLOCK];
    typedef typename TraversalFor<ANY_HIT, COMPRESSED>::type Trav;
    traverse_queue_with<ANY_HIT, Trav>(accel, coverage, source, n, fetch_counter, s_stack + threadIdx.x, accel.budget);
}

} // namespace

int build_accel(Context* ctx) {
    Accel& A = ctx->accel;
    A.valid = false;
    cudaStream_t st = ctx->stream;

    // ---- lay out the instances over the device-resident meshes ----
    std::vector<InstanceRecord> records;
    std::vector<float> h_normal_matrices; // 9 floats per record, row-major
    bool any_texcoords = false;
    if (ctx->has_textured_materials)
        for (const bpt_instance& inst : ctx->instances) any_texcoords |= ctx->meshes[inst.mesh_id].texcoords.size != 0;
    bool any_emission = false;
    for (const bpt_instance& inst : ctx->instances) any_emission |= ctx->meshes[inst.mesh_id].emission.size != 0;
    int64_t prim_total = 0;
    for (const bpt_instance& inst : ctx->instances) {
        const DeviceMesh& mesh = ctx->meshes[inst.mesh_id];
        if (inst.material_id < 0 || inst.material_id >= (int)ctx->host_materials.size())
            return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_build_accel: instance references a material that was not uploaded");
        if (mesh.primitive_count == 0) continue;
        InstanceRecord r = {};
        r.prim_offset = int(prim_total); r.prim_count = mesh.primitive_count;
        r.material = inst.material_id;
        r.flags = (mesh.normals.size ? 1u : 0u) | (mesh.tints.size ? 2u : 0u) | ((any_texcoords && mesh.texcoords.size) ? 4u : 0u) | (mesh.emission.size ? 8u : 0u);
        r.indices = mesh.indices.ptr; r.positions = mesh.positions.ptr; r.normals = mesh.normals.ptr; r.tints = mesh.tints.ptr;
        r.texcoords = reinterpret_cast<const float2*>(mesh.texcoords.ptr); r.emission = mesh.emission.ptr;
        memcpy(r.m, inst.to_world, sizeof(r.m));
        records.push_back(r);
        { // normal matrix = inverse transpose of the upper 3x3 (rtTransformNormal, MonteCarlo.cu:147,176), in double
            const float* m = inst.to_world;
            double a[9] = { m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10] };
            double c[9] = { a[4] * a[8] - a[5] * a[7], a[5] * a[6] - a[3] * a[8], a[3] * a[7] - a[4] * a[6],
                            a[2] * a[7] - a[1] * a[8], a[0] * a[8] - a[2] * a[6], a[1] * a[6] - a[0] * a[7],
                            a[1] * a[5] - a[2] * a[4], a[2] * a[3] - a[0] * a[5], a[0] * a[4] - a[1] * a[3] };
            double det = a[0] * c[0] + a[1] * c[1] + a[2] * c[2];
            for (int k = 0; k < 9; ++k) h_normal_matrices.push_back(float(det != 0.0 ? c[k] / det : (k % 4 == 0 ? 1.0 : 0.0)));
        }
        prim_total += mesh.primitive_count;
        if (prim_total > MAX_TRIANGLES)
            return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_build_accel: more than 2^28 - 1 triangles");
    }
    const int n = int(prim_total);


    DeviceBuffer<InstanceRecord> d_records; DeviceBuffer<float> d_bounds;
    DeviceBuffer<uint64_t> d_keys, d_keys_alt; DeviceBuffer<uint32_t> d_vals, d_vals_alt; DeviceBuffer<uint32_t> d_temp;
    DeviceBuffer<TreeNode> d_tree; DeviceBuffer<int> d_parent_internal, d_parent_leaf, d_arrival; DeviceBuffer<Aabb> d_leaf_boxes, d_node_boxes;
    // PLOC scratch
    DeviceBuffer<uint32_t> d_flag, d_pos; DeviceBuffer<int> d_link[2], d_depth[2], d_nearest, d_scalars; DeviceBuffer<Aabb> d_box[2];
    DeviceBuffer<uint32_t> d_scan_temp, d_scan_total, d_block_counts; DeviceBuffer<PlocState> d_ploc_state;
    DeviceBuffer<WideTask> d_tasks[2]; DeviceBuffer<int> d_counters; // four- and eight-wide collapse
    DeviceBuffer<TraceTriangle> d_triangles_by_node;                 // the triangle array in the order of the eight-wide nodes
    // Leaf clusters: at most two triangles; one on the scenes of the large-scene regime (bpt_trace.cuh: more than 8 M triangles),
    // where tighter leaves measure 2.8 % faster (50 M triangles: 596 against 580 Msamples/s) for 70 % more nodes and a 19 %
    // longer build. BPT_LEAF_MAX at compile time overrides both.
#ifdef BPT_LEAF_MAX_FORCED
    const int leaf_max = LEAF_MAX;
#else
    const int leaf_max = prim_total > 8000000ll ? 1 : LEAF_MAX;
#endif
    const bool try_cw = prim_total >= ctx->cw_min_triangles && leaf_max <= CW_MAX_LEAF_TRIANGLES;
    auto release_all = [&]() {
        d_tasks[0].release(); d_tasks[1].release(); d_counters.release(); d_triangles_by_node.release();
        d_flag.release(); d_pos.release(); d_nearest.release(); d_scalars.release(); d_scan_temp.release(); d_scan_total.release(); d_block_counts.release(); d_ploc_state.release();
        for (int k = 0; k < 2; ++k) { d_link[k].release(); d_depth[k].release(); d_box[k].release(); }
        d_records.release(); d_bounds.release();
        d_keys.release(); d_keys_alt.release(); d_vals.release(); d_vals_alt.release(); d_temp.release();
        d_tree.release(); d_parent_internal.release(); d_parent_leaf.release(); d_arrival.release(); d_leaf_boxes.release(); d_node_boxes.release();
    };
#define BUILD_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { release_all(); return ctx->cuda_fail(_e, #expr); } } while (0)

    auto up = [&](auto& buf, const auto& host) -> cudaError_t {
        cudaError_t e = buf.resize(std::max<size_t>(host.size(), 1));
        if (e != cudaSuccess || host.empty()) return e;
        return cudaMemcpyAsync(buf.ptr, host.data(), host.size() * sizeof(host[0]), cudaMemcpyHostToDevice, st);
    };
    BUILD_CHECK(up(d_records, records));
    BUILD_CHECK(up(A.normal_matrices, h_normal_matrices));
    float init_bounds[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
    BUILD_CHECK(d_bounds.resize(6));
    BUILD_CHECK(cudaMemcpyAsync(d_bounds.ptr, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, st));

    BUILD_CHECK(A.world_vertices.resize(std::max<size_t>(3ull * n, 1)));
    BUILD_CHECK(A.shade.resize(std::max<size_t>(n, 1)));
    A.has_uv = any_texcoords && n > 0;
    A.built_for_textures = ctx->has_textured_materials;
    if (A.has_uv) BUILD_CHECK(A.shade_uv.resize(3ull * n)); else A.shade_uv.release();
    A.has_emission = any_emission && n > 0;
    if (A.has_emission) BUILD_CHECK(A.shade_emission.resize(9ull * n)); else A.shade_emission.release();
    BUILD_CHECK(A.triangles.resize(std::max<size_t>(n, 1)));
    BUILD_CHECK(A.slot_of_primitive.resize(std::max<size_t>(n, 1)));
    BUILD_CHECK(A.nodes.resize((size_t)n + 1));
    // PLOC scratch is sized for the worst case of one cluster per triangle and allocated outside the timed region.
    const bool try_ploc = ctx->use_ploc && n > leaf_max;
    if (try_ploc) {
        BUILD_CHECK(d_flag.resize(n)); BUILD_CHECK(d_pos.resize(n)); BUILD_CHECK(d_scalars.resize(2)); BUILD_CHECK(d_nearest.resize(n));
        for (int k = 0; k < 2; ++k) { BUILD_CHECK(d_link[k].resize(n)); BUILD_CHECK(d_depth[k].resize(n)); BUILD_CHECK(d_box[k].resize(n)); }
        BUILD_CHECK(d_scan_temp.resize(sort::scan_scratch_words(n))); BUILD_CHECK(d_scan_total.resize(1)); BUILD_CHECK(d_ploc_state.resize(1)); BUILD_CHECK(d_block_counts.resize((size_t)ctx->sm_count * 4));
    }
    BUILD_CHECK(d_leaf_boxes.resize(std::max<size_t>(n, 1)));
    // every scratch buffer is allocated here, outside the timed region (cudaMalloc / cudaFree of gigabytes take tens of ms)
    if (n > 0) {
        BUILD_CHECK(d_keys.resize(n)); BUILD_CHECK(d_keys_alt.resize(n)); BUILD_CHECK(d_vals.resize(n)); BUILD_CHECK(d_vals_alt.resize(n));
        BUILD_CHECK(d_temp.resize(sort::sort_scratch_words(n)));
        BUILD_CHECK(d_tree.resize(std::max(n - 1, 1))); BUILD_CHECK(d_parent_internal.resize(std::max(n - 1, 1)));
        BUILD_CHECK(d_parent_leaf.resize(n)); BUILD_CHECK(d_arrival.resize(std::max(n - 1, 1))); BUILD_CHECK(d_node_boxes.resize(std::max(n - 1, 1)));
    }
    if (ctx->use_wide || try_cw) { // worst case: as many wide nodes and tasks as binary nodes; trimmed after the build
        BUILD_CHECK(d_tasks[0].resize((size_t)n + 1)); BUILD_CHECK(d_tasks[1].resize((size_t)n + 1));
        BUILD_CHECK(d_counters.resize(2));
    }
    if (try_cw && n >= 2) { BUILD_CHECK(A.cw_nodes.resize((size_t)n + 1)); BUILD_CHECK(d_triangles_by_node.resize(n)); }
    // The four-wide nodes are the fallback of the eight-wide ones: their worst-case array is only allocated up front when
    // they are the first choice, so that the usual build does not reserve 128 bytes per triangle it never touches.
    if (ctx->use_wide && !(try_cw && n >= 2)) BUILD_CHECK(A.wide_nodes.resize((size_t)n + 1));

    BUILD_CHECK(cudaEventRecord(ctx->ev[0], st));
    const int block = 256;
    auto grid = [&](int64_t count) { return (int)std::min<int64_t>((count + block - 1) / block, (int64_t)ctx->sm_count * 16); };
    auto full_grid = [&](int64_t count) { return (int)((count + block - 1) / block); };

    if (n > 0) {
        flatten_kernel<<<grid(n), block, 0, st>>>(n, (int)records.size(), d_records.ptr, A.world_vertices.ptr, A.shade.ptr,
                                                  A.has_uv ? A.shade_uv.ptr : nullptr, A.has_emission ? A.shade_emission.ptr : nullptr, d_bounds.ptr);
        ctx->counters.kernel_launches++;
        morton_kernel<<<grid(n), block, 0, st>>>(n, A.world_vertices.ptr, d_bounds.ptr, d_keys.ptr, d_vals.ptr);
        ctx->counters.kernel_launches++;

        // 63-bit Morton codes: eight 8-bit passes of the radix sort in bpt_sort.cuh
        const int sorted_in_alt = sort::radix_sort_pairs<uint64_t>(d_keys.ptr, d_vals.ptr, d_keys_alt.ptr, d_vals_alt.ptr, (uint32_t)n, nullptr, 0, 64, d_temp.ptr,
                                                                   ctx->sm_count, st, &ctx->counters.kernel_launches);
        const uint64_t* sorted_keys = sorted_in_alt ? d_keys_alt.ptr : d_keys.ptr;
        const uint32_t* sorted_vals = sorted_in_alt ? d_vals_alt.ptr : d_vals.ptr;

        BUILD_CHECK(cudaMemsetAsync(d_arrival.ptr, 0, sizeof(int) * std::max(n - 1, 1), st));
        if (n > 1) {
            hierarchy_kernel<<<full_grid(n - 1), block, 0, st>>>(n, sorted_keys, d_tree.ptr, d_parent_internal.ptr, d_parent_leaf.ptr);
            ctx->counters.kernel_launches++;
        }
        fit_kernel<<<full_grid(n), block, 0, st>>>(n, sorted_vals, A.world_vertices.ptr, A.shade.ptr, A.triangles.ptr, A.slot_of_primitive.ptr,
                                                   d_leaf_boxes.ptr, d_node_boxes.ptr, d_tree.ptr, d_parent_internal.ptr, d_parent_leaf.ptr, d_arrival.ptr);
        ctx->counters.kernel_launches++;
        BUILD_CHECK(cudaEventRecord(ctx->ev[2], st)); // flatten, Morton codes, sort, Karras hierarchy, fit
        // ---- upper hierarchy: PLOC over the leaf clusters; the plain LBVH emit is the fallback ----
        bool ploc_done = false;
        if (try_ploc) {
#define PLOC_CHECK(expr) BUILD_CHECK(expr)
            PLOC_CHECK(cudaMemsetAsync(d_flag.ptr, 0, sizeof(uint32_t) * n, st));
            ploc_mark_kernel<<<full_grid(n - 1), block, 0, st>>>(n, leaf_max, d_tree.ptr, d_flag.ptr);
            sort::exclusive_scan(d_flag.ptr, d_pos.ptr, (uint32_t)n, nullptr, d_scan_temp.ptr, d_scan_total.ptr, ctx->sm_count, st);
            uint32_t kept = 0; // number of flags set = the scan's grand total
            PLOC_CHECK(cudaMemcpyAsync(&kept, d_scan_total.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            PLOC_CHECK(cudaStreamSynchronize(st));
            int m = int(kept);
            ctx->counters.kernel_launches += 1 + sort::SCAN_LAUNCHES;
            if (m >= 2) {
                // m - 1 nodes from index 1 on, the root is copied to index 0: A.nodes holds n + 1 entries
                ploc_gather_kernel<<<full_grid(n - 1), block, 0, st>>>(n, leaf_max, d_tree.ptr, d_leaf_boxes.ptr, d_node_boxes.ptr, d_pos.ptr, d_link[0].ptr, d_box[0].ptr,
                                                                       d_depth[0].ptr);
                int h_scalars[2] = { 1, 0 }; // next node index, deepest cluster
                PLOC_CHECK(cudaMemcpyAsync(d_scalars.ptr, h_scalars, sizeof(h_scalars), cudaMemcpyHostToDevice, st));
                ctx->counters.kernel_launches++;
                // The passes follow each other on the device, inside one cooperative launch (ploc_persistent_kernel). Round 1 read
                // the cluster count back and synchronised once per pass; that loop over stream launches remains as the fallback for
                // a device without cooperative launches (and as the A/B: BPT_PLOC=host).
                // Search radius: 16 positions to either side; 32 in the large-scene regime (50 M triangles: 599 -> 610 Msamples/s for
                // 12 ms more build; 1 M triangles: no gain). BPT_PLOC_RADIUS at compile time overrides both.
#ifdef BPT_PLOC_RADIUS_FORCED
                const int ploc_radius = PLOC_RADIUS;
#else
                const int ploc_radius = prim_total > 8000000ll ? 2 * PLOC_RADIUS : PLOC_RADIUS;
#endif
                PlocLists lists = { { d_link[0].ptr, d_link[1].ptr }, { d_box[0].ptr, d_box[1].ptr }, { d_depth[0].ptr, d_depth[1].ptr }, ploc_radius };
                PlocState* state = d_ploc_state.ptr;
                ploc_begin_kernel<<<1, 1, 0, st>>>(state, (uint32_t)m);
                ctx->counters.kernel_launches++;
                static const bool host_loop = [] { const char* e = getenv("BPT_PLOC"); return e && strcmp(e, "host") == 0; }();
                int cooperative = 0, blocks_per_sm = 0;
                cudaDeviceGetAttribute(&cooperative, cudaDevAttrCooperativeLaunch, ctx->device);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, ploc_persistent_kernel, PLOC_BLOCK, 0);
                const bool can_loop_on_device = cooperative && blocks_per_sm > 0 && !host_loop && (size_t)ctx->sm_count <= d_block_counts.size;
                // Passes over millions of clusters are work bound (the neighbour search streams the boxes from DRAM) and run
                // faster as full-size stream launches (24 ms of passes at 50 M triangles against 29 ms with all of them in the
                // cooperative kernel); the many small passes after them are latency bound and go to the cooperative kernel.
                const uint32_t device_loop_below = can_loop_on_device ? (4u << 20) : (uint32_t)PLOC_TAIL;
                PlocState h_state = { (uint32_t)m, 0, 0, 0 };
                const int pass_grid = (int)std::min<int64_t>(full_grid(m), (int64_t)ctx->sm_count * 16);
                while (h_state.m > std::max(device_loop_below, (uint32_t)PLOC_TAIL) && !h_state.failed) {
                    ploc_nearest_kernel<<<pass_grid, block, 0, st>>>(state, lists, d_nearest.ptr);
                    ploc_merge_kernel<<<pass_grid, block, 0, st>>>(state, lists, d_nearest.ptr, d_flag.ptr, A.nodes.ptr, d_scalars.ptr, d_scalars.ptr + 1);
                    sort::exclusive_scan(d_flag.ptr, d_pos.ptr, (uint32_t)m, &state->m, d_scan_temp.ptr, d_scan_total.ptr, ctx->sm_count, st);
                    ploc_compact_kernel<<<pass_grid, block, 0, st>>>(state, lists, d_flag.ptr, d_pos.ptr);
                    ploc_advance_kernel<<<1, 1, 0, st>>>(state, d_scan_total.ptr);
                    PLOC_CHECK(cudaMemcpyAsync(&h_state, state, sizeof(h_state), cudaMemcpyDeviceToHost, st));
                    PLOC_CHECK(cudaStreamSynchronize(st));
                    ctx->counters.kernel_launches += 4 + sort::SCAN_LAUNCHES;
                    if (getenv("BPT_PLOC_DEBUG")) fprintf(stderr, "ploc pass %d: %u clusters\n", h_state.passes, h_state.m);
                }
                bool looped_on_device = false;
                if (can_loop_on_device && h_state.m > (uint32_t)PLOC_TAIL && !h_state.failed) {
                    const int grid_blocks = (int)std::min<int64_t>(ctx->sm_count, (h_state.m + PLOC_BLOCK - 1) / PLOC_BLOCK); // one block per SM: a grid-wide barrier costs by the block
                    int* nearest = d_nearest.ptr; uint32_t* keep = d_flag.ptr; uint32_t* block_counts = d_block_counts.ptr;
                    BvhNode* nodes = A.nodes.ptr; int* node_counter = d_scalars.ptr; int* max_depth = d_scalars.ptr + 1;
                    void* args[] = { &state, &lists, &nearest, &keep, &block_counts, &nodes, &node_counter, &max_depth };
                    PLOC_CHECK(cudaLaunchCooperativeKernel((const void*)ploc_persistent_kernel, dim3(grid_blocks), dim3(PLOC_BLOCK), args, 0, st));
                    looped_on_device = true; ctx->counters.kernel_launches++;
                }
                ploc_tail_kernel<<<1, PLOC_TAIL, 0, st>>>(state, lists, A.nodes.ptr, d_scalars.ptr, d_scalars.ptr + 1);
                ctx->counters.kernel_launches++;
                PLOC_CHECK(cudaMemcpyAsync(&h_state, state, sizeof(h_state), cudaMemcpyDeviceToHost, st));
                PLOC_CHECK(cudaMemcpyAsync(h_scalars, d_scalars.ptr, sizeof(h_scalars), cudaMemcpyDeviceToHost, st));
                PLOC_CHECK(cudaStreamSynchronize(st));
                if (!h_state.failed && h_state.m <= (uint32_t)PLOC_TAIL && h_scalars[1] <= PLOC_MAX_DEPTH) {
                    A.node_count = h_scalars[0];
                    A.ploc_passes = h_state.passes; A.ploc_depth = h_scalars[1];
                    A.ploc_on_device = looped_on_device;
                    ploc_done = true;
                }
            }
#undef PLOC_CHECK
        }
        if (!ploc_done && n > 1) {
            emit_kernel<<<full_grid(n - 1), block, 0, st>>>(n, leaf_max, d_tree.ptr, d_leaf_boxes.ptr, d_node_boxes.ptr, A.nodes.ptr);
            ctx->counters.kernel_launches++;
            A.node_count = n - 1; A.ploc_passes = 0; A.ploc_depth = 0;
        }
    }
    if (n <= 1) {
        tiny_root_kernel<<<1, 1, 0, st>>>(n, d_leaf_boxes.ptr, A.nodes.ptr);
        ctx->counters.kernel_launches++;
        A.node_count = 1; A.ploc_passes = 0; A.ploc_depth = 0;
    }
    if (n > 0) BUILD_CHECK(cudaEventRecord(ctx->ev[3], st)); // the binary hierarchy
    // ---- compressed eight-wide collapse of whichever binary hierarchy was built ----
    A.cw_levels = 0; A.cw_node_count = 0;
    if (try_cw && n >= 2) {
        WideTask root = { 0, 0 };
        BUILD_CHECK(cudaMemcpyAsync(d_tasks[0].ptr, &root, sizeof(root), cudaMemcpyHostToDevice, st));
        int h_counters[2] = { 1, 0 }; // nodes allocated (the root is node 0), triangles placed
        BUILD_CHECK(cudaMemcpyAsync(d_counters.ptr, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st));
        int count = 1, cur = 0, levels = 0, next_level_base = 1;
        while (count > 0) {
            cw_collapse_kernel<<<full_grid(count), block, 0, st>>>(count, next_level_base, d_tasks[cur].ptr, d_tasks[cur ^ 1].ptr, d_counters.ptr, A.nodes.ptr,
                                                                  A.cw_nodes.ptr, A.triangles.ptr, d_triangles_by_node.ptr);
            ctx->counters.kernel_launches++;
            BUILD_CHECK(cudaMemcpyAsync(h_counters, d_counters.ptr, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
            BUILD_CHECK(cudaStreamSynchronize(st));
            count = h_counters[0] - next_level_base; next_level_base = h_counters[0]; cur ^= 1; ++levels;
        }
        // a node visit leaves at most two groups on the stack: fall back to the four-wide nodes if that could overflow it
        if (h_counters[1] == n && 2 * levels + 2 <= CW_STACK_SMEM + CW_STACK_LOCAL) {
            A.cw_levels = levels; A.cw_node_count = h_counters[0];
            std::swap(A.triangles.ptr, d_triangles_by_node.ptr); std::swap(A.triangles.capacity, d_triangles_by_node.capacity); std::swap(A.triangles.size, d_triangles_by_node.size);
        }
    }
    // ---- four-wide collapse: the fallback of the eight-wide nodes, or the first choice with BPT_CW=0 ----
    A.wide_levels = 0; A.wide_node_count = 0;
    if (ctx->use_wide && A.cw_levels == 0) {
#define WIDE_CHECK(expr) BUILD_CHECK(expr)
        WIDE_CHECK(A.wide_nodes.resize((size_t)n + 1));
        WideTask root = { 0, 0 };
        WIDE_CHECK(cudaMemcpyAsync(d_tasks[0].ptr, &root, sizeof(root), cudaMemcpyHostToDevice, st));
        int h_counters[2] = { 0, 1 }; // tasks of the next level, wide nodes allocated (the root is node 0)
        int count = 1, cur = 0, levels = 0;
        while (count > 0) {
            h_counters[0] = 0;
            WIDE_CHECK(cudaMemcpyAsync(d_counters.ptr, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st));
            collapse_kernel<<<full_grid(count), block, 0, st>>>(count, d_tasks[cur].ptr, d_tasks[cur ^ 1].ptr, d_counters.ptr, A.nodes.ptr, A.wide_nodes.ptr);
            ctx->counters.kernel_launches++;
            WIDE_CHECK(cudaMemcpyAsync(h_counters, d_counters.ptr, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
            WIDE_CHECK(cudaStreamSynchronize(st));
            count = h_counters[0]; cur ^= 1; ++levels;
        }
#undef WIDE_CHECK
        // a ray pushes at most three links per level: fall back to the binary nodes if that could overflow the stack
        if (3 * levels + 1 <= STACK_SMEM + STACK_LOCAL) { A.wide_levels = levels; A.wide_node_count = h_counters[1]; }
    }
    BUILD_CHECK(cudaEventRecord(ctx->ev[1], st));
    BUILD_CHECK(cudaGetLastError());
    BUILD_CHECK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&A.build_ms, ctx->ev[0], ctx->ev[1]);
    if (n > 0 && getenv("BPT_BUILD_DEBUG")) {
        float leaves_ms = 0.0f, upper_ms = 0.0f, collapse_ms = 0.0f;
        cudaEventElapsedTime(&leaves_ms, ctx->ev[0], ctx->ev[2]); cudaEventElapsedTime(&upper_ms, ctx->ev[2], ctx->ev[3]); cudaEventElapsedTime(&collapse_ms, ctx->ev[3], ctx->ev[1]);
        fprintf(stderr, "bpt_build_accel: %d triangles, %.3f ms = %.3f (flatten, sort, leaf clusters) + %.3f (%d PLOC passes%s) + %.3f (collapse, %d levels)\n",
                n, A.build_ms, leaves_ms, upper_ms, A.ploc_passes, A.ploc_on_device ? ", those below 4 M clusters in one cooperative launch" : ", a host read per pass", collapse_ms, A.cw_levels > 0 ? A.cw_levels : A.wide_levels);
    }
    release_all();
    // Give back the worst-case slack of the node arrays when it is large (outside the timed region).
    auto trim = [&](auto& buffer, size_t used) -> cudaError_t {
        typedef typename std::remove_reference<decltype(*buffer.ptr)>::type T;
        if (buffer.capacity < used + (64u << 20) / sizeof(T)) return cudaSuccess;
        T* exact = nullptr;
        cudaError_t e = cudaMalloc((void**)&exact, std::max<size_t>(used, 1) * sizeof(T));
        if (e != cudaSuccess) return cudaSuccess; // keep the larger allocation
        e = cudaMemcpy(exact, buffer.ptr, used * sizeof(T), cudaMemcpyDeviceToDevice);
        cudaFree(buffer.ptr);
        buffer.ptr = exact; buffer.capacity = buffer.size = std::max<size_t>(used, 1);
        return e;
    };
    BUILD_CHECK(trim(A.nodes, (size_t)A.node_count + 1));
    if (A.wide_levels > 0) BUILD_CHECK(trim(A.wide_nodes, (size_t)A.wide_node_count));
    else A.wide_nodes.release();
    if (A.cw_levels > 0) BUILD_CHECK(trim(A.cw_nodes, (size_t)A.cw_node_count));
    else A.cw_nodes.release();
#undef BUILD_CHECK

    A.triangle_count = n;
    A.valid = true;
    return BPT_OK;
}

int intersect_batch(Context* ctx, int64_t n, const float* origins, const float* directions, const float* tmin, const float* tmax,
                    int32_t* out_primitive, float* out_t, float* out_uv, uint8_t* out_occluded) {
    if (!ctx->accel.valid) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_intersect: call bpt_build_accel first");
    if (n < 0 || n > 0x7fffffff || !origins || !directions || !tmin || !tmax) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_intersect: bad arguments");
    if (n == 0) return BPT_OK;
    if (int status = sync_texture_table(ctx)) return status;
    cudaStream_t st = ctx->stream;
    std::vector<float> h_cov(ctx->host_materials.size());
    for (size_t i = 0; i < h_cov.size(); ++i) h_cov[i] = material_coverage_table_entry(ctx->host_materials[i]);

    // One context-owned scratch block that only grows (no allocation on a repeated query), carved into 256-byte aligned parts.
    size_t offset = 0;
    auto carve = [&](size_t bytes) { size_t at = offset; offset = (offset + bytes + 255) & ~size_t(255); return at; };
    const size_t at_o = carve(3 * n * sizeof(float)), at_d = carve(3 * n * sizeof(float)), at_tmin = carve(n * sizeof(float)), at_tmax = carve(n * sizeof(float));
    const size_t at_cov = carve(std::max<size_t>(h_cov.size(), 1) * sizeof(float));
    const size_t at_prim = carve(n * sizeof(int32_t)), at_t = carve(n * sizeof(float)), at_uv = carve(2 * n * sizeof(float)), at_occ = carve(n);
#define Q_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return ctx->cuda_fail(_e, #expr); } while (0)
    if (ctx->query_scratch.capacity < offset) Q_CHECK(cudaStreamSynchronize(st));
    Q_CHECK(ctx->query_scratch.resize(offset));
    unsigned char* base = ctx->query_scratch.ptr;
    float *d_o = (float*)(base + at_o), *d_d = (float*)(base + at_d), *d_tmin = (float*)(base + at_tmin), *d_tmax = (float*)(base + at_tmax), *d_cov = (float*)(base + at_cov);
    int32_t* d_prim = out_primitive ? (int32_t*)(base + at_prim) : nullptr;
    float *d_t = out_t ? (float*)(base + at_t) : nullptr, *d_uv = out_uv ? (float*)(base + at_uv) : nullptr;
    uint8_t* d_occ = out_occluded ? base + at_occ : nullptr;
    Q_CHECK(cudaMemcpyAsync(d_o, origins, 3 * n * sizeof(float), cudaMemcpyHostToDevice, st));
    Q_CHECK(cudaMemcpyAsync(d_d, directions, 3 * n * sizeof(float), cudaMemcpyHostToDevice, st));
    Q_CHECK(cudaMemcpyAsync(d_tmin, tmin, n * sizeof(float), cudaMemcpyHostToDevice, st));
    Q_CHECK(cudaMemcpyAsync(d_tmax, tmax, n * sizeof(float), cudaMemcpyHostToDevice, st));
    if (!h_cov.empty()) Q_CHECK(cudaMemcpyAsync(d_cov, h_cov.data(), h_cov.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    AccelView view = accel_view(ctx);
    int grid = (int)std::min<int64_t>((n + TRACE_BLOCK - 1) / TRACE_BLOCK, (int64_t)ctx->sm_count * 8);
    BatchSource source = { d_o, d_d, d_tmin, d_tmax, d_prim, d_t, d_uv, d_occ };
    unsigned int* d_fetch = reinterpret_cast<unsigned int*>(ctx->device_counters + 4); // two scratch fetch counters
    Q_CHECK(cudaMemsetAsync(d_fetch, 0, 2 * sizeof(unsigned int), st));
    if (out_primitive || out_t || out_uv) {
        if (view.cw) intersect_kernel<false, true><<<grid, TRACE_BLOCK, 0, st>>>(view, (unsigned int)n, source, d_cov, d_fetch);
        else intersect_kernel<false, false><<<grid, TRACE_BLOCK, 0, st>>>(view, (unsigned int)n, source, d_cov, d_fetch);
        ctx->counters.kernel_launches++;
    }
    if (out_occluded) {
        if (view.cw) intersect_kernel<true, true><<<grid, TRACE_BLOCK, 0, st>>>(view, (unsigned int)n, source, d_cov, d_fetch + 1);
        else intersect_kernel<true, false><<<grid, TRACE_BLOCK, 0, st>>>(view, (unsigned int)n, source, d_cov, d_fetch + 1);
        ctx->counters.kernel_launches++;
    }
    Q_CHECK(cudaGetLastError());
    if (out_primitive) Q_CHECK(cudaMemcpyAsync(out_primitive, d_prim, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (out_t) Q_CHECK(cudaMemcpyAsync(out_t, d_t, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out_uv) Q_CHECK(cudaMemcpyAsync(out_uv, d_uv, 2 * n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out_occluded) Q_CHECK(cudaMemcpyAsync(out_occluded, d_occ, n, cudaMemcpyDeviceToHost, st));
    Q_CHECK(cudaStreamSynchronize(st));
#undef Q_CHECK
    return BPT_OK;
}

} // namespace bpt
