// Compressed eight-wide nodes: the node format, its encoder and the ray / node test, after Ylitie, Karras and Laine,
// "Efficient incoherent ray traversal on GPUs through compressed wide BVHs" (HPG 2017). One more layer under the path
// that replaces OptiX' closed-source hierarchy (Renderer.cpp:116-135,161-182,470-477): the four-wide nodes of
// bpt_context.h spend seven 128-bit loads on four children, this format spends five on eight.
//
// * A node is 80 bytes: an origin p, one power-of-two scale per axis and the eight child boxes as 8-bit plane indices,
//   plane = p + q * 2^e. The encoder rounds outwards, so a quantised box always contains the child's real box.
// * The children of a node that are nodes themselves are stored contiguously from `child_base`, in slot order; the
//   triangles of its leaf children are stored contiguously from `triangle_base` (bpt_bvh.cu reorders the triangle array).
//   One meta byte per slot: 001sssss with sssss = 24 + slot for an inner child, ccc ooooo (unary triangle count, offset
//   from triangle_base) for a leaf, 0 for an empty slot.
// * Slots are octants: the child that lies furthest in direction (+-1, +-1, +-1) goes to the slot with those sign bits, so
//   `slot ^ ray octant` is a front-to-back order without any sorting at traversal time.
// * The ray / node test returns one word: bits 24..31 = inner children that were hit, in traversal priority (highest
//   first), bits 0..23 = triangles of the leaf children that were hit.
//
// Everything here is __host__ __device__: tests/host/cw_host_test.cpp runs the encoder and the test on the CPU against
// exact boxes. The warp-level traversal loop that drives them is in bpt_trace.cuh.
#pragma once
#include <string.h>
#include "bpt_math.cuh"

namespace bpt {

struct __align__(16) CwNode {
    float px, py, pz;
    uint32_t e_imask;                 // exponent bytes of the x, y, z scales | imask << 24 (slots that hold inner children)
    uint32_t child_base, triangle_base;
    uint32_t meta[2];                 // one byte per slot
    uint32_t qlo_x[2], qlo_y[2];      // one byte per slot and plane
    uint32_t qlo_z[2], qhi_x[2];
    uint32_t qhi_y[2], qhi_z[2];
};
static_assert(sizeof(CwNode) == 80, "CwNode must be 80 bytes");

constexpr int CW_WIDTH = 8;
constexpr int CW_MAX_LEAF_TRIANGLES = 3;

BPT_HD float cw_as_float(uint32_t bits) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(bits);
#else
    float f; memcpy(&f, &bits, 4); return f;
#endif
}
BPT_HD uint32_t cw_as_uint(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t bits; memcpy(&bits, &f, 4); return bits;
#endif
}
BPT_HD float cw_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
BPT_HD int cw_highest_bit(uint32_t v) { // v != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
BPT_HD int cw_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// PTX prmt.b32, default mode: result byte i = byte (selector nibble i & 7) of {b, a}; nibble bit 3 replicates that byte's sign.
// (The selector is a template argument so that it becomes an immediate of the instruction.)
template <uint32_t selector>
BPT_HD uint32_t cw_prmt(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "n"(selector));
    return r;
#else
    const uint64_t both = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t s = (selector >> (4 * i)) & 0xfu;
        uint32_t byte = (uint32_t)(both >> (8 * (s & 7u))) & 0xffu;
        if (s & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
        r |= byte << (8 * i);
    }
    return r;
#endif
}
// a - b rounded down / up.
BPT_HD float cw_sub_down(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rd(a, b);
#else
    const double d = (double)a - (double)b; float f = (float)d;
    return (double)f > d ? nextafterf(f, -INFINITY) : f;
#endif
}
BPT_HD float cw_sub_up(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_ru(a, b);
#else
    const double d = (double)a - (double)b; float f = (float)d;
    return (double)f < d ? nextafterf(f, INFINITY) : f;
#endif
}

// ---- encoder --------------------------------------------------------------------------------------------------------

// Smallest exponent byte b (scale 2^(b - 127)) with 255 * scale >= extent.
BPT_HD uint32_t cw_scale_exponent(float extent) {
    int ex = -126;
    if (extent > 0.0f) {
        int k; frexpf(extent, &k); // extent = m * 2^k, m in [0.5, 1)
        ex = k - 8;                // 255 * 2^(k - 8) >= extent unless m > 255 / 256
        if (ldexpf(255.0f, ex) < extent) ++ex;
    }
    ex = ex < -126 ? -126 : (ex > 100 ? 100 : ex);
    return (uint32_t)(ex + 127);
}

// What the encoder decides about one child besides its box.
struct CwPlacement {
    int slot[CW_WIDTH];     // slot of child i
    int offset[CW_WIDTH];   // inner child: its rank among the node's inner children (node index = child_base + rank);
                            // leaf child: offset of its first triangle from triangle_base
    int inner_count, triangle_count;
};

// Encodes a node from `count` (2..8) children. triangles[i] = 0 for an inner child, 1..3 for a leaf with that many
// triangles. Fills everything except child_base / triangle_base, which the caller allocates from `placement`'s totals.
BPT_HD void cw_encode(int count, const float3* lo, const float3* hi, const int* triangles, CwNode& node, CwPlacement& placement) {
    float3 nlo = lo[0], nhi = hi[0];
    for (int i = 1; i < count; ++i) { nlo = min3(nlo, lo[i]); nhi = max3(nhi, hi[i]); }

    // Octant slots: greedily hand the (child, slot) pair with the largest projection of the child's centre (relative to the
    // node's) on the slot's diagonal its slot. Bit 4 / 2 / 1 of a slot = the +x / +y / +z side. (All loops run over the full
    // width with masks, so that they unroll and the 8 x 8 projections stay in registers: with loops over `count` the build
    // kernel kept them in local memory and spent 43 us per tree level on this function.)
    float cost[CW_WIDTH][CW_WIDTH];
#pragma unroll
    for (int i = 0; i < CW_WIDTH; ++i) {
        const int c = i < count ? i : 0;
        const float cx = (lo[c].x + hi[c].x) - (nlo.x + nhi.x), cy = (lo[c].y + hi[c].y) - (nlo.y + nhi.y), cz = (lo[c].z + hi[c].z) - (nlo.z + nhi.z);
#pragma unroll
        for (int s = 0; s < CW_WIDTH; ++s)
            cost[i][s] = ((s & 4) ? cx : -cx) + ((s & 2) ? cy : -cy) + ((s & 1) ? cz : -cz);
    }
    uint32_t child_free = (1u << count) - 1u, slot_free = 0xffu;
    int child_in_slot[CW_WIDTH];
#pragma unroll
    for (int s = 0; s < CW_WIDTH; ++s) child_in_slot[s] = -1;
#pragma unroll
    for (int round = 0; round < CW_WIDTH; ++round) {
        if (round < count) {
            int best_pair = -1; float best = 0.0f; // pair = child * 8 + slot
#pragma unroll
            for (int i = 0; i < CW_WIDTH; ++i)
#pragma unroll
                for (int s = 0; s < CW_WIDTH; ++s) {
                    const bool open = (child_free >> i & 1u) && (slot_free >> s & 1u);
                    if (open && (best_pair < 0 || cost[i][s] > best)) { best = cost[i][s]; best_pair = i * CW_WIDTH + s; }
                }
            const int best_child = best_pair >> 3, best_slot = best_pair & 7;
            child_free &= ~(1u << best_child); slot_free &= ~(1u << best_slot);
#pragma unroll
            for (int s = 0; s < CW_WIDTH; ++s) if (s == best_slot) child_in_slot[s] = best_child;
#pragma unroll
            for (int i = 0; i < CW_WIDTH; ++i) if (i == best_child) placement.slot[i] = best_slot;
        }
    }

    const uint32_t bx = cw_scale_exponent(cw_sub_up(nhi.x, nlo.x)), by = cw_scale_exponent(cw_sub_up(nhi.y, nlo.y)), bz = cw_scale_exponent(cw_sub_up(nhi.z, nlo.z));
    const float rx = cw_as_float((254u - bx) << 23), ry = cw_as_float((254u - by) << 23), rz = cw_as_float((254u - bz) << 23); // 1 / scale, exact

    uint8_t meta[CW_WIDTH], q[6][CW_WIDTH];
    uint32_t imask = 0;
    int inner = 0, tris = 0;
    for (int s = 0; s < CW_WIDTH; ++s) {
        const int i = child_in_slot[s];
        if (i < 0) { // empty slot: an inverted box that no ray passes, and no bits to set if one did
            meta[s] = 0;
            q[0][s] = q[1][s] = q[2][s] = 255; q[3][s] = q[4][s] = q[5][s] = 0;
            continue;
        }
        if (triangles[i] == 0) {
            meta[s] = (uint8_t)(0x20u | (24u + (uint32_t)s));
            imask |= 1u << s;
            placement.offset[i] = inner++;
        } else {
            meta[s] = (uint8_t)((((1u << triangles[i]) - 1u) << 5) | (uint32_t)tris);
            placement.offset[i] = tris;
            tris += triangles[i];
        }
        // lower planes round down, upper planes round up (the differences too), so the quantised box contains the real one
        const float l[3] = { floorf(cw_sub_down(lo[i].x, nlo.x) * rx), floorf(cw_sub_down(lo[i].y, nlo.y) * ry), floorf(cw_sub_down(lo[i].z, nlo.z) * rz) };
        const float h[3] = { ceilf(cw_sub_up(hi[i].x, nlo.x) * rx), ceilf(cw_sub_up(hi[i].y, nlo.y) * ry), ceilf(cw_sub_up(hi[i].z, nlo.z) * rz) };
        for (int a = 0; a < 3; ++a) {
            q[a][s] = (uint8_t)fminf(fmaxf(l[a], 0.0f), 255.0f);
            q[3 + a][s] = (uint8_t)fminf(fmaxf(h[a], 0.0f), 255.0f);
        }
    }
    placement.inner_count = inner; placement.triangle_count = tris;

    auto pack = [](const uint8_t* b) { return (uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24; };
    node.px = nlo.x; node.py = nlo.y; node.pz = nlo.z;
    node.e_imask = bx | by << 8 | bz << 16 | imask << 24;
    node.child_base = 0; node.triangle_base = 0;
    node.meta[0] = pack(meta); node.meta[1] = pack(meta + 4);
    node.qlo_x[0] = pack(q[0]); node.qlo_x[1] = pack(q[0] + 4); node.qlo_y[0] = pack(q[1]); node.qlo_y[1] = pack(q[1] + 4);
    node.qlo_z[0] = pack(q[2]); node.qlo_z[1] = pack(q[2] + 4); node.qhi_x[0] = pack(q[3]); node.qhi_x[1] = pack(q[3] + 4);
    node.qhi_y[0] = pack(q[4]); node.qhi_y[1] = pack(q[4] + 4); node.qhi_z[0] = pack(q[5]); node.qhi_z[1] = pack(q[5] + 4);
}

// ---- ray / node test ------------------------------------------------------------------------------------------------

// Per-ray constants of the node test.
struct CwRay {
    float3 origin;
    float3 inv_d;        // 1 / direction, with |direction| clamped away from zero
    uint32_t oct_inv4;   // (7 - octant) in every byte; octant bit 4 / 2 / 1 = the x / y / z direction is negative
};

BPT_HD float cw_reciprocal_direction(float d) {
    // A zero component would make every plane distance inf or NaN; 2^-60 keeps them finite and moves the ray by less than
    // the padding below over any distance a scene can have.
    const float tiny = 8.673617379884035e-19f;
    if (fabsf(d) < tiny) d = copysignf(tiny, d);
#ifdef __CUDA_ARCH__
    return __fdiv_rn(1.0f, d);
#else
    return 1.0f / d;
#endif
}

BPT_HD CwRay cw_make_ray(float3 origin, float3 direction) {
    CwRay r;
    r.origin = origin;
    r.inv_d = f3(cw_reciprocal_direction(direction.x), cw_reciprocal_direction(direction.y), cw_reciprocal_direction(direction.z));
    const uint32_t octant = (r.inv_d.x < 0.0f ? 4u : 0u) | (r.inv_d.y < 0.0f ? 2u : 0u) | (r.inv_d.z < 0.0f ? 1u : 0u);
    r.oct_inv4 = (7u - octant) * 0x01010101u;
    return r;
}

// The byte `index` of `word` as the float 32768 + byte: one PRMT puts it into mantissa bits 8..15 under the exponent of
// 2^15, with no integer-to-float conversion.
template <int INDEX>
BPT_HD float cw_plane_float(uint32_t word, uint32_t exponent_word) { return cw_as_float(cw_prmt<0x7504u | (uint32_t)(INDEX << 4)>(word, exponent_word)); }
// `exponent_word` is CW_EXPONENT_WORD, but handed in as a run-time value (a kernel parameter, AccelView::cw_exponent_word): PRMT takes
// one immediate, and when both the selector and this word are known at compile time ptxas keeps the word as the immediate
// and moves the selector into a register in front of every PRMT - one more instruction per plane.
constexpr uint32_t CW_EXPONENT_WORD = 0x47000000u;

// Tests the eight child boxes of a node. n0..n4 are the node's five 128-bit words. Returns hit bits as described above.
//
// Distance to the plane with index q along one axis: t = q * adj + org with adj = scale * inv_d (exact, a power of two
// times inv_d) and org = (p - origin) * inv_d. It is evaluated as fma(32768 + q, adj, base), base = org - 32768 * adj.
// Rounding: org 2^-23 |org| (p - origin, then the product); base and base -+ pad 2^-24 (|org| + 2^15 |adj|) each; the fma
// 2^-24 (|org| + 255 |adj|): together below 2^-21.6 |org| + 2^-7.9 |adj|. Entry distances are lowered and exit distances
// raised by pad = 2^-19 |org| + 2^-6 |adj|, i.e. the box grows by 2e-6 of its distance from the ray origin plus 1 / 64 of
// a quantisation step: the test errs on the side of visiting, never of culling, and keeps more than the 8 ulp of slack
// against the rounding of the triangle test's t that the uncompressed slab test has (bpt_trace.cuh: slab).
BPT_HD uint32_t cw_intersect_children(const uint4& n0, const uint4& n1, const uint4& n2, const uint4& n3, const uint4& n4, const CwRay& ray, float tmin, float tmax,
                                       uint32_t exponent_word = CW_EXPONENT_WORD) {
    const uint32_t e = n0.w;
    const float adj_x = cw_as_float((e & 0xffu) << 23) * ray.inv_d.x, adj_y = cw_as_float((e >> 8 & 0xffu) << 23) * ray.inv_d.y, adj_z = cw_as_float((e >> 16 & 0xffu) << 23) * ray.inv_d.z;
    const float org_x = (cw_as_float(n0.x) - ray.origin.x) * ray.inv_d.x, org_y = (cw_as_float(n0.y) - ray.origin.y) * ray.inv_d.y, org_z = (cw_as_float(n0.z) - ray.origin.z) * ray.inv_d.z;
    const float pad_x = cw_fma(fabsf(org_x), 1.9073486328125e-6f, fabsf(adj_x) * 0.015625f);
    const float pad_y = cw_fma(fabsf(org_y), 1.9073486328125e-6f, fabsf(adj_y) * 0.015625f);
    const float pad_z = cw_fma(fabsf(org_z), 1.9073486328125e-6f, fabsf(adj_z) * 0.015625f);
    const float base_x = cw_fma(-32768.0f, adj_x, org_x), base_y = cw_fma(-32768.0f, adj_y, org_y), base_z = cw_fma(-32768.0f, adj_z, org_z);
    const float near_x = base_x - pad_x, near_y = base_y - pad_y, near_z = base_z - pad_z;
    const float far_x = base_x + pad_x, far_y = base_y + pad_y, far_z = base_z + pad_z;
    // the ray enters through the lower planes of the axes it travels up and through the upper planes of the others
    const bool up_x = ray.inv_d.x >= 0.0f, up_y = ray.inv_d.y >= 0.0f, up_z = ray.inv_d.z >= 0.0f;

    const uint32_t ew = exponent_word;
    uint32_t hits = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t meta4 = half ? n1.w : n1.z;
        const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;      // low five bits >= 24
        const uint32_t inner_mask4 = cw_prmt<0xba98u>(is_inner4 << 3, 0u);     // 0xff in the bytes of inner children
        const uint32_t bit_index4 = (meta4 ^ (ray.oct_inv4 & inner_mask4)) & 0x1f1f1f1fu;
        const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
        const uint32_t lo_x = half ? n2.y : n2.x, lo_y = half ? n2.w : n2.z, lo_z = half ? n3.y : n3.x;
        const uint32_t hi_x = half ? n3.w : n3.z, hi_y = half ? n4.y : n4.x, hi_z = half ? n4.w : n4.z;
        const uint32_t in_x = up_x ? lo_x : hi_x, in_y = up_y ? lo_y : hi_y, in_z = up_z ? lo_z : hi_z;
        const uint32_t out_x = up_x ? hi_x : lo_x, out_y = up_y ? hi_y : lo_y, out_z = up_z ? hi_z : lo_z;
#define BPT_CW_CHILD(J) { \
            const float t_in = fmaxf(fmaxf(cw_fma(cw_plane_float<J>(in_x, ew), adj_x, near_x), cw_fma(cw_plane_float<J>(in_y, ew), adj_y, near_y)), \
                                     fmaxf(cw_fma(cw_plane_float<J>(in_z, ew), adj_z, near_z), tmin)); \
            const float t_out = fminf(fminf(cw_fma(cw_plane_float<J>(out_x, ew), adj_x, far_x), cw_fma(cw_plane_float<J>(out_y, ew), adj_y, far_y)), \
                                      fminf(cw_fma(cw_plane_float<J>(out_z, ew), adj_z, far_z), tmax)); \
            if (t_in <= t_out) hits |= ((child_bits4 >> (8 * J)) & 0xffu) << ((bit_index4 >> (8 * J)) & 0xffu); }
        BPT_CW_CHILD(0) BPT_CW_CHILD(1) BPT_CW_CHILD(2) BPT_CW_CHILD(3)
#undef BPT_CW_CHILD
    }
    return hits;
}

// A group on the traversal stack: x = base index, y = bits. A node group has bits in 24..31 (inner children that were
// hit and not visited yet, by priority) and the parent's imask in 0..7; a triangle group has bits in 0..23 only.
BPT_HD bool cw_is_node_group(const uint2& g) { return (g.y & 0xff000000u) != 0u; }

// Takes the nearest unvisited child out of a node group: returns its node index and leaves the rest in `group`.
BPT_HD uint32_t cw_next_child(uint2& group, const CwRay& ray) {
    const uint32_t bits = group.y;
    const int bit = cw_highest_bit(bits);
    group.y = bits & ~(1u << bit);
    const uint32_t slot = (uint32_t)(bit - 24) ^ (ray.oct_inv4 & 7u);
    return group.x + (uint32_t)cw_popc(bits & ~(0xffffffffu << slot)); // inner children in the slots below this one
}

} // namespace bpt
