// Ray traversal of the BVH (four-wide nodes, binary fallback): closest hit and any hit. Replaces OptiX' proprietary Trbvh/RTX traversal that the
// reference configures at Renderer.cpp:116-135,161-182,470-477 and queries with rtTrace (SimpleRGPs.cu:114-125).
//
// * Nodes are four wide (128 bytes, child boxes stored per axis: one visit = seven 128-bit loads, four slab tests and a
//   five-comparator sort of the hit children); the 64-byte binary nodes they are collapsed from remain as the fallback for
//   hierarchies too deep for the stack.
// * Triangles are 48 bytes (three float4 world-space vertices) in Morton order: three 128-bit loads.
// * The ray/triangle test is the watertight test of Woop, Benthin and Wald (JCGT 2013): shear to ray space, 2D edge
//   functions with an fp64 fallback when an edge function is exactly zero. It is evaluated with explicitly rounded
//   operations (no FMA contraction) so the CPU oracle reproduces t and the barycentrics bit for bit; shared edges
//   evaluate to exact negations of each other, which is what makes it watertight.
// * Closest hit is the minimum over (t, global primitive index): ties in t resolve to the lower index so that the
//   result does not depend on traversal order.
// * Control flow is "while-while" (Aila and Laine, HPG 2009): a warp first runs inner-node steps until every lane that is
//   still searching has reached a leaf, then intersects leaves together; leaves travel through the stack like nodes.
// * Persistent threads with dynamic ray fetch: a lane whose ray has terminated takes the next ray of the queue (one
//   atomicAdd per warp for all idle lanes) every TRAVERSAL_BUDGET node visits instead of idling until the slowest ray of
//   its warp finishes.
// * The traversal stack lives in shared memory (STACK_SMEM entries per thread, column layout so that a warp's accesses
//   hit 32 different banks); deeper paths spill to a per-thread local array.
#pragma once
#include <cfloat>
#include "bpt_context.h"
#include "bpt_math.cuh"
#include "bpt_cw.cuh"

namespace bpt {

struct Ray {
    float3 origin;
    float tmin;
    float3 direction;
    float tmax;
};

struct Hit {
    float t;
    int primitive; // global primitive index, -1 = miss
    float u, v;    // barycentric weights of vertex 1 and vertex 2
};

// Per-ray constants of the watertight test.
struct RayShear {
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

BPT_D float comp(float3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

BPT_D RayShear make_ray_shear(float3 d) {
    RayShear s;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    s.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    s.kx = s.kz + 1; if (s.kx == 3) s.kx = 0;
    s.ky = s.kx + 1; if (s.ky == 3) s.ky = 0;
    float dz = comp(d, s.kz);
    if (dz < 0.0f) { int tmp = s.kx; s.kx = s.ky; s.ky = tmp; }
    s.Sx = __fdiv_rn(comp(d, s.kx), dz);
    s.Sy = __fdiv_rn(comp(d, s.ky), dz);
    s.Sz = __fdiv_rn(1.0f, dz);
    return s;
}

// Returns true and fills t, u, v when the ray's supporting line hits the triangle (no interval test).
BPT_D bool watertight_triangle(const RayShear& s, float3 origin, float3 p0, float3 p1, float3 p2, float& t, float& u, float& v) {
    const float3 A = f3(__fsub_rn(p0.x, origin.x), __fsub_rn(p0.y, origin.y), __fsub_rn(p0.z, origin.z));
    const float3 B = f3(__fsub_rn(p1.x, origin.x), __fsub_rn(p1.y, origin.y), __fsub_rn(p1.z, origin.z));
    const float3 C = f3(__fsub_rn(p2.x, origin.x), __fsub_rn(p2.y, origin.y), __fsub_rn(p2.z, origin.z));

    const float Akz = comp(A, s.kz), Bkz = comp(B, s.kz), Ckz = comp(C, s.kz);
    const float Ax = __fsub_rn(comp(A, s.kx), __fmul_rn(s.Sx, Akz));
    const float Ay = __fsub_rn(comp(A, s.ky), __fmul_rn(s.Sy, Akz));
    const float Bx = __fsub_rn(comp(B, s.kx), __fmul_rn(s.Sx, Bkz));
    const float By = __fsub_rn(comp(B, s.ky), __fmul_rn(s.Sy, Bkz));
    const float Cx = __fsub_rn(comp(C, s.kx), __fmul_rn(s.Sx, Ckz));
    const float Cy = __fsub_rn(comp(C, s.ky), __fmul_rn(s.Sy, Ckz));

    float U = __fsub_rn(__fmul_rn(Cx, By), __fmul_rn(Cy, Bx));
    float V = __fsub_rn(__fmul_rn(Ax, Cy), __fmul_rn(Ay, Cx));
    float W = __fsub_rn(__fmul_rn(Bx, Ay), __fmul_rn(By, Ax));

    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = __dmul_rn((double)Cx, (double)By), CyBx = __dmul_rn((double)Cy, (double)Bx);
        U = (float)__dsub_rn(CxBy, CyBx);
        double AxCy = __dmul_rn((double)Ax, (double)Cy), AyCx = __dmul_rn((double)Ay, (double)Cx);
        V = (float)__dsub_rn(AxCy, AyCx);
        double BxAy = __dmul_rn((double)Bx, (double)Ay), ByAx = __dmul_rn((double)By, (double)Ax);
        W = (float)__dsub_rn(BxAy, ByAx);
    }

    // Straight-line from here on (no early exits): lanes of a warp test different triangles, and a lane that leaves early
    // only idles until the others are done.
    const bool outside = (U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f);
    const float det = __fadd_rn(__fadd_rn(U, V), W);

    const float Az = __fmul_rn(s.Sz, Akz);
    const float Bz = __fmul_rn(s.Sz, Bkz);
    const float Cz = __fmul_rn(s.Sz, Ckz);
    const float T = __fadd_rn(__fadd_rn(__fmul_rn(U, Az), __fmul_rn(V, Bz)), __fmul_rn(W, Cz));

    const float rcp_det = __fdiv_rn(1.0f, det);
    t = __fmul_rn(T, rcp_det);
    u = __fmul_rn(V, rcp_det); // weight of p1
    v = __fmul_rn(W, rcp_det); // weight of p2
    return !outside && det != 0.0f;
}

// ---- traversal ---------------------------------------------------------------------------------

#ifndef BPT_TRACE_MIN_BLOCKS
#define BPT_TRACE_MIN_BLOCKS 8
#endif
#ifndef BPT_TRAVERSAL_BUDGET
#define BPT_TRAVERSAL_BUDGET 48
#endif
constexpr int TRACE_BLOCK = 128;
#ifndef BPT_MIN_ACTIVE_LANES
#define BPT_MIN_ACTIVE_LANES 12
#endif
#ifndef BPT_STACK_SMEM
#define BPT_STACK_SMEM 32
#endif
constexpr int STACK_SMEM = BPT_STACK_SMEM;
constexpr int STACK_LOCAL = 72;
constexpr int TRAVERSAL_BUDGET = BPT_TRAVERSAL_BUDGET; // node visits between two refills of a warp's idle lanes
constexpr int NODE_EMPTY = (int)0x80000000;   // "no node": neither an inner index (>= 0) nor a leaf (~packed)

// Leaf link: ~(first_triangle | (count - 1) << 28), always in [-2^30, -1].
BPT_HD int pack_leaf(int first, int count) { return ~(first | ((count - 1) << 28)); }
BPT_HD bool is_leaf(int link) { return link < 0 && link != NODE_EMPTY; }
BPT_HD int leaf_first(int link) { return (~link) & 0x0fffffff; }
BPT_HD int leaf_count(int link) { return ((~link) >> 28) + 1; }
constexpr int MAX_TRIANGLES = 0x0fffffff;

// The first STACK_SMEM entries of a lane's stack live in shared memory; deeper ones spill to a per-thread array in local
// memory. The spill array is referenced through a pointer so that the rest of the traversal state stays in registers
// (a struct that contains a dynamically indexed array is placed in local memory as a whole).
struct TraversalStack {
    int* smem;                // [STACK_SMEM][TRACE_BLOCK], this thread's column
    int* spill;               // [STACK_LOCAL + 1]: the last slot is the "a push was dropped" flag
    int sp;
    // The build only hands out hierarchies whose depth fits (bpt_bvh.cu: 3 x wide levels + 1 <= STACK_SMEM + STACK_LOCAL, binary
    // depth <= 96, else it falls back), so the last branch is unreachable by construction; if it ever ran, the dropped subtree
    // would be a wrong image, so it is counted (bpt_counters.traversal_stack_overflows) instead of passing silently.
    // (The overflow is recorded in one extra slot of the spill array and counted when the ray retires: an atomic on a global
    // counter right here cost 17 % of the closest-hit kernel - measured - although it never executes.)
    BPT_D void push(int link) {
        if (sp < STACK_SMEM) smem[sp * TRACE_BLOCK] = link;
        else if (sp - STACK_SMEM < STACK_LOCAL) spill[sp - STACK_SMEM] = link;
        else spill[STACK_LOCAL] = 1;
        ++sp;
    }
    BPT_D int pop() {
        if (sp == 0) return NODE_EMPTY;
        --sp;
        return sp < STACK_SMEM ? smem[sp * TRACE_BLOCK] : spill[min(sp - STACK_SMEM, STACK_LOCAL - 1)];
    }
};

struct AccelView {
    const BvhNode* __restrict__ nodes;
    const WideNode* __restrict__ wide; // four-wide nodes; nullptr = traverse the binary nodes
    const CwNode* __restrict__ cw;     // compressed eight-wide nodes (bpt_cw.cuh); when set, the kernels instantiated for them run
    uint32_t cw_exponent_word;         // CW_EXPONENT_WORD as a run-time value (bpt_cw.cuh: cw_plane_float)
    const TraceTriangle* __restrict__ triangles; // in the order of the hierarchy in use (Morton order, or per node for `cw`)
    const Material* __restrict__ materials; // with `textures`: only read by any-hit rays that meet a coverage-textured material
    TextureView textures;
    unsigned long long* overflow_counter; // pushes beyond the traversal stack (never, see TraversalStack::push)
    int min_active; // see traversal_min_active_for
    int budget; // upper bound on the node visits between two refills of a warp's idle lanes
};
#ifndef BPT_BUDGET_SMALL
#define BPT_BUDGET_SMALL 48
#endif
#ifndef BPT_BUDGET_LARGE
#define BPT_BUDGET_LARGE 96
#endif
BPT_HD int traversal_budget_for(long long triangle_count) { return triangle_count > 8000000ll ? BPT_BUDGET_LARGE : BPT_BUDGET_SMALL; }
// A round also ends once fewer than this many lanes of the warp are still traversing (measured on B200: +4 % on the 20 k
// and 1 M triangle scenes with 12 lanes; +3 % on the 50 M triangle scene with 4). With this trigger in place the node
// budget is only a backstop, and larger budgets (48 / 96) measured 1-3 % faster than 24 / 48.
#ifndef BPT_MIN_ACTIVE_LANES_LARGE
#define BPT_MIN_ACTIVE_LANES_LARGE 4
#endif
BPT_HD int traversal_min_active_for(long long triangle_count) { return triangle_count > 8000000ll ? BPT_MIN_ACTIVE_LANES_LARGE : BPT_MIN_ACTIVE_LANES; }

inline AccelView accel_view(const Context* ctx) {
    AccelView a;
    a.nodes = ctx->accel.nodes.ptr; a.triangles = ctx->accel.triangles.ptr;
    a.wide = ctx->accel.wide_levels > 0 ? ctx->accel.wide_nodes.ptr : nullptr;
    a.cw = ctx->accel.cw_levels > 0 ? ctx->accel.cw_nodes.ptr : nullptr;
    a.cw_exponent_word = CW_EXPONENT_WORD;
    a.materials = ctx->materials.ptr;
    a.textures.objects = ctx->texture_objects.ptr;
    a.textures.uv = ctx->accel.has_uv ? ctx->accel.shade_uv.ptr : nullptr;
    a.overflow_counter = reinterpret_cast<unsigned long long*>(ctx->device_counters) + 5;
    a.min_active = traversal_min_active_for(ctx->accel.triangle_count);
    a.budget = traversal_budget_for(ctx->accel.triangle_count);
    return a;
}

BPT_D float4 ldg4(const float4* p) { return __ldg(p); }

// Slab test against one child box, (plane - origin) * inv_d per plane. (A fused lo * inv_d - o * inv_d form needs an absolute
// error pad proportional to |o * inv_d|; measured on B200 it made the 1M / 50M triangle scenes 2-3x slower because rays with
// one small direction component then visit large parts of the tree.) Conservative: the far distance is padded by a few ulp (Ize, "Robust BVH ray
// traversal", JCGT 2013) and the near one shrunk, so a triangle the watertight test would accept is never culled.
// An absent child (scenes with fewer than two triangles) is a point box at 3e38 and never passes.
BPT_D bool slab(float3 lo, float3 hi, float3 o, float3 inv_d, float tmin, float tmax, float& tnear) {
    float t0x = (lo.x - o.x) * inv_d.x, t1x = (hi.x - o.x) * inv_d.x;
    float t0y = (lo.y - o.y) * inv_d.y, t1y = (hi.y - o.y) * inv_d.y;
    float t0z = (lo.z - o.z) * inv_d.z, t1z = (hi.z - o.z) * inv_d.z;
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
    tn = fmaxf(tmin, tn * 0.99999905f);
    tf = fminf(tmax, tf * 1.00000095f);
    tnear = tn;
    return tn <= tf;
}

// Everything one lane needs to carry a ray through the hierarchy.
template <bool ANY_HIT>
struct Traversal {
    Ray ray;
    RayShear shear;
    float3 inv_d;
    float tmax;          // closest hit: shrinks to the best t; any hit: < 0 once the ray is blocked
    Hit hit;
    float transmission;
    float termination_weight; // any hit: the largest channel of the radiance the ray carries (1 for plain occlusion queries)
    int skip_primitive;
    int node;
    int postponed;       // a leaf found while other lanes were still descending; NODE_EMPTY when none
    TraversalStack stack;
#ifdef BPT_TRAVERSAL_STATS
    unsigned int stat_nodes, stat_triangles;
#endif
    static constexpr bool COMPRESSED = false;

    // The part of the stack that does not fit in shared memory, declared by whoever drives the traversal.
    struct Spill { int slots[STACK_LOCAL + 1]; };
    // `smem_column`: this thread's column of the CTA's [STACK_SMEM][TRACE_BLOCK] array.
    BPT_D void attach(int* smem_column, Spill& spill) {
        stack.smem = smem_column; stack.spill = spill.slots; stack.sp = 0;
        spill.slots[STACK_LOCAL] = 0;
        node = NODE_EMPTY; postponed = NODE_EMPTY;
    }
    BPT_D bool take_overflow() {
        if (stack.spill[STACK_LOCAL] == 0) return false;
        stack.spill[STACK_LOCAL] = 0;
        return true;
    }
    BPT_D float3 ray_origin() const { return ray.origin; }

    BPT_D void begin(const AccelView&, const Ray& r, int skip) {
        ray = r;
        shear = make_ray_shear(r.direction);
        inv_d = f3(__fdiv_rn(1.0f, r.direction.x), __fdiv_rn(1.0f, r.direction.y), __fdiv_rn(1.0f, r.direction.z)); // IEEE whatever -prec-div says
        tmax = r.tmax;
        hit.t = r.tmax; hit.primitive = 0x7fffffff; hit.u = hit.v = 0.0f;
        transmission = 1.0f;
        termination_weight = 1.0f;
        skip_primitive = skip;
        stack.sp = 0;
        node = 0; // root
        postponed = NODE_EMPTY;
#ifdef BPT_TRAVERSAL_STATS
        stat_nodes = stat_triangles = 0;
#endif
    }

    // One inner-node visit: tests both children, descends into the nearer hit child, defers the other.
    BPT_D void inner_step(const AccelView& a) {
#ifdef BPT_TRAVERSAL_STATS
        ++stat_nodes;
#endif
        const float4* n = reinterpret_cast<const float4*>(a.nodes + node);
        float4 n0 = ldg4(n), n1 = ldg4(n + 1), n2 = ldg4(n + 2);
        int2 links = __ldg(reinterpret_cast<const int2*>(n + 3));
        float tn_l, tn_r;
        bool hit_l = slab(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), ray.origin, inv_d, ray.tmin, tmax, tn_l);
        bool hit_r = slab(f3(n1.z, n1.w, n2.x), f3(n2.y, n2.z, n2.w), ray.origin, inv_d, ray.tmin, tmax, tn_r);
        if (hit_l && hit_r) {
            bool left_first = tn_l <= tn_r;
            node = left_first ? links.x : links.y;
            stack.push(left_first ? links.y : links.x);
        } else if (hit_l)
            node = links.x;
        else if (hit_r)
            node = links.y;
        else
            node = stack.pop();
    }

    // One visit of a four-wide node: tests the four children, continues with the nearest one that is hit and defers the
    // others, farthest first, so that they come off the stack nearest first.
    // (Measured in round 2 and dropped: letting the ray pick the entry / exit plane float4 of each axis by the sign of its
    // direction removes the 24 per-axis min / max of the four slab tests, but the six loads then need computed addresses
    // instead of [node + immediate]: extend +4 % SLOWER on the 1 M and 20 k triangle scenes, +7 % on the 50 M one.)
    BPT_D void wide_step(const AccelView& a) {
#ifdef BPT_TRAVERSAL_STATS
        ++stat_nodes;
#endif
        const float4* n = reinterpret_cast<const float4*>(a.wide + node);
        const float4 lox = ldg4(n), loy = ldg4(n + 1), loz = ldg4(n + 2), hix = ldg4(n + 3), hiy = ldg4(n + 4), hiz = ldg4(n + 5);
        const int4 links = __ldg(reinterpret_cast<const int4*>(n + 6));
        float t0, t1, t2, t3;
        const bool h0 = slab(f3(lox.x, loy.x, loz.x), f3(hix.x, hiy.x, hiz.x), ray.origin, inv_d, ray.tmin, tmax, t0);
        const bool h1 = slab(f3(lox.y, loy.y, loz.y), f3(hix.y, hiy.y, hiz.y), ray.origin, inv_d, ray.tmin, tmax, t1);
        const bool h2 = slab(f3(lox.z, loy.z, loz.z), f3(hix.z, hiy.z, hiz.z), ray.origin, inv_d, ray.tmin, tmax, t2);
        const bool h3 = slab(f3(lox.w, loy.w, loz.w), f3(hix.w, hiy.w, hiz.w), ray.origin, inv_d, ray.tmin, tmax, t3);
        const int hits = int(h0) + int(h1) + int(h2) + int(h3);
        if (hits == 0) { node = stack.pop(); return; }
        // children that are missed sort to the end
        t0 = h0 ? t0 : FLT_MAX; t1 = h1 ? t1 : FLT_MAX; t2 = h2 ? t2 : FLT_MAX; t3 = h3 ? t3 : FLT_MAX;
        int l0 = links.x, l1 = links.y, l2 = links.z, l3 = links.w;
        if (hits > 1) { // sorting network for four keys: (0,1) (2,3) (0,2) (1,3) (1,2)
#define BPT_CSWAP(ta, la, tb, lb) { bool s = tb < ta; float tt = s ? tb : ta; tb = s ? ta : tb; ta = tt; int ll = s ? lb : la; lb = s ? la : lb; la = ll; }
            BPT_CSWAP(t0, l0, t1, l1) BPT_CSWAP(t2, l2, t3, l3) BPT_CSWAP(t0, l0, t2, l2) BPT_CSWAP(t1, l1, t3, l3) BPT_CSWAP(t1, l1, t2, l2)
#undef BPT_CSWAP
            if (hits > 3) stack.push(l3);
            if (hits > 2) stack.push(l2);
            stack.push(l1);
            node = l0;
        } else
            node = h0 ? l0 : (h1 ? l1 : (h2 ? l2 : l3));
    }

    // Intersects the triangles of one leaf. Returns false when an any-hit ray got blocked (traversal is over).
    BPT_D bool intersect_leaf(const AccelView& a, const float* __restrict__ coverage_by_material, int leaf) {
        const int first = leaf_first(leaf), count = leaf_count(leaf);
        for (int i = 0; i < count; ++i) {
            const float4* tri = reinterpret_cast<const float4*>(a.triangles + first + i);
            float4 v0 = ldg4(tri), v1 = ldg4(tri + 1), v2 = ldg4(tri + 2);
            int primitive = __float_as_int(v0.w);
#ifdef BPT_TRAVERSAL_STATS
            ++stat_triangles;
#endif
            float t, u, v;
            bool candidate = watertight_triangle(shear, ray.origin, f3(v0), f3(v1), f3(v2), t, u, v) && primitive != skip_primitive;
            if (ANY_HIT) {
                if (candidate && t > ray.tmin && t < ray.tmax) {
                    // shadow_any_hit, MonteCarlo.cu:278-285: the payload radiance is attenuated by (1 - coverage) and the ray ends
                    // once all its channels are below 1e-7, i.e. once largest channel x transmission is; opaque surfaces end it.
                    float coverage = coverage_by_material[__float_as_int(v1.w)];
                    if (coverage < 0.0f) // coverage texture: Material::get_coverage(texcoord), Types.h:405-414
                        coverage = material_coverage(a.materials[__float_as_int(v1.w)], a.textures, interpolate_texcoord(a.textures, primitive, u, v));
                    transmission *= 1.0f - coverage;
                    if (transmission * termination_weight < 0.0000001f) { transmission = 0.0f; return false; }
                }
            } else if (candidate && t > ray.tmin && (t < hit.t || (t == hit.t && primitive < hit.primitive))) {
                // hit.t starts at ray.tmax and hit.primitive at INT_MAX, so a first hit needs t < tmax.
                hit.t = t; hit.primitive = primitive; hit.u = u; hit.v = v;
                tmax = t;
            }
        }
        return true;
    }

    // Runs up to `budget` inner-node visits (and the leaves met on the way). Returns when the ray is done or the budget
    // is used up; `node == NODE_EMPTY && postponed == NODE_EMPTY` tells which.
    // Speculative traversal (Aila and Laine): a lane that reaches a leaf parks it in `postponed` and keeps descending while
    // other lanes of the warp are still looking for theirs, so the inner-node loop runs with more lanes active.
    // (A second parked leaf per lane was measured in round 2: extend +1.8 % slower on the 1 M triangle scene, +2 % on the 20 k one;
    // the extra speculative node visits cost more than the wider leaf phase returns.)
    BPT_D void run(const AccelView& a, const float* __restrict__ coverage_by_material, int budget, int min_active = 0) {
        while ((node != NODE_EMPTY || postponed != NODE_EMPTY) && budget > 0) {
            while (node >= 0 && budget > 0) {
                if (a.wide != nullptr) wide_step(a); else inner_step(a);
                --budget;
                if (postponed == NODE_EMPTY && is_leaf(node)) { postponed = node; node = stack.pop(); }
                const unsigned int active = __activemask();
                // Too few lanes left in the round: hand the warp back to the driver, which refills the idle lanes.
                if (__popc(active) < min_active) budget = 0;
                if (!__any_sync(active, postponed == NODE_EMPTY && node >= 0))
                    break;
            }
            if (postponed == NODE_EMPTY && is_leaf(node)) { postponed = node; node = stack.pop(); }
            // One leaf loop for both the parked leaf and a leaf that is the current node, so that all lanes holding a leaf
            // run the triangle tests together.
            while (postponed != NODE_EMPTY) {
                bool alive = intersect_leaf(a, coverage_by_material, postponed);
                postponed = NODE_EMPTY;
                if (!alive) { node = NODE_EMPTY; stack.sp = 0; }
                else if (is_leaf(node)) { postponed = node; node = stack.pop(); }
            }
        }
    }

    BPT_D bool finished() const { return node == NODE_EMPTY && postponed == NODE_EMPTY; }

    BPT_D Hit result() const {
        Hit h = hit;
        if (!ANY_HIT && h.primitive == 0x7fffffff) h.primitive = -1;
        return h;
    }
};

// ---- compressed eight-wide nodes -----------------------------------------------------------------------------------
// The node format and the ray / node test are in bpt_cw.cuh; this is the per-lane state and the warp-level loop.
// * The stack holds 64-bit groups: (first child node, hit children by priority | the parent's imask) or (first triangle,
//   hit triangles). A node visit pushes at most one node group - the rest of the group the visited child came from - and
//   needs no sorting: the octant slots of the encoder already are a front-to-back order.
// * Control flow is "while-while" with speculation, as in Traversal: a lane that has found triangles keeps visiting nodes
//   while other lanes of its warp are still searching for theirs (further triangles it meets wait on its stack), and all
//   lanes then test their triangles together. Measured on B200 (configs[2] / configs[1], Msamples/s): no speculation
//   603 / 479, one parked group and then wait 691 / 500, unbounded (this) 703 / 504; ending the node loop while 3 / 6 / 10
//   lanes are still searching 698 / 696 / 690 on configs[2].
// * Measured and dropped: triangle tests shared by the warp (the (ray, triangle) pairs of all lanes listed in shared memory and
//   dealt out 32 at a time, also to lanes without a ray; owners fold the results in list order). Bit-exact, 60 parity tests
//   green, and slower: configs[2] 704 -> 662, configs[3] 580 -> 542 Msamples/s (closest-hit kernel 1.78 -> 1.81 ms, any-hit
//   0.92 -> 0.96): the prefix sum, the list, two barriers per round and the fold loop cost what the wider tests return.
//   Also dropped: prefetch.global.L2 of a found group's first (and last) triangle at the end of the node visit: configs[3]
//   579 -> 574 (568), configs[2] 703 -> 696 (693).
constexpr int CW_STACK_SMEM = STACK_SMEM / 2;  // 8-byte entries in the same shared memory as Traversal's 4-byte ones
constexpr int CW_STACK_LOCAL = 48;             // bpt_bvh.cu only hands out trees with 2 * levels + 2 <= CW_STACK_SMEM + CW_STACK_LOCAL

struct CwStack {
    uint2* smem;   // [CW_STACK_SMEM][TRACE_BLOCK], this thread's column
    uint2* spill;  // [CW_STACK_LOCAL + 1]: the last slot is the "a push was dropped" flag (see TraversalStack)
    int sp;
    BPT_D void push(uint2 group) {
        if (sp < CW_STACK_SMEM) smem[sp * TRACE_BLOCK] = group;
        else if (sp - CW_STACK_SMEM < CW_STACK_LOCAL) spill[sp - CW_STACK_SMEM] = group;
        else spill[CW_STACK_LOCAL].x = 1u;
        ++sp;
    }
    BPT_D uint2 pop() {
        if (sp == 0) return make_uint2(0u, 0u);
        --sp;
        return sp < CW_STACK_SMEM ? smem[sp * TRACE_BLOCK] : spill[min(sp - CW_STACK_SMEM, CW_STACK_LOCAL - 1)];
    }
};

template <bool ANY_HIT>
struct TraversalCW {
    static constexpr bool COMPRESSED = true;
    CwRay cw;            // origin, reciprocal direction, octant
    float tmin, ray_tmax;
    RayShear shear;
    float tmax;          // closest hit: shrinks to the best t; any hit: the ray's tmax
    Hit hit;
    float transmission;
    float termination_weight;
    int skip_primitive;
    uint2 ngroup;        // group being worked on (y == 0: none; no bits in 24..31: a triangle group that came off the stack)
    uint2 tgroup;        // triangles found and not tested yet
    CwStack stack;
#ifdef BPT_TRAVERSAL_STATS
    unsigned int stat_nodes, stat_triangles;
#endif

    struct Spill { uint2 slots[CW_STACK_LOCAL + 1]; };
    BPT_D void attach(int* smem_column, Spill& spill) {
        stack.smem = reinterpret_cast<uint2*>(smem_column - threadIdx.x) + threadIdx.x;
        stack.spill = spill.slots; stack.sp = 0;
        spill.slots[CW_STACK_LOCAL].x = 0u;
        ngroup = make_uint2(0u, 0u); tgroup = make_uint2(0u, 0u);
    }
    BPT_D bool take_overflow() {
        if (stack.spill[CW_STACK_LOCAL].x == 0u) return false;
        stack.spill[CW_STACK_LOCAL].x = 0u;
        return true;
    }
    BPT_D float3 ray_origin() const { return cw.origin; }

    BPT_D void begin(const AccelView& a, const Ray& r, int skip) {
        cw = cw_make_ray(r.origin, r.direction);
        shear = make_ray_shear(r.direction);
        tmin = r.tmin; ray_tmax = r.tmax; tmax = r.tmax;
        hit.t = r.tmax; hit.primitive = 0x7fffffff; hit.u = hit.v = 0.0f;
        transmission = 1.0f;
        termination_weight = 1.0f;
        skip_primitive = skip;
        stack.sp = 0;
        ngroup = make_uint2(0u, 0x80000000u); // the root: "child 7 ^ octant of a group whose imask is empty" = node 0
        tgroup = make_uint2(0u, 0u);
#ifdef BPT_TRAVERSAL_STATS
        stat_nodes = stat_triangles = 0;
#endif
    }

    // Visits the nearest unvisited child of `ngroup`: the rest of the group goes to the stack, `ngroup` becomes the group
    // of that child's own inner children and `found` the triangles of its leaf children.
    BPT_D void node_step(const AccelView& a, uint2& found) {
#ifdef BPT_TRAVERSAL_STATS
        ++stat_nodes;
#endif
        const uint32_t index = cw_next_child(ngroup, cw);
        if (cw_is_node_group(ngroup)) stack.push(ngroup);
        const uint4* n = reinterpret_cast<const uint4*>(a.cw + index);
        const uint4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3), n4 = __ldg(n + 4);
        const uint32_t hits = cw_intersect_children(n0, n1, n2, n3, n4, cw, tmin, tmax, a.cw_exponent_word);
        ngroup = make_uint2(n1.x, (hits & 0xff000000u) ? ((hits & 0xff000000u) | (n0.w >> 24)) : 0u);
        found = make_uint2(n1.y, hits & 0x00ffffffu);
    }

    // Tests one triangle (Traversal::intersect_leaf has the same body). Returns false when an any-hit ray got blocked.
    BPT_D bool test_triangle(const AccelView& a, const float* __restrict__ coverage_by_material, uint32_t index) {
        const float4* tri = reinterpret_cast<const float4*>(a.triangles + index);
        const float4 v0 = ldg4(tri), v1 = ldg4(tri + 1), v2 = ldg4(tri + 2);
        const int primitive = __float_as_int(v0.w);
#ifdef BPT_TRAVERSAL_STATS
        ++stat_triangles;
#endif
        float t, u, v;
        const bool candidate = watertight_triangle(shear, cw.origin, f3(v0), f3(v1), f3(v2), t, u, v) && primitive != skip_primitive;
        if (ANY_HIT) {
            if (candidate && t > tmin && t < ray_tmax) { // shadow_any_hit, MonteCarlo.cu:278-285
                float coverage = coverage_by_material[__float_as_int(v1.w)];
                if (coverage < 0.0f)
                    coverage = material_coverage(a.materials[__float_as_int(v1.w)], a.textures, interpolate_texcoord(a.textures, primitive, u, v));
                transmission *= 1.0f - coverage;
                if (transmission * termination_weight < 0.0000001f) { transmission = 0.0f; return false; }
            }
        } else if (candidate && t > tmin && (t < hit.t || (t == hit.t && primitive < hit.primitive))) {
            hit.t = t; hit.primitive = primitive; hit.u = u; hit.v = v;
            tmax = t;
        }
        return true;
    }

    BPT_D void run(const AccelView& a, const float* __restrict__ coverage_by_material, int budget, int min_active = 0) {
        while ((ngroup.y | tgroup.y) != 0u && budget > 0) {
            while (cw_is_node_group(ngroup) && budget > 0) {
                uint2 found;
                node_step(a, found);
                --budget;
                if (found.y != 0u) {
                    if (tgroup.y == 0u) tgroup = found;
                    else stack.push(found); // further triangles wait on the stack while the lane keeps visiting nodes
                }
                if (ngroup.y == 0u) ngroup = stack.pop();
                const unsigned int active = __activemask();
                if (__popc(active) < min_active) budget = 0;
                if (!__any_sync(active, tgroup.y == 0u && cw_is_node_group(ngroup)))
                    break;
            }
            // One triangle loop for the parked triangles and for triangle groups that came off the stack.
            while (true) {
                if (tgroup.y == 0u) {
                    if (ngroup.y == 0u || cw_is_node_group(ngroup)) break;
                    tgroup = ngroup; ngroup = stack.pop();
                }
                const int bit = 31 - __clz((int)tgroup.y);
                tgroup.y &= ~(1u << bit);
                if (!test_triangle(a, coverage_by_material, tgroup.x + (uint32_t)bit)) { ngroup.y = 0u; tgroup.y = 0u; stack.sp = 0; break; }
            }
        }
    }

    BPT_D bool finished() const { return (ngroup.y | tgroup.y) == 0u; }

    BPT_D Hit result() const {
        Hit h = hit;
        if (!ANY_HIT && h.primitive == 0x7fffffff) h.primitive = -1;
        return h;
    }
};

// The traversal a kernel instantiated for a node format runs.
template <bool ANY_HIT, bool COMPRESSED> struct TraversalFor { typedef Traversal<ANY_HIT> type; };
template <bool ANY_HIT> struct TraversalFor<ANY_HIT, true> { typedef TraversalCW<ANY_HIT> type; };

// Persistent-thread driver. `Trav` is Traversal<ANY_HIT> or TraversalCW<ANY_HIT>. `Source` supplies rays and consumes results:
//   void load(unsigned int index, Ray& ray, int& skip_primitive)
//   float termination_weight(unsigned int index)        (any hit only: the largest radiance channel the ray carries)
//   void store(unsigned int index, const Trav& traversal)
// `fetch_counter` is a zero-initialised global counter shared by all CTAs of the launch. Must be called by whole warps.
template <bool ANY_HIT, class Trav, class Source>
BPT_D void traverse_queue_with(const AccelView& a, const float* __restrict__ coverage_by_material, Source& source, unsigned int count,
                               unsigned int* fetch_counter, int* stack_smem, int budget = TRAVERSAL_BUDGET) {
    typename Trav::Spill spill;
    Trav tr;
    tr.attach(stack_smem, spill);
    unsigned int index = 0;
    bool has_ray = false, exhausted = false;
    const int lane = threadIdx.x & 31;

    while (true) {
        // Retire finished rays and refill idle lanes: one atomic per warp.
        if (has_ray && tr.finished()) {
            source.store(index, tr); has_ray = false;
            if (tr.take_overflow()) atomicAdd(a.overflow_counter, 1ull);
        }
        unsigned int idle = __ballot_sync(0xffffffffu, !has_ray && !exhausted);
        if (idle) {
            int leader = __ffs(idle) - 1;
            unsigned int base = 0;
            if (lane == leader) base = atomicAdd(fetch_counter, __popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!has_ray && !exhausted) {
                index = base + __popc(idle & ((1u << lane) - 1u));
                if (index < count) {
                    Ray ray; int skip;
                    source.load(index, ray, skip);
                    tr.begin(a, ray, skip);
                    if (ANY_HIT) tr.termination_weight = source.termination_weight(index);
                    has_ray = true;
                } else
                    exhausted = true;
            }
        }
        if (__ballot_sync(0xffffffffu, has_ray) == 0)
            break;
        const bool exhausted_warp = __any_sync(0xffffffffu, exhausted);
        if (has_ray)
            tr.run(a, coverage_by_material, budget, exhausted_warp ? 0 : a.min_active);
    }
}

template <bool ANY_HIT, class Source>
BPT_D void traverse_queue(const AccelView& a, const float* __restrict__ coverage_by_material, Source& source, unsigned int count,
                          unsigned int* fetch_counter, int* stack_smem, int budget = TRAVERSAL_BUDGET) {
    traverse_queue_with<ANY_HIT, Traversal<ANY_HIT>>(a, coverage_by_material, source, count, fetch_counter, stack_smem, budget);
}

} // namespace bpt
