// CPU test of the compressed eight-wide node format (bifrost3d_b200/csrc/bpt_cw.cuh): the encoder and the ray / node test
// are __host__ __device__, so the same code that the traversal kernels run is checked here without a GPU.
//   1. every child box that a ray's exact slab test passes is reported by cw_intersect_children (conservative), the bits
//      of empty slots never appear, and inner / triangle bits land where the meta bytes say;
//   2. a whole tree (median-split hierarchy -> eight-wide collapse -> cw_encode, children and triangles laid out as
//      bpt_bvh.cu does) traversed with cw_next_child / cw_intersect_children returns the brute-force closest hit
//      (minimum over (t, primitive)) for random, axis-parallel and surface-grazing rays;
//   3. the front-to-back property of the octant slots: the first child visited is never behind all others.
// Build + run: g++ -O2 -std=c++17 -I/usr/local/cuda/include tests/host/cw_host_test.cpp -o cw_host_test && ./cw_host_test
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <cuda_runtime.h>
#include "../../bifrost3d_b200/csrc/bpt_cw.cuh"

using namespace bpt;

struct Tri { float3 v[3]; int id; };
struct Box { float3 lo, hi; };
struct BinNode { Box box; int left = -1, right = -1; int first = 0, count = 0; }; // count > 0: leaf

static Box box_of(const Tri& t) {
    Box b; b.lo = min3(min3(t.v[0], t.v[1]), t.v[2]); b.hi = max3(max3(t.v[0], t.v[1]), t.v[2]); return b;
}
static Box merge(const Box& a, const Box& b) { return { min3(a.lo, b.lo), max3(a.hi, b.hi) }; }
static float area(const Box& b) { float3 d = b.hi - b.lo; return d.x * d.y + d.y * d.z + d.z * d.x; }

static int build_binary(std::vector<BinNode>& nodes, std::vector<Tri>& tris, int first, int count, int leaf_max) {
    BinNode n; n.box = box_of(tris[first]);
    for (int i = 1; i < count; ++i) n.box = merge(n.box, box_of(tris[first + i]));
    int index = (int)nodes.size(); nodes.push_back(n);
    if (count <= leaf_max) { nodes[index].first = first; nodes[index].count = count; return index; }
    float3 d = n.box.hi - n.box.lo;
    int axis = d.x >= d.y && d.x >= d.z ? 0 : (d.y >= d.z ? 1 : 2);
    auto key = [&](const Tri& t) { float3 c = t.v[0] + t.v[1] + t.v[2]; return axis == 0 ? c.x : (axis == 1 ? c.y : c.z); };
    std::nth_element(tris.begin() + first, tris.begin() + first + count / 2, tris.begin() + first + count, [&](const Tri& a, const Tri& b) { return key(a) < key(b); });
    int l = build_binary(nodes, tris, first, count / 2, leaf_max);
    int r = build_binary(nodes, tris, first + count / 2, count - count / 2, leaf_max);
    nodes[index].left = l; nodes[index].right = r;
    return index;
}

struct CwTree { std::vector<CwNode> nodes; std::vector<Tri> tris; int levels = 0; };

// Same layout rules as the device build: a node's inner children are consecutive nodes in slot order, its leaf
// children's triangles are consecutive in slot order.
static void collapse(const std::vector<BinNode>& bin, const std::vector<Tri>& sorted, CwTree& out) {
    struct Task { int cw, binary; };
    std::vector<Task> level = { { 0, 0 } };
    out.nodes.resize(1);
    while (!level.empty()) {
        ++out.levels;
        std::vector<Task> next;
        for (const Task& task : level) {
            int link[8]; int k = 0;
            link[k++] = bin[task.binary].left; link[k++] = bin[task.binary].right;
            while (k < 8) {
                int best = -1; float best_area = -1.0f;
                for (int i = 0; i < k; ++i) if (bin[link[i]].count == 0 && area(bin[link[i]].box) > best_area) { best_area = area(bin[link[i]].box); best = i; }
                if (best < 0) break;
                int opened = link[best];
                link[best] = bin[opened].left; link[k++] = bin[opened].right;
            }
            float3 lo[8], hi[8]; int triangles[8];
            for (int i = 0; i < k; ++i) { lo[i] = bin[link[i]].box.lo; hi[i] = bin[link[i]].box.hi; triangles[i] = bin[link[i]].count; }
            CwNode node; CwPlacement place;
            cw_encode(k, lo, hi, triangles, node, place);
            node.child_base = (uint32_t)out.nodes.size();
            node.triangle_base = (uint32_t)out.tris.size();
            out.nodes.resize(out.nodes.size() + place.inner_count);
            out.tris.resize(out.tris.size() + place.triangle_count);
            for (int i = 0; i < k; ++i) {
                if (triangles[i] == 0) next.push_back({ (int)node.child_base + place.offset[i], link[i] });
                else for (int t = 0; t < triangles[i]; ++t) out.tris[node.triangle_base + place.offset[i] + t] = sorted[bin[link[i]].first + t];
            }
            out.nodes[task.cw] = node;
        }
        level.swap(next);
    }
}

// Moeller-Trumbore in double: the same function serves the brute force and the traversal, so equal results mean that the
// traversal never skipped a triangle it should have tested.
static bool hit_triangle(const Tri& tri, float3 o, float3 d, double& t) {
    double e1[3] = { (double)tri.v[1].x - tri.v[0].x, (double)tri.v[1].y - tri.v[0].y, (double)tri.v[1].z - tri.v[0].z };
    double e2[3] = { (double)tri.v[2].x - tri.v[0].x, (double)tri.v[2].y - tri.v[0].y, (double)tri.v[2].z - tri.v[0].z };
    double D[3] = { d.x, d.y, d.z }, O[3] = { (double)o.x - tri.v[0].x, (double)o.y - tri.v[0].y, (double)o.z - tri.v[0].z };
    double p[3] = { D[1] * e2[2] - D[2] * e2[1], D[2] * e2[0] - D[0] * e2[2], D[0] * e2[1] - D[1] * e2[0] };
    double det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (det == 0.0) return false;
    double u = (O[0] * p[0] + O[1] * p[1] + O[2] * p[2]) / det;
    double q[3] = { O[1] * e1[2] - O[2] * e1[1], O[2] * e1[0] - O[0] * e1[2], O[0] * e1[1] - O[1] * e1[0] };
    double v = (D[0] * q[0] + D[1] * q[1] + D[2] * q[2]) / det;
    if (u < 0.0 || v < 0.0 || u + v > 1.0) return false;
    t = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) / det;
    return true;
}

struct Result { float t; int id; long nodes, tests; };

static Result brute(const std::vector<Tri>& tris, float3 o, float3 d, float tmin, float tmax) {
    Result r = { tmax, 0x7fffffff, 0, 0 };
    for (const Tri& tri : tris) {
        double t;
        if (hit_triangle(tri, o, d, t)) { float tf = (float)t; if (tf > tmin && (tf < r.t || (tf == r.t && tri.id < r.id))) { r.t = tf; r.id = tri.id; } }
    }
    return r;
}

static uint4 word(const CwNode& n, int i) { uint4 w; memcpy(&w, reinterpret_cast<const char*>(&n) + 16 * i, 16); return w; }

static Result traverse(const CwTree& tree, float3 o, float3 d, float tmin, float tmax, size_t* max_stack = nullptr) {
    Result r = { tmax, 0x7fffffff, 0, 0 };
    const CwRay ray = cw_make_ray(o, d);
    std::vector<uint2> stack;
    uint2 group = make_uint2(0u, 0x80000000u);
    while (true) {
        uint2 triangles = make_uint2(0u, 0u);
        if (cw_is_node_group(group)) {
            const uint32_t index = cw_next_child(group, ray);
            if (cw_is_node_group(group)) stack.push_back(group);
            const CwNode& n = tree.nodes[index];
            const uint4 n0 = word(n, 0), n1 = word(n, 1);
            const uint32_t hits = cw_intersect_children(n0, n1, word(n, 2), word(n, 3), word(n, 4), ray, tmin, r.t);
            ++r.nodes;
            group = make_uint2(n1.x, (hits & 0xff000000u) ? ((hits & 0xff000000u) | (n0.w >> 24)) : 0u);
            triangles = make_uint2(n1.y, hits & 0x00ffffffu);
        } else { triangles = group; group = make_uint2(0u, 0u); }
        while (triangles.y) {
            const int bit = cw_highest_bit(triangles.y); triangles.y &= ~(1u << bit);
            const Tri& tri = tree.tris[triangles.x + bit];
            double t; ++r.tests;
            if (hit_triangle(tri, o, d, t)) { float tf = (float)t; if (tf > tmin && (tf < r.t || (tf == r.t && tri.id < r.id))) { r.t = tf; r.id = tri.id; } }
        }
        if (!cw_is_node_group(group)) {
            if (stack.empty()) break;
            group = stack.back(); stack.pop_back();
        }
        if (max_stack) *max_stack = std::max(*max_stack, stack.size());
    }
    return r;
}

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { if (++failures < 20) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

static std::mt19937 rng(1234);
static float uniform(float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); }
static float3 random_direction(int kind) {
    float3 d;
    do { d = f3(uniform(-1, 1), uniform(-1, 1), uniform(-1, 1)); } while (dot(d, d) < 1e-4f || dot(d, d) > 1.0f);
    if (kind == 1) d.x = 0.0f;                      // one zero component
    if (kind == 2) { d.y = 0.0f; d.z = -0.0f; }     // axis parallel
    if (kind == 3) d.z *= 1e-7f;                    // one tiny component
    if (dot(d, d) == 0.0f) d = f3(0, 1, 0);
    return normalize(d);
}

// Exact slab test in double on the REAL child box.
static bool exact_box_hit(const Box& b, float3 o, float3 d, float tmin, float tmax) {
    double t0 = tmin, t1 = tmax;
    const double O[3] = { o.x, o.y, o.z }, D[3] = { d.x, d.y, d.z }, L[3] = { b.lo.x, b.lo.y, b.lo.z }, H[3] = { b.hi.x, b.hi.y, b.hi.z };
    for (int a = 0; a < 3; ++a) {
        if (D[a] == 0.0) { if (O[a] < L[a] || O[a] > H[a]) return false; continue; }
        double ta = (L[a] - O[a]) / D[a], tb = (H[a] - O[a]) / D[a];
        if (ta > tb) std::swap(ta, tb);
        t0 = std::max(t0, ta); t1 = std::min(t1, tb);
    }
    return t0 <= t1;
}

static void test_node_conservative() {
    long reported = 0, exact = 0;
    for (int trial = 0; trial < 20000; ++trial) {
        const int count = 2 + (int)(rng() % 7);
        const float scale = std::pow(10.0f, uniform(-4, 4));
        const float3 centre = f3(uniform(-1, 1), uniform(-1, 1), uniform(-1, 1)) * (scale * (trial % 3 == 0 ? 1000.0f : 1.0f));
        float3 lo[8], hi[8]; int triangles[8];
        for (int i = 0; i < count; ++i) {
            float3 c = centre + f3(uniform(-1, 1), uniform(-1, 1), uniform(-1, 1)) * scale;
            float3 h = f3(uniform(0, 0.5f), uniform(0, 0.5f), uniform(0, 0.5f)) * scale;
            if (trial % 5 == 0) h.y = 0.0f;                             // flat boxes
            if (trial % 7 == 0) { c.z = centre.z; h.z = 0.0f; }         // the whole node is flat
            lo[i] = c - h; hi[i] = c + h;
            triangles[i] = (int)(rng() % 4); // 0 = inner
        }
        CwNode node; CwPlacement place;
        cw_encode(count, lo, hi, triangles, node, place);
        node.child_base = 100; node.triangle_base = 1000;
        CHECK(place.triangle_count <= 24, "triangle_count %d", place.triangle_count);
        uint32_t slots_used = 0;
        for (int i = 0; i < count; ++i) { CHECK(!(slots_used >> place.slot[i] & 1u), "slot reused"); slots_used |= 1u << place.slot[i]; }
        const uint4 n0 = word(node, 0), n1 = word(node, 1), n2 = word(node, 2), n3 = word(node, 3), n4 = word(node, 4);
        for (int r = 0; r < 16; ++r) {
            float3 o = centre + f3(uniform(-3, 3), uniform(-3, 3), uniform(-3, 3)) * scale;
            if (r % 4 == 0) o = lo[rng() % count];                    // origin on a box corner
            const float3 d = random_direction(r % 4);
            const float tmin = r % 3 == 0 ? 0.0f : uniform(0, scale), tmax = r % 2 == 0 ? 1e30f : uniform(scale, 6 * scale);
            const CwRay ray = cw_make_ray(o, d);
            const uint32_t hits = cw_intersect_children(n0, n1, n2, n3, n4, ray, tmin, tmax);
            uint32_t allowed = 0;
            for (int i = 0; i < count; ++i) {
                uint32_t bits;
                if (triangles[i] == 0) bits = 1u << (24 + (place.slot[i] ^ (int)(ray.oct_inv4 & 7u)));
                else bits = ((1u << triangles[i]) - 1u) << place.offset[i];
                CHECK(!(allowed & bits), "children share hit bits");
                allowed |= bits;
                const bool must = exact_box_hit({ lo[i], hi[i] }, o, d, tmin, tmax);
                exact += must; reported += (hits & bits) == bits;
                CHECK(!must || (hits & bits) == bits, "culled a child the exact test passes (trial %d ray %d child %d)", trial, r, i);
                CHECK((hits & bits) == 0 || (hits & bits) == bits, "partial bits");
                // node index of an inner child
                if (triangles[i] == 0) {
                    uint2 g = make_uint2(node.child_base, (1u << (24 + (place.slot[i] ^ (int)(ray.oct_inv4 & 7u)))) | (n0.w >> 24));
                    CHECK(cw_next_child(g, ray) == node.child_base + (uint32_t)place.offset[i], "inner child index");
                }
            }
            CHECK((hits & ~allowed) == 0, "bits of no child: %08x", hits & ~allowed);
        }
    }
    printf("node test: %ld child boxes pass exactly, %ld reported (%.2f %% more)\n", exact, reported, 100.0 * (reported - exact) / std::max(1L, exact));
}

static std::vector<Tri> make_scene(int kind, int n) {
    std::vector<Tri> tris(n);
    for (int i = 0; i < n; ++i) {
        float3 c; float size = 0.05f;
        if (kind == 0) c = f3(uniform(-1, 1), uniform(-1, 1), uniform(-1, 1));
        else if (kind == 1) { c = f3(uniform(-1, 1), 0.0f, uniform(-1, 1)); size = 0.03f; }                   // near-planar
        else if (kind == 2) { float3 k = f3(float(rng() % 4), float(rng() % 4), float(rng() % 4)); c = k * 10.0f + f3(uniform(-0.2f, 0.2f), uniform(-0.2f, 0.2f), uniform(-0.2f, 0.2f)); size = 0.02f; } // clusters
        else { c = f3(5000.0f, -3000.0f, 8000.0f) + f3(uniform(-1, 1), uniform(-1, 1), uniform(-1, 1)); }       // far from the origin
        for (int v = 0; v < 3; ++v) tris[i].v[v] = c + f3(uniform(-size, size), kind == 1 ? uniform(-1e-3f, 1e-3f) : uniform(-size, size), uniform(-size, size));
        if (kind == 1 && i % 10 == 0) for (int v = 0; v < 3; ++v) tris[i].v[v].y = 0.0f;                         // exactly flat, axis aligned
        tris[i].id = i;
    }
    return tris;
}

static void test_tree(int kind, int n, int rays) {
    std::vector<Tri> tris = make_scene(kind, n);
    std::vector<Tri> sorted = tris;
    std::vector<BinNode> bin;
    build_binary(bin, sorted, 0, n, 1 + kind % 3); // leaves of up to 1, 2, 3 triangles
    CwTree tree; collapse(bin, sorted, tree);
    CHECK(tree.tris.size() == tris.size(), "triangle count %zu", tree.tris.size());
    float3 lo = bin[0].box.lo, hi = bin[0].box.hi, centre = (lo + hi) * 0.5f, half = (hi - lo) * 0.5f;
    long hits = 0, nodes = 0, tests = 0; size_t max_stack = 0;
    for (int r = 0; r < rays; ++r) {
        float3 o = centre + f3(uniform(-1.5f, 1.5f) * half.x, uniform(-1.5f, 1.5f) * half.y, uniform(-1.5f, 1.5f) * half.z);
        float3 d = random_direction(r % 8 < 4 ? 0 : r % 4);
        if (r % 5 == 0) { const Tri& t = tris[rng() % n]; o = (t.v[0] + t.v[1] + t.v[2]) * (1.0f / 3.0f); }   // starts on a surface
        if (r % 7 == 0) { const Tri& t = tris[rng() % n]; d = normalize(t.v[rng() % 3] - o); }                   // aims at a vertex
        const float tmin = r % 3 == 0 ? 0.0f : 1e-4f;
        Result a = brute(tris, o, d, tmin, 1e30f), b = traverse(tree, o, d, tmin, 1e30f, &max_stack);
        CHECK(a.t == b.t && a.id == b.id, "scene %d ray %d: brute (%g, %d) traversal (%g, %d)", kind, r, a.t, a.id, b.t, b.id);
        hits += a.id != 0x7fffffff; nodes += b.nodes; tests += b.tests;
    }
    printf("scene %d: %d triangles, %zu nodes (%.2f per triangle), %d levels, %ld of %d rays hit, %.1f node visits and %.1f triangle tests per ray, deepest stack %zu\n",
           kind, n, tree.nodes.size(), double(tree.nodes.size()) / n, tree.levels, hits, rays, double(nodes) / rays, double(tests) / rays, max_stack);
}

int main() {
    test_node_conservative();
    test_tree(0, 20000, 3000);
    test_tree(1, 20000, 3000);
    test_tree(2, 20000, 3000);
    test_tree(3, 20000, 3000);
    test_tree(0, 2, 200);
    test_tree(2, 9, 200);
    if (failures) { printf("%d FAILURES\n", failures); return 1; }
    printf("OK\n");
    return 0;
}
