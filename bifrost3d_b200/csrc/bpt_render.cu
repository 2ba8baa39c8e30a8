// Wavefront path integrator. Replaces the OptiX megakernel launched by Renderer::render
// (Renderer.cpp:1250-1265 -> context->launch(PathTracing, w, h), IBackend.h:37-39) and its programs:
//   path_tracing_RPG / accumulate / initialize_monte_carlo_payload   Shading/SimpleRGPs.cu:44-140
//   miss                                                              Shading/SimpleRGPs.cu:349-362
//   path_tracing_closest_hit<DefaultMaterialCreator>, NEE + RIS       Shading/MonteCarlo.cu:61-244
//   shadow_any_hit, light_closest_hit                                 Shading/MonteCarlo.cu:278-302
//   interpolate_attributes                                            Shading/TriangleAttributes.cu:35-84
//   analytic light intersect                                          Shading/LightSources/LightSources.cu:31-70
//
// One progressive sample of every pixel =
//     generate -> { extend(k) || shadow(k - 1) -> shade(k) -> advance } while paths are alive -> shadow(last) -> accumulate.
// Every stage is a persistent-thread kernel over a queue of pixel indices whose length lives in device memory. The whole
// sample is ONE CUDA graph whose loop is a conditional WHILE node: the device decides when the queues have drained, the host
// never synchronises inside bpt_render, and the shadow rays of bounce k - 1 are traced concurrently with the closest-hit
// rays of bounce k (both only depend on shade(k - 1)). Queues are compacted with warp ballot + one atomic per warp.
// Path state is SoA, indexed by pixel, in 16-byte records so every access is a 128-bit transaction.
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#include "bpt_context.h"
#include "bpt_lights.cuh"
#include "bpt_rng.cuh"
#include "bpt_trace.cuh"

#include <cuda_fp16.h>
#include <algorithm>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

#ifndef BPT_DEFAULT_LANES
#define BPT_DEFAULT_LANES 4 // samples in flight (BPT_LANES in the environment overrides; 1 = a single chain of kernels)
#endif
#ifndef BPT_TRACE_GRID_CTAS_SHARED
#define BPT_TRACE_GRID_CTAS_SHARED 4 // CTAs per SM of a traversal kernel's persistent grid when several samples are in flight
#endif
#ifndef BPT_TILED_QUEUE
#define BPT_TILED_QUEUE 1
#endif

namespace bpt {

namespace {

#ifndef BPT_SHADE_BLOCK
#define BPT_SHADE_BLOCK 512
#endif
constexpr int SHADE_BLOCK = BPT_SHADE_BLOCK; // threads per CTA of the shade kernels (with BPT_SHADE_MIN_BLOCKS CTAs per SM)
constexpr int LIGHT_HIT_FLAG = 0x40000000; // hit.primitive = LIGHT_HIT_FLAG | light index
constexpr float RT_DEFAULT_MAX = 1e27f;    // tmax of optix::Ray when none is given (SimpleRGPs.cu:114)

// Queue counters in device memory. Queues ping-pong: in iteration k the extend / shade kernels read queue[parity] and append
// the continuing paths to queue[parity ^ 1]; shade(k) appends its shadow rays under shadow[parity] and the shadow kernel that
// runs beside extend(k + 1) (parity flipped by then) reads shadow[parity ^ 1].
struct QueueCounters {
    unsigned int active;       // entries in the current extend/shade queue
    unsigned int next_active;  // entries appended for the next iteration
    unsigned int shadow[2];    // entries in the shadow queue, by the parity of the iteration that filled it
    unsigned int fetch_extend; // dynamic ray fetch cursors of the two traversal kernels
    unsigned int fetch_shadow;
    unsigned int surface;      // paths whose ray hit a Default / Diffuse surface (shade_kernel<true, false>)
    unsigned int escaped;      // paths whose ray left the scene or hit an analytic light (shade_kernel<false, false>)
    unsigned int transmissive; // paths whose ray hit a Transmissive surface (shade_kernel<true, true>)
    unsigned int parity;       // 0 / 1, flipped by advance_kernel
    unsigned int iteration;    // iterations of the current sample (guards against a queue that never drains)
};

// What changes from frame to frame lives in device memory, so that one instantiated graph serves every bpt_render call of a
// scene: the host copies this small record in front of the launches, finish_sample_kernel advances sample_index.
struct FrameState {
    bpt_camera camera;
    float path_regularization_pdf_scale;
    unsigned int sample_index; // accumulation index of the sample being rendered (Types.h:486-501 `accumulations`)
    unsigned int sample_stride; // what finish_sample_kernel adds to it: the number of lanes (samples in flight)
};

struct Wavefront {
    int64_t pixel_capacity = 0;
    // ray_o: origin.xyz, tmin | ray_d: direction.xyz, bsdf_pdf | thr: throughput.xyz, bounces (bits) |
    // rad: radiance.xyz, previous primitive (bits) | hit: t, primitive (bits), u, v
    DeviceBuffer<float4> ray_o, ray_d, thr, rad, hit;
    // shadow rays: origin.xyz + tmax, direction.xyz + pixel (bits), radiance.xyz
    DeviceBuffer<float4> sh_o, sh_d, sh_rad;
    DeviceBuffer<unsigned int> queue_a, queue_b;
    DeviceBuffer<unsigned int> queue_surface, queue_escaped; // extend sorts its results by what shading they need
    // Sorting of the surface hits by (shading class, hit cell): keys beside queue_surface, the sorted queue, 2 x SORT_BINS
    // counters (histogram, then running cursors) and the per-material shading class.
    DeviceBuffer<unsigned short> surface_key;
    DeviceBuffer<unsigned int> queue_surface_sorted, sort_bins;
    DeviceBuffer<unsigned char> material_class;
    DeviceBuffer<QueueCounters> counters;
    DeviceBuffer<FrameState> frame_state;
    DeviceBuffer<float> coverage; // per material
    uint64_t coverage_version = ~0ull;

    // The sample graph and what it was built for (rebuilt when any launch parameter changes).
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    std::vector<unsigned char> graph_signature;
    bool graph_unavailable = false; // conditional graph nodes could not be created: stream launches + host polling instead
};

struct WavefrontView {
    float4 *ray_o, *ray_d, *thr, *rad, *hit;
    float4 *sh_o, *sh_d, *sh_rad;
    unsigned int *queue_a, *queue_b; // selected by parity with a conditional, never indexed: a dynamically indexed kernel
                                     // parameter would make the compiler copy the whole struct to local memory
    unsigned int *queue_surface, *queue_escaped;
    unsigned int queue_capacity;      // entries per queue; the transmissive queue grows down from the end of queue_surface
    unsigned short* surface_key;      // sort key of queue_surface[i]
    unsigned int* queue_surface_sorted;
    unsigned int* sort_bins;          // [0, SORT_BINS): histogram of the keys; [SORT_BINS, 2 SORT_BINS): scatter cursors
    QueueCounters* counters;
    const FrameState* frame;
    unsigned long long* ray_counters; // [0] extend, [1] shadow, [6] dropped non-finite samples, [7] iterations
};

struct SceneView {
    AccelView accel;
    const float4* __restrict__ world_vertices;
    const ShadeTriangle* __restrict__ shade;
    const float* __restrict__ normal_matrices;
    const float* __restrict__ shade_emission; // 9 floats per primitive or nullptr (scale 1)
    const Material* __restrict__ materials;
    const float* __restrict__ coverage;
    const Light* __restrict__ lights;
    int light_count;          // lights sampled by next event estimation (analytic + environment when importance sampled)
    int analytic_light_count; // lights that rays can hit
    EnvironmentView env;
    const float* __restrict__ tables;
    const float2* __restrict__ dielectric_tables;
    const float4* __restrict__ nee_offsets;
    bool split_by_shading_model; // the scene has Transmissive materials: extend keys surface hits by shading model
    // Sorting of the surface hits before shading (north star (3): material-keyed sorting before shading)
    const uint32_t* __restrict__ slot_of_primitive; // position of a primitive in the Morton-ordered triangle array
    const unsigned char* __restrict__ material_class;
    int sort_cell_shift;                          // hit cell = Morton slot >> shift, below SORT_CELLS
};

// Per-configuration constants (kernel parameters, baked into the graph).
struct FrameParams {
    int width, height;
    unsigned int max_bounce_count;
    int next_event_sample_count;
    unsigned int russian_roulette_start_bounce; // 0 = off (the reference has no Russian roulette)
    unsigned int sort_hits_from_iteration;      // surface hits are sorted from this iteration of a sample on; 0xffffffff = never
};

// Key of a surface hit: [shading class : 2][hit cell : 10]. The class separates what makes the shading code branch (Diffuse
// versus Default shading model, coat or none); the cell is the top of the hit triangle's position in the Morton-ordered triangle
// array, i.e. WHERE the path is. Sorting by it gives the shade kernel warps that read neighbouring triangles and materials
// and take the same branches - and, because shading appends its rays in queue order, gives the next closest-hit launch and
// the shadow launch warps whose rays start next to each other.
constexpr unsigned int SORT_CELLS = 1024, SORT_CLASSES = 4, SORT_BINS = SORT_CELLS * SORT_CLASSES;
__device__ __forceinline__ bool sorting_now(const QueueCounters* c, unsigned int from_iteration) { return c->iteration >= from_iteration; }


// Appends `value` to a queue for every lane with `pred` set: one atomicAdd per warp.
__device__ __forceinline__ void warp_append(bool pred, unsigned int* queue, unsigned int* counter, unsigned int value) {
    unsigned int mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

// ---- generate ------------------------------------------------------------------------------------

__device__ __forceinline__ float4 mul4x4(const float* m, float4 v) {
    return make_float4(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w, m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w,
                       m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w, m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w);
}

__global__ void generate_kernel(WavefrontView w, FrameParams f) {
    int64_t pixel_count = (int64_t)f.width * f.height;
    const bpt_camera& camera = w.frame->camera;
    const unsigned int accumulation_count = w.frame->sample_index;
    unsigned int* __restrict__ queue_in = w.queue_a; // a sample starts with parity 0
    // Queue order: 8 x 4 pixel tiles, one per warp, so that the camera rays (and the first hits) of a warp are neighbours in
    // both image directions; plain row order when the frame is not a whole number of tiles. Path state stays indexed by pixel.
    const bool tiled = BPT_TILED_QUEUE && (f.width % 8 == 0) && (f.height % 4 == 0);
    const int tiles_x = f.width / 8;
    for (int64_t slot = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; slot < pixel_count; slot += (int64_t)gridDim.x * blockDim.x) {
        int x, y;
        if (tiled) {
            const int64_t tile = slot >> 5; const int lane = int(slot & 31);
            x = int(tile % tiles_x) * 8 + (lane & 7); y = int(tile / tiles_x) * 4 + (lane >> 3);
        } else { x = int(slot % f.width); y = int(slot / f.width); }
        const int64_t p = (int64_t)y * f.width + x;
        unsigned int pixel_hash = pcg2d((unsigned int)x, (unsigned int)y).x;
        float2 jitter = f2(0.5f, 0.5f);
        if (accumulation_count != 0) {
            float4 r = path_rng_sample4f(accumulation_count, pixel_hash, 0u, DIM_CAMERA);
            jitter = f2(r.x, r.y);
        }
        float2 screen_pos = f2(float(x) + jitter.x, float(y) + jitter.y);
        float2 viewport_pos = f2(screen_pos.x / float(f.width), screen_pos.y / float(f.height));

        float4 ndc_near = make_float4(viewport_pos.x * 2.0f - 1.0f, viewport_pos.y * 2.0f - 1.0f, -1.0f, 1.0f);
        float4 near_world = mul4x4(camera.inverse_view_projection, ndc_near);
        float3 origin = f3(near_world) / near_world.w;
        float4 ndc_far = make_float4(ndc_near.x, ndc_near.y, 1.0f, 1.0f);
        float4 far_view = mul4x4(camera.inverse_projection, ndc_far);
        const float* r = camera.view_to_world_rotation;
        float3 v = f3(far_view);
        float3 direction = normalize(f3(r[0] * v.x + r[1] * v.y + r[2] * v.z, r[3] * v.x + r[4] * v.y + r[5] * v.z, r[6] * v.x + r[7] * v.y + r[8] * v.z));

        w.ray_o[p] = f4(origin, 0.0f);
        w.ray_d[p] = f4(direction, Pdf::delta_dirac(1.0f).v);
        w.thr[p] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(0u));
        w.rad[p] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
        queue_in[slot] = (unsigned int)p;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        QueueCounters c = {};
        c.active = (unsigned int)pixel_count;
        *w.counters = c;
    }
}

// ---- extend: closest hit over triangles and analytic lights -------------------------------------------------

template <bool SORT_HITS>
struct ExtendSource {
    WavefrontView w;
    const unsigned int* __restrict__ queue_in;
    const Light* __restrict__ lights;
    int analytic_light_count;
    // Set when the scene has Transmissive materials: surface hits are then keyed by shading model.
    const ShadeTriangle* __restrict__ shade;
    const Material* __restrict__ materials;
    // Read when the surface hits of this iteration are sorted before shading (SORT_HITS instantiation only).
    bool sort_hits;
    const ShadeTriangle* __restrict__ shade_all;
    const uint32_t* __restrict__ slot_of_primitive;
    const unsigned char* __restrict__ material_class;
    int sort_cell_shift;
    __device__ void load(unsigned int i, Ray& ray, int& skip) const {
        unsigned int pixel = queue_in[i];
        float4 o = w.ray_o[pixel], d = w.ray_d[pixel];
        ray.origin = f3(o); ray.tmin = o.w; ray.direction = f3(d); ray.tmax = RT_DEFAULT_MAX;
        // the primitive the path is leaving (MonteCarlo.cu:137-142). Only the w lane is read: the shadow kernel that runs
        // concurrently adds to the xyz lanes of the same record.
        skip = __float_as_int(reinterpret_cast<const float*>(w.rad + pixel)[3]);
    }
    __device__ float termination_weight(unsigned int) const { return 1.0f; }
    template <class Trav>
    __device__ void store(unsigned int i, const Trav& tr) const {
        Hit h = tr.result();
        float t_closest = h.primitive >= 0 ? h.t : RT_DEFAULT_MAX;
        const unsigned int pixel = queue_in[i];
        // The eight-wide traversal does not keep the direction (it carries the reciprocal): read the ray again when a light needs it.
        float3 ray_origin, ray_direction; float ray_tmin;
        if constexpr (Trav::COMPRESSED) {
            if (analytic_light_count > 0) { const float4 o = w.ray_o[pixel], d = w.ray_d[pixel]; ray_origin = f3(o); ray_tmin = o.w; ray_direction = f3(d); }
        } else { ray_origin = tr.ray.origin; ray_direction = tr.ray.direction; ray_tmin = tr.ray.tmin; }
        // Analytic sphere / disk lights (LightSources.cu:31-70): intersectable by MonteCarlo rays only.
        for (int l = 0; l < analytic_light_count; ++l) {
            Light light = lights[l];
            float t = -1e30f, radius = 0.0f;
            if (light_type(light) == BPT_LIGHT_SPHERE) {
                SphereLight sl = as_sphere(light); radius = sl.radius;
                t = isect::ray_sphere(ray_origin, ray_direction, sl.position, sl.radius);
            } else if (light_type(light) == BPT_LIGHT_SPOT) {
                SpotLight sp = as_spot(light); radius = sp.radius;
                t = isect::ray_disk(ray_origin, ray_direction, sp.position, sp.direction, sp.radius);
            }
            if (radius > 0.0f && t > ray_tmin && t < t_closest) { t_closest = t; h.t = t; h.primitive = LIGHT_HIT_FLAG | l; }
        }
        w.hit[pixel] = make_float4(h.t, __int_as_float(h.primitive), h.u, h.v);
        // Sort the paths by the shading they need: surface hits go to the (large) surface shading kernels - keyed by the
        // material's shading model when the scene mixes them - and escaped rays and light hits to a small one, so none runs
        // with lanes masked off for another's work. Same-address atomics of a warp are aggregated by the compiler
        // (REDUX + one atomic).
        if (h.primitive >= 0 && !(h.primitive & LIGHT_HIT_FLAG)) {
            bool transmissive = shade != nullptr && materials[shade[h.primitive].material_index].shading_model == SHADING_TRANSMISSIVE;
            if (transmissive) w.queue_surface[w.queue_capacity - 1u - atomicAdd(&w.counters->transmissive, 1u)] = pixel;
            else {
                const unsigned int slot = atomicAdd(&w.counters->surface, 1u);
                w.queue_surface[slot] = pixel;
                if (SORT_HITS && sort_hits) {
                    const unsigned int cell = min(slot_of_primitive[h.primitive] >> sort_cell_shift, SORT_CELLS - 1u);
                    const unsigned int key = (unsigned int)material_class[shade_all[h.primitive].material_index] * SORT_CELLS + cell;
                    w.surface_key[slot] = (unsigned short)key;
                    atomicAdd(w.sort_bins + key, 1u);
                }
            }
        } else
            w.queue_escaped[atomicAdd(&w.counters->escaped, 1u)] = pixel;
#ifdef BPT_TRAVERSAL_STATS
        atomicAdd(w.ray_counters + 2, (unsigned long long)tr.stat_nodes);
        atomicAdd(w.ray_counters + 3, (unsigned long long)tr.stat_triangles);
#endif
    }
};

// SORT_HITS: the instantiation that also writes the sort keys of the surface hits (bpt_set_hit_sorting); the default one
// carries none of that state.
// COMPRESSED: the instantiation that traverses the compressed eight-wide nodes (s.accel.cw).
template <bool SORT_HITS, bool COMPRESSED>
__global__ void __launch_bounds__(TRACE_BLOCK, BPT_TRACE_MIN_BLOCKS) extend_kernel(WavefrontView w, SceneView s, FrameParams f) {
    __shared__ __align__(16) int s_stack[STACK_SMEM * TRACE_BLOCK];
    const unsigned int count = w.counters->active;
    ExtendSource<SORT_HITS> source = { w, w.counters->parity ? w.queue_b : w.queue_a, s.lights, s.analytic_light_count, s.split_by_shading_model ? s.shade : nullptr, s.materials,
                            SORT_HITS && sorting_now(w.counters, f.sort_hits_from_iteration), s.shade, s.slot_of_primitive, s.material_class, s.sort_cell_shift };
    typedef typename TraversalFor<false, COMPRESSED>::type Trav;
    traverse_queue_with<false, Trav>(s.accel, s.coverage, source, count, &w.counters->fetch_extend, s_stack + threadIdx.x, s.accel.budget);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(w.ray_counters, (unsigned long long)count);
}

// ---- shadow: accumulated transmission along the light sample's segment ---------------------------------------

struct ShadowSource {
    WavefrontView w;
    __device__ void load(unsigned int i, Ray& ray, int& skip) const {
        float4 o = w.sh_o[i], d = w.sh_d[i];
        ray.origin = f3(o); ray.tmin = 0.0f; ray.direction = f3(d); ray.tmax = o.w;
        skip = -1;
    }
    __device__ float termination_weight(unsigned int i) const { return w.sh_rad[i].w; } // max(r, g, b), stored by the shade kernel
    template <class Trav>
    __device__ void store(unsigned int i, const Trav& tr) const {
        if (tr.transmission > 0.0f) {
            unsigned int pixel = __float_as_uint(w.sh_d[i].w);
            float4 l = w.sh_rad[i];
            // One shadow ray per pixel and iteration, so the read-modify-write needs no atomic; the w lane (previous primitive)
            // is left alone because the concurrent extend kernel reads it.
            float* rad = reinterpret_cast<float*>(w.rad + pixel);
            rad[0] += l.x * tr.transmission; rad[1] += l.y * tr.transmission; rad[2] += l.z * tr.transmission;
        }
    }
};

template <bool COMPRESSED>
__global__ void __launch_bounds__(TRACE_BLOCK, BPT_TRACE_MIN_BLOCKS) shadow_kernel(WavefrontView w, SceneView s) {
    __shared__ __align__(16) int s_stack[STACK_SMEM * TRACE_BLOCK];
    const unsigned int count = w.counters->shadow[w.counters->parity ^ 1u]; // filled by the previous iteration's shade kernels
    ShadowSource source = { w };
    typedef typename TraversalFor<true, COMPRESSED>::type Trav;
    traverse_queue_with<true, Trav>(s.accel, s.coverage, source, count, &w.counters->fetch_shadow, s_stack + threadIdx.x, s.accel.budget);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(w.ray_counters + 1, (unsigned long long)count);
}

// Counting sort of the surface queue by key, two small kernels between the traversal and the shading of an iteration. The
// histogram was filled by extend_kernel's atomics; one CTA turns it into running cursors (and clears it for the next
// iteration), then every hit takes the next free position of its bin. The order inside a bin is whatever the atomics
// make it - results do not depend on it, path state is indexed by pixel.
__global__ void __launch_bounds__(1024) sort_scan_kernel(WavefrontView w, FrameParams f) {
    if (!sorting_now(w.counters, f.sort_hits_from_iteration)) return;
    __shared__ unsigned int warp_sums[32];
    constexpr unsigned int PER_THREAD = SORT_BINS / 1024;
    unsigned int v[PER_THREAD], sum = 0;
#pragma unroll
    for (unsigned int k = 0; k < PER_THREAD; ++k) { v[k] = w.sort_bins[threadIdx.x * PER_THREAD + k]; sum += v[k]; w.sort_bins[threadIdx.x * PER_THREAD + k] = 0; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int inclusive = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, inclusive, o); if (lane >= o) inclusive += t; }
    if (lane == 31) warp_sums[warp] = inclusive;
    __syncthreads();
    if (warp == 0) {
        unsigned int t2 = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned int t = __shfl_up_sync(0xffffffffu, t2, o); if (lane >= o) t2 += t; }
        warp_sums[lane] = t2;
    }
    __syncthreads();
    unsigned int running = (warp ? warp_sums[warp - 1] : 0u) + inclusive - sum;
#pragma unroll
    for (unsigned int k = 0; k < PER_THREAD; ++k) { w.sort_bins[SORT_BINS + threadIdx.x * PER_THREAD + k] = running; running += v[k]; }
}

__global__ void __launch_bounds__(256) sort_scatter_kernel(WavefrontView w, FrameParams f) {
    if (!sorting_now(w.counters, f.sort_hits_from_iteration)) return;
    const unsigned int count = w.counters->surface;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        w.queue_surface_sorted[atomicAdd(w.sort_bins + SORT_BINS + w.surface_key[i], 1u)] = w.queue_surface[i];
}

// Ends an iteration: the appended paths become the active queue, the queue roles flip, and the graph's WHILE node is told
// whether another iteration is needed (a null handle: the host polls `active` instead).
constexpr unsigned int MAX_ITERATIONS_PER_SAMPLE = 4096; // rejected hits re-trace without consuming a bounce; this is only a backstop
__global__ void advance_kernel(QueueCounters* c, unsigned long long* ray_counters, cudaGraphConditionalHandle loop_handle, int has_handle) {
    const unsigned int parity = c->parity;
    c->active = c->next_active;
    c->next_active = 0;
    c->shadow[parity ^ 1u] = 0; // consumed by this iteration's shadow kernel; the next iteration's shade kernels fill it
    c->fetch_extend = 0;
    c->fetch_shadow = 0;
    c->surface = 0;
    c->escaped = 0;
    c->transmissive = 0;
    c->parity = parity ^ 1u;
    c->iteration += 1u;
    atomicAdd(ray_counters + 7, 1ull); // several samples may be in flight
    if (c->iteration >= MAX_ITERATIONS_PER_SAMPLE) c->active = 0; // the leftover paths are dropped; never observed
    if (has_handle) cudaGraphSetConditional(loop_handle, c->active != 0u ? 1u : 0u);
}

// ---- shade ---------------------------------------------------------------------------------------------

// Utils.h:67-74
__device__ __forceinline__ float3 fix_backfacing_shading_normal(float3 wdir, float3 n, float target_cos_theta) {
    float cos_theta = dot(wdir, n);
    if (cos_theta < target_cos_theta) {
        float c = cos_theta - target_cos_theta;
        return normalize(n - c * wdir);
    }
    return n;
}

// Utils.h:372-397 (Ray Tracing Gems ch. 6)
__device__ __forceinline__ float3 offset_ray_origin(float3 p, float3 n) {
    const float origin = 1.0f / 32.0f;
    const float float_scale = 1.0f / 65536.0f;
    const float int_scale = 256.0f;
    int3 of_i = make_int3(int(int_scale * n.x), int(int_scale * n.y), int(int_scale * n.z));
    float3 p_i = f3(__int_as_float(__float_as_int(p.x) + ((p.x < 0) ? -of_i.x : of_i.x)),
                    __int_as_float(__float_as_int(p.y) + ((p.y < 0) ? -of_i.y : of_i.y)),
                    __int_as_float(__float_as_int(p.z) + ((p.z < 0) ? -of_i.z : of_i.z)));
    return f3(fabsf(p.x) < origin ? p.x + float_scale * n.x : p_i.x,
              fabsf(p.y) < origin ? p.y + float_scale * n.y : p_i.y,
              fabsf(p.z) < origin ? p.z + float_scale * n.z : p_i.z);
}
__device__ __forceinline__ float3 offset_ray_origin(float3 p, float3 direction, float3 geometric_normal) {
    float cos_theta = dot(geometric_normal, direction);
    geometric_normal = cos_theta >= 0 ? geometric_normal : -geometric_normal;
    return offset_ray_origin(p, geometric_normal);
}

// OctahedralNormal::decode, Types.h:62-69
__device__ __forceinline__ float3 oct_decode(const int16_t e[2]) {
    float2 fe = f2(float(e[0]), float(e[1]));
    float3 n = f3(fe.x, fe.y, 32767 - fabsf(fe.x) - fabsf(fe.y));
    float t = fmaxf(-n.z, 0.0f);
    n.x += n.x >= 0 ? -t : t;
    n.y += n.y >= 0 ? -t : t;
    return normalize(n);
}

// Utils.h:286-290
__device__ __forceinline__ unsigned char unorm8(float v) { return (unsigned char)(saturate(v) * 255.0f + 0.5f); }

// sample_single_light, MonteCarlo.cu:61-87
template <typename Bsdf>
__device__ LightSample sample_single_light(const SceneView& s, const Bsdf& material, float3 position, float3 wo, const Tbn& tbn, float3 u) {
    int light_index = min(s.light_count - 1, int(u.z * s.light_count));
    Light light = s.lights[light_index];
    LightSample ls = light_sample_radiance(light, s.env, position, f2(u.x, u.y));
    ls.radiance *= float(s.light_count);

    float N_dot_L = dot(tbn.normal, ls.direction_to_light);
    ls.radiance *= fdiv(fabsf(N_dot_L), ls.pdf.value());

    const float3 shading_light_direction = tbn.to_local(ls.direction_to_light);
    BsdfResponse response = material.evaluate_with_pdf(wo, shading_light_direction);
    bool apply_MIS = !ls.pdf.is_delta_dirac();
    if (apply_MIS)
        ls.radiance *= mis_weight(ls.pdf, response.pdf);
    else
        response.reflectance = min3(response.reflectance, f3(32.0f));
    ls.radiance *= response.reflectance;
    return ls;
}

#ifndef BPT_SHADE_MIN_BLOCKS
#define BPT_SHADE_MIN_BLOCKS 2
#endif
#ifndef BPT_SHADE_SYNC
#define BPT_SHADE_SYNC 2 // 0: no barriers; 1: one at the top of every path; 2: also between the phases of the surface shading
#endif
#if BPT_SHADE_SYNC >= 1
#define SHADE_TOP_BARRIER() __syncthreads()
#else
#define SHADE_TOP_BARRIER() do { } while (0)
#endif
#if BPT_SHADE_SYNC >= 2
#define SHADE_PHASE_BARRIER() __syncthreads()
#else
#define SHADE_PHASE_BARRIER() do { } while (0)
#endif
// SURFACE, !TRANSMISSIVE: default_closest_hit / diffuse_closest_hit (MonteCarlo.cu:246-257); SURFACE, TRANSMISSIVE:
// transmissive_closest_hit (:259-268), launched only for scenes that hold such materials; !SURFACE: miss and light hits.
template <bool SURFACE, bool TRANSMISSIVE>
__global__ void __launch_bounds__(SHADE_BLOCK, BPT_SHADE_MIN_BLOCKS) shade_kernel(WavefrontView w, SceneView s, FrameParams f) {
    // The 12 KB of rho / alpha tables are read through L1 (they stay hot) rather than staged in shared memory: the kernel's
    // spills and call frames (~0.5 KB per thread) need the L1 capacity more. Measured: shade -5 %.
    const ShadingTables tables = { s.tables, s.tables + TABLE_FLOATS, s.tables + 2 * TABLE_FLOATS };

    const unsigned int accumulation_count = w.frame->sample_index;
    const float path_regularization_pdf_scale = w.frame->path_regularization_pdf_scale;
    const unsigned int parity = w.counters->parity;
    unsigned int* __restrict__ queue_out = parity ? w.queue_a : w.queue_b;
    unsigned int* shadow_counter = w.counters->shadow + parity;
    const unsigned int count = SURFACE ? (TRANSMISSIVE ? w.counters->transmissive : w.counters->surface) : w.counters->escaped;
    const unsigned int* __restrict__ queue = SURFACE ? (TRANSMISSIVE ? w.queue_surface + (w.queue_capacity - count)
                                                                     : (sorting_now(w.counters, f.sort_hits_from_iteration) ? w.queue_surface_sorted : w.queue_surface))
                                                     : w.queue_escaped;
    // whole CTAs take part in every iteration (the barriers below), whole warps in the ballots of the queue compaction
    const unsigned int rounded = ((count + SHADE_BLOCK - 1u) / SHADE_BLOCK) * SHADE_BLOCK;
    // The surface shading of one path is ~4 500 executed instructions of straight-line code out of a ~100 KB kernel: every warp
    // streams the whole kernel through the instruction caches once per path, and with the warps of an SM at unrelated places
    // the top stall reason is instruction fetch (ncu round 2: no_instruction 5.7 warps per issue, 50 % of the issue slots).
    // So the warps of a CTA are kept side by side: a CTA-wide barrier at the top of every path and between the phases of the
    // shading (set-up | each next-event candidate | BSDF sampling) lets one fetched line serve all of them. Measured on B200:
    // barrier at the top only, 512-thread CTAs: shade -7 % (1 M triangle scene) / -17 % (Cornell box); see DESIGN.md 6 for the
    // phase barriers. The arithmetic and its order per path are untouched.
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
        SHADE_TOP_BARRIER();
        // Barriers must be reached by every thread of the CTA, and hoisting the shading state out of a conditional costs
        // registers (measured: +6 % on the kernel). So the shading below is unconditional straight-line code: a thread past
        // the end of the queue shades the queue's last path once more, a thread whose hit is rejected (back face, coverage)
        // shades it anyway, and both drop the result. Lanes that would otherwise sit masked do the work, so it costs nothing
        // unless a whole warp is rejected.
        const bool valid = i < count;
        bool continue_path = false, cast_shadow = false;
        const unsigned int pixel = queue[valid ? i : count - 1u];
        float4 shadow_o = make_float4(0, 0, 0, 0), shadow_d = make_float4(0, 0, 0, 0), shadow_rad = make_float4(0, 0, 0, 0);

        const float4 ro = w.ray_o[pixel], rd = w.ray_d[pixel];
        float4 thr4 = w.thr[pixel], rad4 = w.rad[pixel];
        const float4 hit4 = w.hit[pixel];
        const float3 ray_origin = f3(ro), ray_direction = f3(rd);
        Pdf bsdf_pdf(rd.w);
        float3 throughput = f3(thr4), radiance = f3(rad4);
        unsigned int bounces = __float_as_uint(thr4.w);
        int previous_primitive = __float_as_int(rad4.w);
        const float t_hit = hit4.x;
        const int primitive = __float_as_int(hit4.y);
        const unsigned int pixel_hash = pcg2d(pixel % (unsigned int)f.width, pixel / (unsigned int)f.width).x;

        float3 next_origin = ray_origin, next_direction = ray_direction;
        float next_tmin = ro.w;

        if constexpr (!SURFACE) {
            if (primitive < 0) {
                // miss, SimpleRGPs.cu:349-362
                float3 environment_radiance = s.env.tint;
                if (s.env.texels != nullptr) {
                    environment_radiance = environment_light::evaluate(s.env, ray_direction);
                    if (bsdf_pdf.use_for_MIS())
                        environment_radiance *= mis_weight(bsdf_pdf, environment_light::pdf(s.env, ray_direction));
                }
                radiance += throughput * environment_radiance;
                throughput = f3(0.0f);
            } else {
                // light_closest_hit, MonteCarlo.cu:291-302; evaluate_intersection, LightImpl.h:86-108
                Light light = s.lights[primitive & ~LIGHT_HIT_FLAG];
                float3 light_radiance = light_evaluate(light, s.env, ray_origin, ray_direction);
                if (bsdf_pdf.use_for_MIS())
                    light_radiance *= mis_weight(bsdf_pdf, light_pdf(light, s.env, ray_origin, ray_direction));
                throughput = min3(throughput, f3(4.0f));
                radiance += throughput * light_radiance;
                throughput = f3(0.0f);
            }
        } else {
            // ---- phase 0: hit attributes, reject rules, frame and material set-up ----
            // interpolate_attributes, TriangleAttributes.cu:35-84 (geometry is pre-transformed to world space)
            const float3 p0 = f3(__ldg(s.world_vertices + 3ll * primitive)), p1 = f3(__ldg(s.world_vertices + 3ll * primitive + 1)),
                         p2 = f3(__ldg(s.world_vertices + 3ll * primitive + 2));
            const int4* shade_raw = reinterpret_cast<const int4*>(s.shade + primitive);
            int4 sr0 = __ldg(shade_raw), sr1 = __ldg(shade_raw + 1);
            ShadeTriangle st;
            memcpy(&st, &sr0, 16); memcpy(reinterpret_cast<char*>(&st) + 16, &sr1, 16);

            float3 geometric_normal = normalize(cross(p1 - p0, p2 - p0));
            const float bx = hit4.z, by = hit4.w;
            const float bz = 1.0f - bx - by;
            const float3 intersection_point = p1 * bx + p2 * by + p0 * bz;
            const bool has_normals = st.flags & 1u, has_tints = st.flags & 2u;
            float3 shading_normal;
            if (has_normals) {
                shading_normal = oct_decode(st.n1) * bx + oct_decode(st.n2) * by + oct_decode(st.n0) * bz;
                shading_normal = normalize(shading_normal);
            } else
                shading_normal = geometric_normal;
            float4 tint_and_roughness_scale = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
            if (has_tints) {
                const float n255 = 1.0f / 255.0f;
                tint_and_roughness_scale.x = (st.t1[0] * bx + st.t2[0] * by + st.t0[0] * bz) * n255;
                tint_and_roughness_scale.y = (st.t1[1] * bx + st.t2[1] * by + st.t0[1] * bz) * n255;
                tint_and_roughness_scale.z = (st.t1[2] * bx + st.t2[2] * by + st.t0[2] * bz) * n255;
                tint_and_roughness_scale.w = (st.t1[3] * bx + st.t2[3] * by + st.t0[3] * bz) * n255;
            }

            // path_tracing_closest_hit, MonteCarlo.cu:129-233. The same-primitive test (:137-142) already
            // happened inside the traversal, which skips `previous_primitive`.
            const float2 texcoord = interpolate_texcoord(s.accel.textures, primitive, bx, by);
            const Material material_parameter = material_at(s.materials[st.material_index], s.accel.textures, texcoord);
            float3 world_geometric_normal = geometric_normal;
            bool hit_from_front = dot(world_geometric_normal, ray_direction) < 0.0f;
            bool backside_cull = !hit_from_front && !material_is_thin_walled(material_parameter);
            backside_cull &= !material_is_transmissive(material_parameter);

            float4 bsdf_coverage_random = path_rng_sample4f(accumulation_count, pixel_hash, bounces, DIM_BSDF);
            float coverage_cutoff = bsdf_coverage_random.w;
            float3 bsdf_random_uvs = f3(bsdf_coverage_random);
            float coverage = material_coverage(material_parameter, s.accel.textures, texcoord);
            bool discard_from_coverage = coverage < coverage_cutoff;
            const bool rejected = backside_cull || discard_from_coverage; // the ray goes on past this surface (:146-164)

            world_geometric_normal = hit_from_front ? world_geometric_normal : -world_geometric_normal;
            float3 world_shading_normal = shading_normal;
            if (has_normals) {
                const float* nm = s.normal_matrices + 9 * (st.flags >> 2);
                world_shading_normal = normalize(f3(nm[0] * shading_normal.x + nm[1] * shading_normal.y + nm[2] * shading_normal.z,
                                                    nm[3] * shading_normal.x + nm[4] * shading_normal.y + nm[5] * shading_normal.z,
                                                    nm[6] * shading_normal.x + nm[7] * shading_normal.y + nm[8] * shading_normal.z));
            }
            world_shading_normal = hit_from_front ? world_shading_normal : -world_shading_normal;
            world_shading_normal = fix_backfacing_shading_normal(-ray_direction, world_shading_normal, 0.002f);
            const Tbn tbn(world_shading_normal);

            const float3 world_intersection_point = intersection_point;
            const float3 wo = tbn.to_local(-ray_direction);
            float cos_theta = hit_from_front || material_is_thin_walled(material_parameter) ? wo.z : -wo.z;

            // DefaultMaterialCreator::create, MonteCarlo.cu:239-244. The per-vertex scale goes through the
            // payload as unorm8 only for the AOV backends; shading uses the interpolated floats.
            Pdf max_pdf_hint(bsdf_pdf.v * path_regularization_pdf_scale);
            const auto material = [&]() {
                if constexpr (TRANSMISSIVE)
                    return TransmissiveShading::create_regularized(tables, s.dielectric_tables, material_parameter, tint_and_roughness_scale,
                                                                   cos_theta, max_pdf_hint);
                else
                    return material_parameter.shading_model == SHADING_DIFFUSE
                        ? DefaultShading::create_diffuse(material_parameter, tint_and_roughness_scale)
                        : DefaultShading::create_regularized(tables, material_parameter, tint_and_roughness_scale, cos_theta, max_pdf_hint);
            }();

            float3 emission = f3(1.0f); // multiplicative identity, TriangleAttributes.cu:83
            if (s.shade_emission != nullptr) { // per-vertex emission scale, TriangleAttributes.cu:78-81
                const float* e = s.shade_emission + 9ll * primitive;
                emission = f3(e[3], e[4], e[5]) * bx + f3(e[6], e[7], e[8]) * by + f3(e[0], e[1], e[2]) * bz;
            }
            if (!rejected) radiance += throughput * emission * f3(material_parameter.emission[0], material_parameter.emission[1], material_parameter.emission[2]);

            // ---- phases 1 .. N: reestimated_light_samples, MonteCarlo.cu:91-123, one candidate per phase ----
            LightSample light_sample = light_sample_none();
            if (s.light_count != 0) {
                float4 light_random_base = path_rng_sample4f(accumulation_count, pixel_hash, bounces, DIM_NEE);
                for (int k = 0; k < f.next_event_sample_count; ++k) {
                    SHADE_PHASE_BARRIER();
                    float4 shift = __ldg(s.nee_offsets + k);
                    float4 r = light_random_base + shift; // toroidal_shift, Utils.h:46-49
                    r = make_float4(r.x - floorf(r.x), r.y - floorf(r.y), r.z - floorf(r.z), r.w - floorf(r.w));
                    LightSample candidate = sample_single_light(s, material, world_intersection_point, wo, tbn, f3(r));
                    float light_weight = sum(light_sample.radiance);
                    float new_light_weight = sum(candidate.radiance);
                    float new_light_probability = fdiv(new_light_weight, light_weight + new_light_weight);
                    if (r.w < new_light_probability) {
                        light_sample = candidate;
                        light_sample.radiance /= new_light_probability;
                    } else
                        light_sample.radiance /= 1.0f - new_light_probability;
                }
                light_sample.radiance /= float(f.next_event_sample_count);
            }
            float3 light_sample_origin = offset_ray_origin(world_intersection_point, light_sample.direction_to_light, world_geometric_normal);
            light_sample.radiance *= throughput;

            // ---- last phase: BSDF sampling, throughput, the next ray ----
            SHADE_PHASE_BARRIER();
            BsdfSample bsdf_sample = material.sample(wo, bsdf_random_uvs);
            if (rejected) {
                next_tmin = nextafterf(t_hit, INFINITY); // same ray, advanced past this surface
            } else {
                previous_primitive = primitive;
                bool is_reflection = bsdf_sample.direction.z >= 0;
                next_direction = tbn.to_world(bsdf_sample.direction);
                bsdf_pdf = bsdf_sample.pdf;
                if (bsdf_sample.pdf.is_valid())
                    throughput *= bsdf_sample.reflectance * fabsf(bsdf_sample.direction.z) / bsdf_sample.pdf.value();
                else
                    throughput = f3(0.0f);

                float cos_geometric_theta_i = dot(next_direction, world_geometric_normal);
                if (is_reflection ? cos_geometric_theta_i < 0.0f : cos_geometric_theta_i >= 0.0f)
                    next_direction = reflect(next_direction, world_geometric_normal);

                next_origin = offset_ray_origin(world_intersection_point, next_direction, world_geometric_normal);
                next_tmin = 0.0f;
                // Russian roulette, opt-in (bpt_settings.russian_roulette_start_bounce): survival probability =
                // the largest throughput component, decided by the otherwise unused RNG dimension 3 of this bounce.
                if (f.russian_roulette_start_bounce != 0u && bounces + 1u >= f.russian_roulette_start_bounce && !is_black(throughput)) {
                    float survival = clampf(fmaxf(fmaxf(throughput.x, throughput.y), throughput.z), 0.05f, 1.0f);
                    float u = path_rng_sample4f(accumulation_count, pixel_hash, bounces, DIM_ROULETTE).x;
                    if (u < survival) throughput = throughput / survival;
                    else throughput = f3(0.0f);
                }
                bounces += 1u;
                if (!light_sample.pdf.is_valid())
                    bsdf_pdf.disable_MIS();

                // path_trace_single_bounce, SimpleRGPs.cu:117-125
                if (valid && (light_sample.radiance.x > 0 || light_sample.radiance.y > 0 || light_sample.radiance.z > 0)) {
                    cast_shadow = true;
                    shadow_o = f4(light_sample_origin, light_sample.distance);
                    shadow_d = f4(light_sample.direction_to_light, __uint_as_float(pixel));
                    shadow_rad = f4(light_sample.radiance, fmaxf(fmaxf(light_sample.radiance.x, light_sample.radiance.y), light_sample.radiance.z));
                }
            }
        }

        if (valid) {
            continue_path = bounces <= f.max_bounce_count && !is_black(throughput); // SimpleRGPs.cu:136
            w.rad[pixel] = f4(radiance, __int_as_float(previous_primitive));
            if (continue_path) {
                w.ray_o[pixel] = f4(next_origin, next_tmin);
                w.ray_d[pixel] = f4(next_direction, bsdf_pdf.v);
                w.thr[pixel] = f4(throughput, __uint_as_float(bounces));
            }
        }

        // Queue compaction: ballot + popc inside the warp, one atomic per warp and queue.
        warp_append(continue_path, queue_out, &w.counters->next_active, pixel);
        {
            unsigned int mask = __ballot_sync(0xffffffffu, cast_shadow);
            if (mask) {
                int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
                unsigned int base = 0;
                if (lane == leader) base = atomicAdd(shadow_counter, __popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (cast_shadow) {
                    unsigned int slot = base + __popc(mask & ((1u << lane) - 1u));
                    w.sh_o[slot] = shadow_o; w.sh_d[slot] = shadow_d; w.sh_rad[slot] = shadow_rad;
                }
            }
        }
    }
}


__global__ void set_frame_state_kernel(FrameState* destination, FrameState value) { *destination = value; }
// Last node of a sample: the lane's next sample (no kernel of this sample reads the index any more).
__global__ void finish_sample_kernel(FrameState* frame) { frame->sample_index += frame->sample_stride; }

// ---- accumulate / resolve --------------------------------------------------------------------------------

// accumulate<>, SimpleRGPs.cu:74-107. The reference keeps a running mean in fp64; here the fp64 SUM and the
// sample count are kept (mean = sum / count on resolve), which lets sample ranges rendered on different GPUs be
// combined with one sum-reduce.
// A sample whose radiance is not finite is dropped (neither the sum nor the pixel's sample count change) and counted in
// bpt_counters.nonfinite_samples: a single NaN would otherwise poison the pixel's fp64 sum for the rest of the render.
__global__ void accumulate_kernel(const float4* __restrict__ rad, double* __restrict__ accum, int64_t pixel_count, unsigned long long* __restrict__ nonfinite) {
    unsigned int dropped = 0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pixel_count; p += (int64_t)gridDim.x * blockDim.x) {
        float4 r = rad[p];
        if (!(isfinite(r.x) && isfinite(r.y) && isfinite(r.z))) { ++dropped; continue; }
        double2* a = reinterpret_cast<double2*>(accum + 4 * p);
        double2 rg = a[0], bw = a[1];
        rg.x += (double)r.x; rg.y += (double)r.y; bw.x += (double)r.z; bw.y += 1.0;
        a[0] = rg; a[1] = bw;
    }
    if (dropped) atomicAdd(nonfinite, (unsigned long long)dropped);
}

__global__ void resolve_half4_kernel(const double* __restrict__ accum, ushort4* __restrict__ out, int64_t pixel_count, float scale) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pixel_count; p += (int64_t)gridDim.x * blockDim.x) {
        const double2* a = reinterpret_cast<const double2*>(accum + 4 * p);
        double2 rg = a[0], bw = a[1];
        double inv = bw.y > 0.0 ? 1.0 / bw.y : 0.0;
        float3 mean = f3(float(rg.x * inv), float(rg.y * inv), float(bw.x * inv));
        if (scale != 1.0f) mean = mean / scale; // depth backend: d = depth / max_depth
        // float_to_half, SimpleRGPs.cu:39-42
        out[p] = make_ushort4(__half_as_ushort(__float2half_rn(mean.x)), __half_as_ushort(__float2half_rn(mean.y)),
                              __half_as_ushort(__float2half_rn(mean.z)), __half_as_ushort(__float2half_rn(1.0f)));
    }
}

__global__ void resolve_float4_kernel(const double* __restrict__ accum, float4* __restrict__ out, int64_t pixel_count) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < pixel_count; p += (int64_t)gridDim.x * blockDim.x) {
        const double2* a = reinterpret_cast<const double2*>(accum + 4 * p);
        double2 rg = a[0], bw = a[1];
        double inv = bw.y > 0.0 ? 1.0 / bw.y : 0.0;
        out[p] = make_float4(float(rg.x * inv), float(rg.y * inv), float(bw.x * inv), 1.0f);
    }
}

// Several samples in flight. A sample is a chain of dependent kernels whose late iterations hold few paths, and every link of
// the chain ends in a tail where the SMs drain; other, independent chains fill those gaps. The sample with accumulation
// index i runs on lane i mod L (L = BPT_DEFAULT_LANES; measured 1 / 2 / 3 / 4 lanes: 572 / 657 / 669 / 672 Msamples/s on
// configs[2], 506 / 550 / 563 / 570 on configs[3]): own path state, own graph, own stream. The radiance of a finished
// sample is added to the accumulation target by a launch on the context's MAIN stream that waits for the lane, so
// the sums happen in index order whatever the lanes do (results are bit for bit those of a single lane), and everything else
// the caller enqueues on the main stream - resolves, uploads, the NCCL reduce - stays ordered behind the samples it follows.
// A lane waits for the main stream only after the scene changed (ctx->scene_epoch), and for the accumulation of its own
// previous sample (which reads the radiance buffer the next sample overwrites).
constexpr int MAX_LANES = 8;
struct Integrator {
    Wavefront lane[MAX_LANES];
    cudaStream_t stream[MAX_LANES] = {};
    cudaEvent_t sample_done[MAX_LANES] = {}, accumulated[MAX_LANES] = {}, scene_ready = nullptr;
    bool accumulated_recorded[MAX_LANES] = {};
    uint64_t scene_epoch_seen = ~0ull;
};

Integrator* integrator(Context* ctx) { return static_cast<Integrator*>(ctx->wavefront); }

// Everything the kernels of one sample are launched with. Its bytes are the signature the instantiated graph is keyed by.
struct SampleLaunch {
    WavefrontView w;
    SceneView s;
    FrameParams f;
    int64_t pixels;
    int trace_grid, shade_grid, stream_grid, escaped_grid;
    int transmissive; // the scene holds Transmissive materials: a third shade kernel per iteration
    int sort_hits;    // two more kernels per iteration: the counting sort of the surface queue
};

void destroy_graph(Wavefront* wf) {
    if (wf->graph_exec) cudaGraphExecDestroy(wf->graph_exec);
    if (wf->graph) cudaGraphDestroy(wf->graph);
    wf->graph_exec = nullptr; wf->graph = nullptr;
    wf->graph_signature.clear();
}

cudaError_t add_kernel(cudaGraphNode_t* node, cudaGraph_t graph, const cudaGraphNode_t* dependencies, size_t dependency_count, const void* function,
                       int grid, int block, void** arguments) {
    cudaKernelNodeParams p = {};
    p.func = const_cast<void*>(function); p.gridDim = dim3(grid); p.blockDim = dim3(block); p.sharedMemBytes = 0;
    p.kernelParams = arguments; p.extra = nullptr;
    return cudaGraphAddKernelNode(node, graph, dependencies, dependency_count, &p);
}

// One sample as a graph:
//   generate -> WHILE(paths alive) { extend || shadow -> shade_escaped || shade_surface [|| shade_transmissive] -> advance }
//            -> shadow (the rays of the last shade) -> finish (the lane's next sample index)
// The accumulation of the sample's radiance is a separate launch on the context's main stream (see render()).
// The WHILE handle starts every launch at 1 (cudaGraphCondAssignDefault); advance_kernel sets it from the queue length.
cudaError_t build_sample_graph(Wavefront* wf, SampleLaunch& L) {
    destroy_graph(wf);
#define GRAPH_CHECK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { destroy_graph(wf); return _e; } } while (0)
    GRAPH_CHECK(cudaGraphCreate(&wf->graph, 0));
    cudaGraph_t graph = wf->graph;

    cudaGraphNode_t generate, loop, final_shadow, finish;
    void* generate_args[] = { &L.w, &L.f };
    GRAPH_CHECK(add_kernel(&generate, graph, nullptr, 0, (const void*)generate_kernel, L.stream_grid, 256, generate_args));

    cudaGraphConditionalHandle handle;
    GRAPH_CHECK(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams loop_params = {};
    loop_params.type = cudaGraphNodeTypeConditional;
    loop_params.conditional.handle = handle;
    loop_params.conditional.type = cudaGraphCondTypeWhile;
    loop_params.conditional.size = 1;
    GRAPH_CHECK(cudaGraphAddNode(&loop, graph, &generate, 1, &loop_params));
    cudaGraph_t body = loop_params.conditional.phGraph_out[0];

    cudaGraphNode_t traced[2], shaded[3], advance;
    void* trace_args[] = { &L.w, &L.s };
    void* shade_args[] = { &L.w, &L.s, &L.f };
    const bool compressed = L.s.accel.cw != nullptr;
    const void* extend_function = L.sort_hits ? (compressed ? (const void*)extend_kernel<true, true> : (const void*)extend_kernel<true, false>)
                                              : (compressed ? (const void*)extend_kernel<false, true> : (const void*)extend_kernel<false, false>);
    const void* shadow_function = compressed ? (const void*)shadow_kernel<true> : (const void*)shadow_kernel<false>;
    GRAPH_CHECK(add_kernel(&traced[0], body, nullptr, 0, extend_function, L.trace_grid, TRACE_BLOCK, shade_args));
    GRAPH_CHECK(add_kernel(&traced[1], body, nullptr, 0, shadow_function, L.trace_grid, TRACE_BLOCK, trace_args));
    size_t shade_count = 0;
    GRAPH_CHECK(add_kernel(&shaded[shade_count++], body, traced, 2, (const void*)shade_kernel<false, false>, L.escaped_grid, SHADE_BLOCK, shade_args));
    if (L.sort_hits) { // the surface shading waits for the sorted queue; escaped paths are shaded beside the sort
        cudaGraphNode_t scan, scatter, surface_inputs[2];
        void* sort_args[] = { &L.w, &L.f };
        GRAPH_CHECK(add_kernel(&scan, body, &traced[0], 1, (const void*)sort_scan_kernel, 1, 1024, sort_args));
        GRAPH_CHECK(add_kernel(&scatter, body, &scan, 1, (const void*)sort_scatter_kernel, L.stream_grid, 256, sort_args));
        surface_inputs[0] = scatter; surface_inputs[1] = traced[1]; // the shadow kernel adds to the radiance shading reads
        GRAPH_CHECK(add_kernel(&shaded[shade_count++], body, surface_inputs, 2, (const void*)shade_kernel<true, false>, L.shade_grid, SHADE_BLOCK, shade_args));
    } else
        GRAPH_CHECK(add_kernel(&shaded[shade_count++], body, traced, 2, (const void*)shade_kernel<true, false>, L.shade_grid, SHADE_BLOCK, shade_args));
    if (L.transmissive)
        GRAPH_CHECK(add_kernel(&shaded[shade_count++], body, traced, 2, (const void*)shade_kernel<true, true>, L.shade_grid, SHADE_BLOCK, shade_args));
    QueueCounters* counters = L.w.counters;
    unsigned long long* ray_counters = L.w.ray_counters;
    int has_handle = 1;
    void* advance_args[] = { &counters, &ray_counters, &handle, &has_handle };
    GRAPH_CHECK(add_kernel(&advance, body, shaded, shade_count, (const void*)advance_kernel, 1, 1, advance_args));

    GRAPH_CHECK(add_kernel(&final_shadow, graph, &loop, 1, shadow_function, L.trace_grid, TRACE_BLOCK, trace_args));
    FrameState* frame = const_cast<FrameState*>(L.w.frame);
    void* finish_args[] = { &frame };
    GRAPH_CHECK(add_kernel(&finish, graph, &final_shadow, 1, (const void*)finish_sample_kernel, 1, 1, finish_args));

    GRAPH_CHECK(cudaGraphInstantiate(&wf->graph_exec, graph, 0));
#undef GRAPH_CHECK
    wf->graph_signature.assign(reinterpret_cast<const unsigned char*>(&L), reinterpret_cast<const unsigned char*>(&L) + sizeof(L));
    return cudaSuccess;
}

// The same sample with plain stream launches; the host reads the queue length back to decide when the loop ends. Used when
// per-stage timing is on (bpt_set_profiling: CUDA events between the stages), with BPT_GRAPH=0, or when the driver refuses
// conditional graph nodes.
int launch_sample_serial(Context* ctx, SampleLaunch& L, cudaStream_t st) {
    const bool compressed = L.s.accel.cw != nullptr;
    generate_kernel<<<L.stream_grid, 256, 0, st>>>(L.w, L.f);
    // A path shades at most max_bounce_count + 1 surfaces; rejected hits (back faces, coverage) re-trace the same ray
    // without consuming a bounce, so a few extra iterations run before the queue length is checked on the host.
    uint32_t planned = L.f.max_bounce_count + 2, done = 0;
    while (true) {
        for (uint32_t it = 0; it < planned; ++it) {
            if (ctx->profiling) cudaEventRecord(ctx->stage_events[ctx->stage_event(it * 4 + 0)], st);
            if (compressed) {
                if (L.sort_hits) extend_kernel<true, true><<<L.trace_grid, TRACE_BLOCK, 0, st>>>(L.w, L.s, L.f);
                else extend_kernel<false, true><<<L.trace_grid, TRACE_BLOCK, 0, st>>>(L.w, L.s, L.f);
            } else {
                if (L.sort_hits) extend_kernel<true, false><<<L.trace_grid, TRACE_BLOCK, 0, st>>>(L.w, L.s, L.f);
                else extend_kernel<false, false><<<L.trace_grid, TRACE_BLOCK, 0, st>>>(L.w, L.s, L.f);
            }
            if (ctx->profiling) cudaEventRecord(ctx->stage_events[ctx->stage_event(it * 4 + 1)], st);
            if (compressed) shadow_kernel<true><<<L.trace_grid, TRACE_BLOCK, 0, st>>>(L.w, L.s);
            else shadow_kernel<false><<<L.trace_grid, TRACE_BLOCK, 0, st>>>(L.w, L.s);
            if (ctx->profiling) cudaEventRecord(ctx->stage_events[ctx->stage_event(it * 4 + 2)], st);
            if (L.sort_hits) { // timed with the shading it serves
                sort_scan_kernel<<<1, 1024, 0, st>>>(L.w, L.f);
                sort_scatter_kernel<<<L.stream_grid, 256, 0, st>>>(L.w, L.f);
            }
            shade_kernel<false, false><<<L.escaped_grid, SHADE_BLOCK, 0, st>>>(L.w, L.s, L.f);
            shade_kernel<true, false><<<L.shade_grid, SHADE_BLOCK, 0, st>>>(L.w, L.s, L.f);
            if (L.transmissive) shade_kernel<true, true><<<L.shade_grid, SHADE_BLOCK, 0, st>>>(L.w, L.s, L.f);
            if (ctx->profiling) cudaEventRecord(ctx->stage_events[ctx->stage_event(it * 4 + 3)], st);
            advance_kernel<<<1, 1, 0, st>>>(L.w.counters, L.w.ray_counters, 0ull, 0);
        }
        done += planned;
        QueueCounters h;
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(&h, L.w.counters, sizeof(h), cudaMemcpyDeviceToHost, st));
        BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        if (ctx->profiling)
            for (uint32_t it = 0; it < planned; ++it) {
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, ctx->stage_events[it * 4 + 0], ctx->stage_events[it * 4 + 1]); ctx->counters.extend_ms += ms;
                cudaEventElapsedTime(&ms, ctx->stage_events[it * 4 + 1], ctx->stage_events[it * 4 + 2]); ctx->counters.shadow_ms += ms;
                cudaEventElapsedTime(&ms, ctx->stage_events[it * 4 + 2], ctx->stage_events[it * 4 + 3]); ctx->counters.shade_ms += ms;
            }
        if (h.active == 0) break;
        if (done > MAX_ITERATIONS_PER_SAMPLE) return ctx->fail(BPT_ERROR_CUDA, "bpt_render: path queue did not drain");
        planned = 2;
    }
    if (ctx->profiling) cudaEventRecord(ctx->stage_events[ctx->stage_event(0)], st);
    if (compressed) shadow_kernel<true><<<L.trace_grid, TRACE_BLOCK, 0, st>>>(L.w, L.s); // the shadow rays of the last shade
    else shadow_kernel<false><<<L.trace_grid, TRACE_BLOCK, 0, st>>>(L.w, L.s);
    if (ctx->profiling) {
        cudaEventRecord(ctx->stage_events[ctx->stage_event(1)], st);
        cudaEventSynchronize(ctx->stage_events[1]);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ctx->stage_events[0], ctx->stage_events[1]); ctx->counters.shadow_ms += ms;
    }
    finish_sample_kernel<<<1, 1, 0, st>>>(const_cast<FrameState*>(L.w.frame));
    return BPT_OK;
}

} // namespace

void release_wavefront(Context* ctx) {
    Integrator* in = integrator(ctx);
    if (!in) return;
    for (int l = 0; l < MAX_LANES; ++l) {
        if (in->stream[l]) cudaStreamSynchronize(in->stream[l]);
        Wavefront* wf = &in->lane[l];
        destroy_graph(wf);
        wf->ray_o.release(); wf->ray_d.release(); wf->thr.release(); wf->rad.release(); wf->hit.release();
        wf->sh_o.release(); wf->sh_d.release(); wf->sh_rad.release(); wf->queue_a.release(); wf->queue_b.release();
        wf->queue_surface.release(); wf->queue_escaped.release();
        wf->counters.release(); wf->frame_state.release(); wf->coverage.release();
        wf->surface_key.release(); wf->queue_surface_sorted.release(); wf->sort_bins.release(); wf->material_class.release();
        if (in->stream[l]) cudaStreamDestroy(in->stream[l]);
        if (in->sample_done[l]) cudaEventDestroy(in->sample_done[l]);
        if (in->accumulated[l]) cudaEventDestroy(in->accumulated[l]);
    }
    if (in->scene_ready) cudaEventDestroy(in->scene_ready);
    delete in;
    ctx->wavefront = nullptr;
}

namespace {

// Buffers, per-material tables and the launch description of one lane for this call.
int prepare_lane(Context* ctx, Wavefront* wf, const bpt_settings* settings, int width, int height, int lanes, SampleLaunch& L) {
    cudaStream_t st = ctx->stream;
    const int64_t pixels = (int64_t)width * height;
    if (wf->pixel_capacity < pixels) {
        BPT_CUDA_CHECK(ctx, cudaDeviceSynchronize()); // launches in flight (on any lane) still use the old buffers
        BPT_CUDA_CHECK(ctx, wf->ray_o.resize(pixels)); BPT_CUDA_CHECK(ctx, wf->ray_d.resize(pixels)); BPT_CUDA_CHECK(ctx, wf->thr.resize(pixels));
        BPT_CUDA_CHECK(ctx, wf->rad.resize(pixels)); BPT_CUDA_CHECK(ctx, wf->hit.resize(pixels));
        BPT_CUDA_CHECK(ctx, wf->sh_o.resize(pixels)); BPT_CUDA_CHECK(ctx, wf->sh_d.resize(pixels)); BPT_CUDA_CHECK(ctx, wf->sh_rad.resize(pixels));
        BPT_CUDA_CHECK(ctx, wf->queue_a.resize(pixels)); BPT_CUDA_CHECK(ctx, wf->queue_b.resize(pixels));
        BPT_CUDA_CHECK(ctx, wf->queue_surface.resize(pixels)); BPT_CUDA_CHECK(ctx, wf->queue_escaped.resize(pixels));
        BPT_CUDA_CHECK(ctx, wf->surface_key.resize(pixels)); BPT_CUDA_CHECK(ctx, wf->queue_surface_sorted.resize(pixels));
        BPT_CUDA_CHECK(ctx, wf->sort_bins.resize(2 * SORT_BINS));
        BPT_CUDA_CHECK(ctx, cudaMemsetAsync(wf->sort_bins.ptr, 0, 2 * SORT_BINS * sizeof(unsigned int), st));
        BPT_CUDA_CHECK(ctx, wf->counters.resize(1)); BPT_CUDA_CHECK(ctx, wf->frame_state.resize(1));
        BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        wf->pixel_capacity = pixels;
    }
    if (wf->coverage_version != ctx->material_version) {
        std::vector<float> h_cov(ctx->host_materials.size());
        std::vector<unsigned char> h_class(std::max<size_t>(ctx->host_materials.size(), 1), 0);
        for (size_t i = 0; i < h_cov.size(); ++i) {
            const Material& m = ctx->host_materials[i];
            h_cov[i] = material_coverage_table_entry(m);
            // what the surface shading branches on: the shading model (DefaultShading.h vs DiffuseShading.h) and the coat lobe
            h_class[i] = (unsigned char)((m.shading_model == SHADING_DIFFUSE ? 1u : 0u) | (m.coat != 0 ? 2u : 0u));
        }
        BPT_CUDA_CHECK(ctx, cudaDeviceSynchronize()); // as above when the tables have to grow
        BPT_CUDA_CHECK(ctx, wf->coverage.resize(h_cov.size()));
        BPT_CUDA_CHECK(ctx, wf->material_class.resize(h_class.size()));
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(wf->material_class.ptr, h_class.data(), h_class.size(), cudaMemcpyHostToDevice, st));
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(wf->coverage.ptr, h_cov.data(), h_cov.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st)); // h_cov goes out of scope
        wf->coverage_version = ctx->material_version;
    }

    memset(&L, 0, sizeof(L)); // padding included: the bytes are compared
    SceneView& s = L.s;
    s.accel = accel_view(ctx);
    s.world_vertices = ctx->accel.world_vertices.ptr;
    s.shade = ctx->accel.shade.ptr;
    s.normal_matrices = ctx->accel.normal_matrices.ptr;
    s.shade_emission = ctx->accel.has_emission ? ctx->accel.shade_emission.ptr : nullptr;
    s.materials = ctx->materials.ptr;
    s.coverage = wf->coverage.ptr;
    s.lights = ctx->lights.ptr;
    s.analytic_light_count = ctx->light_count;
    s.light_count = ctx->light_count;
    const bool env_by_cdf = ctx->env_nee_mode == BPT_ENVIRONMENT_NEE_CDF;
    s.env = environment_view(ctx, env_by_cdf);
    if (ctx->env_width > 0 && ctx->env_sample_count > 1) {
        // next_event_estimation_possible (PresampledEnvironmentMap.h:64): the environment is appended to the light list
        // (Renderer.cpp:1180-1195). bpt_set_lights reserved the slot.
        if (!ctx->lights.ptr) BPT_CUDA_CHECK(ctx, ctx->lights.resize(1));
        if (!ctx->env_light_uploaded) {
            Light env_light = {};
            env_light.flags = env_by_cdf ? BPT_LIGHT_ENVIRONMENT : BPT_LIGHT_PRESAMPLED_ENVIRONMENT;
            BPT_CUDA_CHECK(ctx, cudaDeviceSynchronize());
            BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->lights.ptr + ctx->light_count, &env_light, sizeof(Light), cudaMemcpyHostToDevice, st));
            BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
            ctx->env_light_uploaded = true;
        }
        s.lights = ctx->lights.ptr;
        s.light_count = ctx->light_count + 1;
    }
    s.tables = ctx->tables.ptr;
    s.dielectric_tables = ctx->dielectric_tables.ptr;
    s.split_by_shading_model = ctx->has_transmissive_materials;
    s.nee_offsets = ctx->nee_offsets.ptr;
    s.slot_of_primitive = ctx->accel.slot_of_primitive.ptr;
    s.material_class = wf->material_class.ptr;
    s.sort_cell_shift = 0;
    while ((ctx->accel.triangle_count >> s.sort_cell_shift) > (int64_t)SORT_CELLS) ++s.sort_cell_shift;

    WavefrontView& w = L.w;
    w.ray_o = wf->ray_o.ptr; w.ray_d = wf->ray_d.ptr; w.thr = wf->thr.ptr; w.rad = wf->rad.ptr; w.hit = wf->hit.ptr;
    w.sh_o = wf->sh_o.ptr; w.sh_d = wf->sh_d.ptr; w.sh_rad = wf->sh_rad.ptr;
    w.queue_a = wf->queue_a.ptr; w.queue_b = wf->queue_b.ptr;
    w.queue_surface = wf->queue_surface.ptr; w.queue_escaped = wf->queue_escaped.ptr;
    w.queue_capacity = (unsigned int)pixels;
    w.surface_key = wf->surface_key.ptr; w.queue_surface_sorted = wf->queue_surface_sorted.ptr; w.sort_bins = wf->sort_bins.ptr;
    w.counters = wf->counters.ptr;
    w.frame = wf->frame_state.ptr;
    w.ray_counters = reinterpret_cast<unsigned long long*>(ctx->device_counters);

    FrameParams& f = L.f;
    f.width = width; f.height = height;
    f.max_bounce_count = settings->max_bounce_count;
    f.next_event_sample_count = settings->next_event_sample_count;
    f.russian_roulette_start_bounce = settings->russian_roulette_start_bounce;
    f.sort_hits_from_iteration = ctx->sort_hits_from_iteration < 0 ? 0xffffffffu : (unsigned int)ctx->sort_hits_from_iteration;

    // Persistent grids: a whole number of CTAs per SM.
    L.pixels = pixels;
    // With several samples in flight a traversal kernel does not need the whole GPU to itself: 4 CTAs per SM instead of the 8
    // that fit let kernels of different lanes share an SM instead of queueing behind each other's CTAs (measured with 4 lanes,
    // 8 / 6 / 5 / 4 / 2 CTAs per SM: 671 / 679 / 678 / 685 / 693 Msamples/s on configs[2], 499 / 508 / 511 / 513 / 520 on
    // configs[1], 569 / 558 / - / 567 / - on configs[3]). A single chain (profiling, BPT_LANES=1) keeps the full grid.
    // BPT_TRACE_CTAS / BPT_SHADE_CTAS in the environment override both for tuning runs.
    static const int trace_ctas_env = [] { const char* e = getenv("BPT_TRACE_CTAS"); return e ? atoi(e) : 0; }();
    static const int shade_ctas_env = [] { const char* e = getenv("BPT_SHADE_CTAS"); return e ? atoi(e) : 0; }();
    const int trace_ctas = trace_ctas_env > 0 ? trace_ctas_env : (lanes > 1 ? BPT_TRACE_GRID_CTAS_SHARED : BPT_TRACE_MIN_BLOCKS);
    L.trace_grid = ctx->sm_count * trace_ctas;
    L.shade_grid = ctx->sm_count * (shade_ctas_env > 0 ? shade_ctas_env : BPT_SHADE_MIN_BLOCKS);
    L.stream_grid = ctx->sm_count * 8;
    L.escaped_grid = ctx->sm_count * 4;
    L.transmissive = ctx->has_transmissive_materials ? 1 : 0;
    L.sort_hits = ctx->sort_hits_from_iteration >= 0 ? 1 : 0;
    ctx->launches_per_iteration = 5 + L.transmissive + 2 * L.sort_hits;
    return BPT_OK;
}

} // namespace

int render(Context* ctx, const bpt_camera* camera, const bpt_settings* settings, int width, int height,
           uint32_t first_sample, uint32_t sample_count, int reset_accumulation) {
    if (!camera || !settings || width <= 0 || height <= 0) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_render: bad arguments");
    if (!ctx->accel.valid) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_render: call bpt_build_accel first");
    if (!ctx->has_tables) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_render: call bpt_set_tables first");
    if (ctx->has_transmissive_materials && !ctx->has_dielectric_tables)
        return ctx->fail(BPT_ERROR_NOT_READY, "bpt_render: transmissive materials need bpt_set_dielectric_tables");
    if (int status = sync_texture_table(ctx)) return status;
    if (settings->next_event_sample_count < 0 || settings->next_event_sample_count > 256)
        return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_render: next_event_sample_count must be in [0, 256]");
    cudaStream_t st = ctx->stream;
    const int64_t pixels = (int64_t)width * height;
    if (pixels > 0x7fffffffll) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_render: frame too large");

    ctx->half4_scale = 1.0f;
    if (!ctx->wavefront) ctx->wavefront = new Integrator();
    Integrator* in = integrator(ctx);

    static const bool graphs_disabled = [] { const char* e = getenv("BPT_GRAPH"); return e && e[0] == '0'; }();
    static const int lane_setting = [] { const char* e = getenv("BPT_LANES"); int n = e ? atoi(e) : BPT_DEFAULT_LANES; return n < 1 ? 1 : (n > MAX_LANES ? MAX_LANES : n); }();
    bool use_graph = !ctx->profiling && !graphs_disabled && !in->lane[0].graph_unavailable;
    int lanes = use_graph ? lane_setting : 1;

    SampleLaunch L[MAX_LANES];
    for (int l = 0; l < lanes; ++l)
        if (int status = prepare_lane(ctx, &in->lane[l], settings, width, height, lanes, L[l])) return status;

    bool size_changed = ctx->width != width || ctx->height != height || ctx->accumulation.size != (size_t)(4 * pixels);
    if (size_changed) {
        if (ctx->accumulation.capacity < (size_t)(4 * pixels)) BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
        BPT_CUDA_CHECK(ctx, ctx->accumulation.resize(4 * pixels));
        ctx->width = width; ctx->height = height;
    }
    if (size_changed || reset_accumulation)
        BPT_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->accumulation.ptr, 0, sizeof(double) * 4 * pixels, st));

    if (use_graph)
        for (int l = 0; l < lanes; ++l) {
            Wavefront* wf = &in->lane[l];
            const bool current = wf->graph_exec && wf->graph_signature.size() == sizeof(SampleLaunch) && memcmp(wf->graph_signature.data(), &L[l], sizeof(SampleLaunch)) == 0;
            if (current) continue;
            cudaError_t e = build_sample_graph(wf, L[l]);
            if (e != cudaSuccess) {
                // e.g. a driver without conditional nodes: remember it, say so once, and take the stream path
                in->lane[0].graph_unavailable = true;
                ctx->last_error = std::string("bpt_render: sample graph unavailable (") + cudaGetErrorString(e) + "), using stream launches";
                fprintf(stderr, "%s\n", ctx->last_error.c_str());
                cudaGetLastError();
                use_graph = false; lanes = 1;
                break;
            }
        }

    // What changes per call travels through device memory, as the by-value argument of a one-thread kernel: fully
    // asynchronous, no staging buffer whose lifetime the host would have to track.
    FrameState frame_state;
    memset(&frame_state, 0, sizeof(frame_state));
    frame_state.camera = *camera;
    frame_state.path_regularization_pdf_scale = settings->path_regularization_pdf_scale;
    frame_state.sample_stride = (unsigned int)lanes;
    unsigned long long* nonfinite = reinterpret_cast<unsigned long long*>(ctx->device_counters) + 6;

    if (lanes > 1) {
        for (int l = 0; l < lanes; ++l)
            if (!in->stream[l]) {
                BPT_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&in->stream[l], cudaStreamNonBlocking));
                BPT_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&in->sample_done[l], cudaEventDisableTiming));
                BPT_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&in->accumulated[l], cudaEventDisableTiming));
            }
        if (!in->scene_ready) BPT_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&in->scene_ready, cudaEventDisableTiming));
        if (in->scene_epoch_seen != ctx->scene_epoch) { // uploads / builds since the last render: the lanes start behind them
            BPT_CUDA_CHECK(ctx, cudaEventRecord(in->scene_ready, st));
            for (int l = 0; l < lanes; ++l) BPT_CUDA_CHECK(ctx, cudaStreamWaitEvent(in->stream[l], in->scene_ready, 0));
            in->scene_epoch_seen = ctx->scene_epoch;
        }
        bool lane_started[MAX_LANES] = {};
        for (uint32_t k = 0; k < sample_count; ++k) {
            const int l = int((first_sample + k) % (uint32_t)lanes);
            cudaStream_t lane_stream = in->stream[l];
            if (in->accumulated_recorded[l]) BPT_CUDA_CHECK(ctx, cudaStreamWaitEvent(lane_stream, in->accumulated[l], 0)); // its radiance buffer is free again
            if (!lane_started[l]) {
                frame_state.sample_index = first_sample + k;
                set_frame_state_kernel<<<1, 1, 0, lane_stream>>>(in->lane[l].frame_state.ptr, frame_state);
                lane_started[l] = true;
            }
            BPT_CUDA_CHECK(ctx, cudaGraphLaunch(in->lane[l].graph_exec, lane_stream));
            BPT_CUDA_CHECK(ctx, cudaEventRecord(in->sample_done[l], lane_stream));
            BPT_CUDA_CHECK(ctx, cudaStreamWaitEvent(st, in->sample_done[l], 0));
            accumulate_kernel<<<L[l].stream_grid, 256, 0, st>>>(L[l].w.rad, ctx->accumulation.ptr, pixels, nonfinite);
            BPT_CUDA_CHECK(ctx, cudaEventRecord(in->accumulated[l], st));
            in->accumulated_recorded[l] = true;
        }
        for (int l = 0; l < lanes; ++l) ctx->counters.kernel_launches += lane_started[l] ? 1ull : 0ull;
    } else {
        // One lane on the main stream. The other lane may hold a sample of an earlier two-lane call: the main stream already
        // waits for it through that sample's accumulation.
        frame_state.sample_index = first_sample;
        set_frame_state_kernel<<<1, 1, 0, st>>>(in->lane[0].frame_state.ptr, frame_state);
        ctx->counters.kernel_launches += 1;
        for (uint32_t k = 0; k < sample_count; ++k) {
            if (use_graph) BPT_CUDA_CHECK(ctx, cudaGraphLaunch(in->lane[0].graph_exec, st));
            else if (int status = launch_sample_serial(ctx, L[0], st)) return status;
            accumulate_kernel<<<L[0].stream_grid, 256, 0, st>>>(L[0].w.rad, ctx->accumulation.ptr, pixels, nonfinite);
        }
        in->scene_epoch_seen = ~0ull; // the lanes have not seen what the main stream did meanwhile
    }
    ctx->counters.kernel_launches += 4ull * sample_count; // per sample: generate, the last shadow, finish, accumulate; the iterations are counted on the device
    ctx->counters.samples += (uint64_t)pixels * sample_count;
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

int resolve_half4(Context* ctx, uint16_t* out, int on_device) {
    if (!out || !ctx->accumulation.ptr) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_resolve_half4: nothing rendered");
    int64_t pixels = (int64_t)ctx->width * ctx->height;
    cudaStream_t st = ctx->stream;
    ushort4* d = reinterpret_cast<ushort4*>(out);
    if (!on_device) {
        // context-owned staging frame: no allocation on the per-frame path
        BPT_CUDA_CHECK(ctx, ctx->output_half4.resize(4 * pixels));
        d = reinterpret_cast<ushort4*>(ctx->output_half4.ptr);
    }
    resolve_half4_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->accumulation.ptr, d, pixels, ctx->half4_scale);
    ctx->counters.kernel_launches++;
    if (!on_device) {
        BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d, pixels * sizeof(ushort4), cudaMemcpyDeviceToHost, st));
        BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    }
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

// Enqueues the half4 resolve of the current accumulation on the render stream and its device -> host copy on the copy
// stream, and returns without waiting: the next bpt_render overlaps the copy. `out_host` should be pinned memory.
int resolve_half4_async(Context* ctx, uint16_t* out_host, int slot) {
    if (!out_host || !ctx->accumulation.ptr) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_resolve_half4_async: nothing rendered");
    if (slot < 0 || slot >= BPT_FRAME_SLOTS) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_resolve_half4_async: slot must be in [0, BPT_FRAME_SLOTS)");
    const int64_t pixels = (int64_t)ctx->width * ctx->height;
    cudaStream_t st = ctx->stream;
    if (!ctx->copy_stream) {
        BPT_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < BPT_FRAME_SLOTS; ++i) {
            BPT_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->frame_resolved[i], cudaEventDisableTiming));
            BPT_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->frame_copied[i], cudaEventDisableTiming));
        }
    }
    if (ctx->frame_staging[slot].size != (size_t)(4 * pixels)) {
        if (ctx->frame_in_flight[slot]) BPT_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->frame_copied[slot]));
        BPT_CUDA_CHECK(ctx, ctx->frame_staging[slot].resize(4 * pixels));
    }
    // the staging frame of this slot may still be read by its previous copy
    if (ctx->frame_in_flight[slot]) BPT_CUDA_CHECK(ctx, cudaStreamWaitEvent(st, ctx->frame_copied[slot], 0));
    resolve_half4_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->accumulation.ptr, reinterpret_cast<ushort4*>(ctx->frame_staging[slot].ptr), pixels,
                                                             ctx->half4_scale);
    ctx->counters.kernel_launches++;
    BPT_CUDA_CHECK(ctx, cudaEventRecord(ctx->frame_resolved[slot], st));
    BPT_CUDA_CHECK(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->frame_resolved[slot], 0));
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(out_host, ctx->frame_staging[slot].ptr, pixels * sizeof(ushort4), cudaMemcpyDeviceToHost, ctx->copy_stream));
    BPT_CUDA_CHECK(ctx, cudaEventRecord(ctx->frame_copied[slot], ctx->copy_stream));
    ctx->frame_in_flight[slot] = true;
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

int wait_frame(Context* ctx, int slot) {
    if (slot < 0 || slot >= BPT_FRAME_SLOTS) return ctx->fail(BPT_ERROR_INVALID_ARGUMENT, "bpt_wait_frame: slot must be in [0, BPT_FRAME_SLOTS)");
    if (ctx->frame_in_flight[slot]) {
        BPT_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->frame_copied[slot]));
        ctx->frame_in_flight[slot] = false;
    }
    return BPT_OK;
}

int resolve_float4(Context* ctx, float* out) {
    if (!out || !ctx->accumulation.ptr) return ctx->fail(BPT_ERROR_NOT_READY, "bpt_resolve_float4: nothing rendered");
    int64_t pixels = (int64_t)ctx->width * ctx->height;
    cudaStream_t st = ctx->stream;
    // context-owned staging frame: no allocation on the per-frame path
    if (ctx->output_float4.capacity < (size_t)(4 * pixels)) BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    BPT_CUDA_CHECK(ctx, ctx->output_float4.resize(4 * pixels));
    float4* d = reinterpret_cast<float4*>(ctx->output_float4.ptr);
    resolve_float4_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->accumulation.ptr, d, pixels);
    ctx->counters.kernel_launches++;
    BPT_CUDA_CHECK(ctx, cudaMemcpyAsync(out, d, pixels * sizeof(float4), cudaMemcpyDeviceToHost, st));
    BPT_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    BPT_CUDA_CHECK(ctx, cudaGetLastError());
    return BPT_OK;
}

} // namespace bpt
