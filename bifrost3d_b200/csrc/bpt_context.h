// Host-side context of the path tracer: owns every device allocation.
// Data layout in HBM is described in DESIGN.md ("Data layout").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>

#include "bpt_types.h"
#include "bpt_cw.cuh"

namespace bpt {

// Growable device buffer.
template <typename T>
struct DeviceBuffer {
    T* ptr = nullptr;
    size_t capacity = 0; // elements
    size_t size = 0;

    cudaError_t reserve(size_t n) {
        if (n <= capacity) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr; capacity = 0;
        cudaError_t e = cudaMalloc((void**)&ptr, n * sizeof(T));
        if (e == cudaSuccess) capacity = n;
        return e;
    }
    cudaError_t resize(size_t n) { cudaError_t e = reserve(n); if (e == cudaSuccess) size = n; return e; }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; capacity = size = 0; }
    size_t bytes() const { return size * sizeof(T); }
};

// A mesh lives on the device from bpt_upload_mesh on (load_mesh, Renderer.cpp:92-136): rebuilding the acceleration
// structure after a transform change re-flattens the resident meshes without touching the host again.
struct DeviceMesh {
    DeviceBuffer<uint32_t> indices;  // 3 per primitive
    DeviceBuffer<float> positions;   // 3 per vertex
    DeviceBuffer<int16_t> normals;   // 2 per vertex (octahedral); empty if the mesh has none
    DeviceBuffer<float> texcoords;   // 2 per vertex or empty
    DeviceBuffer<uint8_t> tints;     // 4 per vertex or empty
    DeviceBuffer<float> emission;    // 3 per vertex or empty
    int primitive_count = 0;
    int vertex_count = 0;
    void release() { indices.release(); positions.release(); normals.release(); texcoords.release(); tints.release(); emission.release(); primitive_count = vertex_count = 0; }
};

// BVH node, 64 bytes, two children per node: the child AABBs are stored in the parent so one 64 byte (4 x 128-bit)
// load decides both children. A link >= 0 is a child node index; a negative link is a leaf packed as
// ~(first_triangle | (count - 1) << 28) (bpt_trace.cuh: pack_leaf). An absent child is a far-away point box.
struct __align__(16) BvhNode {
    float4 lo_l_hi_l_x;   // left.lo.x, left.lo.y, left.lo.z, left.hi.x
    float4 hi_l_lo_r;     // left.hi.y, left.hi.z, right.lo.x, right.lo.y
    float4 lo_r_hi_r;     // right.lo.z, right.hi.x, right.hi.y, right.hi.z
    int left, right;
    int pad0, pad1;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

// Four-wide node, 128 bytes: the binary hierarchy with every other level removed (bpt_bvh.cu: collapse). Child boxes are
// stored per axis so that one 128-bit load brings the same plane of all four children; links as in BvhNode, an absent
// child is a far-away point box with link NODE_EMPTY.
struct __align__(16) WideNode {
    float4 lo_x, lo_y, lo_z, hi_x, hi_y, hi_z;
    int4 link;
    int4 pad;
};
static_assert(sizeof(WideNode) == 128, "WideNode must be 128 bytes");

// World-space triangle in traversal order: three float4 loads. The w lanes carry the global primitive
// index (bits of v0.w) and the material index (v1.w); v2.w is unused.
struct __align__(16) TraceTriangle {
    float4 v0, v1, v2;
};
static_assert(sizeof(TraceTriangle) == 48, "TraceTriangle must be 48 bytes");

// Per-primitive shading attributes, indexed by global primitive index (instance-major order).
struct __align__(16) ShadeTriangle {
    // Octahedral vertex normals (3 x short2), per-vertex tint/roughness (3 x uchar4), material index, flags.
    int16_t n0[2], n1[2], n2[2];
    uint8_t t0[4], t1[4], t2[4];
    int32_t material_index;
    uint32_t flags; // bit0 has normals, bit1 has tints
};
static_assert(sizeof(ShadeTriangle) == 32, "ShadeTriangle must be 32 bytes");

// One image + sampler pair (Renderer.cpp:650-751): a CUDA array and the texture object that reads it.
struct DeviceTexture {
    cudaArray_t array = nullptr;
    cudaTextureObject_t object = 0;
    int width = 0, height = 0, channels = 0;
};

struct Accel {
    DeviceBuffer<BvhNode> nodes;
    DeviceBuffer<WideNode> wide_nodes;           // four-wide collapse of `nodes`; used by traversal when wide_levels > 0
    int64_t wide_node_count = 0;
    int wide_levels = 0;                          // depth of the four-wide tree; 0 = traverse the binary nodes
    DeviceBuffer<CwNode> cw_nodes;               // compressed eight-wide collapse of `nodes` (bpt_cw.cuh); used when cw_levels > 0
    int64_t cw_node_count = 0;
    int cw_levels = 0;                            // depth of the eight-wide tree; 0 = not built / too deep for its stack
    DeviceBuffer<TraceTriangle> triangles;      // traversal order: Morton sorted, or grouped per eight-wide node when cw_levels > 0
    DeviceBuffer<float4> world_vertices;        // 3 per primitive, instance-major order: position.xyz, w unused
    DeviceBuffer<ShadeTriangle> shade;          // instance-major order
    DeviceBuffer<float2> shade_uv;              // 3 texcoords per primitive, instance-major order; only when a mesh has texcoords
    bool has_uv = false;
    bool built_for_textures = false;            // the last build saw textured materials (and kept whatever texcoords the meshes have)
    DeviceBuffer<float> shade_emission;         // 9 floats per primitive (per-vertex emission scale); only when a mesh has emission
    bool has_emission = false;
    DeviceBuffer<uint32_t> slot_of_primitive;   // global primitive index -> position in `triangles` (Morton order): the spatial
                                                // part of the key the surface hits are sorted by before shading
    DeviceBuffer<float> normal_matrices;        // 9 floats per instance record: inverse transpose of the upper 3x3
    int64_t triangle_count = 0;
    int64_t node_count = 0;
    float build_ms = 0.0f;
    int ploc_passes = 0, ploc_depth = 0;        // 0 when the plain LBVH hierarchy is in use
    bool ploc_on_device = false;                // the PLOC passes ran inside one cooperative launch
    float3 scene_lo, scene_hi;
    bool valid = false;
};

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int sm_count = 148;

    // Tables: [ggx_with_fresnel_rho | ggx_rho | estimate_alpha], 3 x 1024 floats.
    DeviceBuffer<float> tables;
    bool has_tables = false;
    // Dielectric GGX rho: [into_light_medium | into_dense_medium], 2 x 16^3 float2 (TransmissiveShading).
    DeviceBuffer<float2> dielectric_tables;
    bool has_dielectric_tables = false;
    bool has_transmissive_materials = false;
    bool use_wide = true; // BPT_WIDE=0 in the environment traverses the binary nodes (for A/B measurements)
    // Node format of the traversal: the compressed eight-wide nodes from this many triangles on, the uncompressed four-wide
    // ones below (measured on B200: 20 k triangles 504 vs 510 Msamples/s, 1 M 701 vs 681, 50 M 577 vs 560). BPT_CW=0 / 1 in the
    // environment sets it to never / always (A/B measurements, tests).
    int64_t cw_min_triangles = 131072;
    bool use_ploc = true; // BPT_BVH=lbvh in the environment selects the plain Morton hierarchy (for A/B measurements)
    // Surface hits are sorted by (shading class, hit cell) before shading from this wavefront iteration on; -1 = never.
    // bpt_set_hit_sorting; BPT_SORT_HITS in the environment sets the initial value.
    int sort_hits_from_iteration = -1;
    DeviceBuffer<float4> nee_offsets; // 256 ReverseHalton toroidal shifts (Renderer.cpp:323-336)

    std::map<int, DeviceMesh> meshes;
    std::vector<bpt_instance> instances;
    DeviceBuffer<Material> materials;
    std::map<int, DeviceTexture> textures;                  // by texture id (>= 1)
    DeviceBuffer<unsigned long long> texture_objects;       // cudaTextureObject_t by texture id, 0 where there is none
    bool texture_table_dirty = true;
    bool has_textured_materials = false;
    std::vector<Material> host_materials;
    DeviceBuffer<Light> lights;
    int light_count = 0;

    // Environment
    DeviceBuffer<float4> env_texels;
    DeviceBuffer<float> env_pdf;
    DeviceBuffer<bpt_light_sample> env_samples;
    int env_width = 0, env_height = 0, env_pdf_width = 0, env_pdf_height = 0, env_sample_count = 0;
    float env_tint[3] = {0.0f, 0.0f, 0.0f};
    DeviceBuffer<float> env_marginal_cdf, env_conditional_cdf; // bpt_set_environment_cdfs; sized for env_pdf_width x env_pdf_height
    bool env_has_cdfs = false;
    int env_nee_mode = 0; // BPT_ENVIRONMENT_NEE_PRESAMPLED / BPT_ENVIRONMENT_NEE_CDF

    Accel accel;

    // Render state. `accumulation`, `width`, `height` and `half4_scale` are those of the SELECTED accumulation target
    // (bpt_select_accumulation: one per camera, Renderer.cpp:199-222); the others are parked in `parked_targets`.
    struct AccumulationTarget { DeviceBuffer<double> buffer; int width = 0, height = 0; float half4_scale = 1.0f; };
    std::map<int, AccumulationTarget> parked_targets;
    int selected_target = 0;
    int width = 0, height = 0;
    DeviceBuffer<double> accumulation; // double4 per pixel: radiance sum xyz, sample count w
    DeviceBuffer<uint16_t> output_half4; // staging frame for bpt_resolve_half4 to host memory
    DeviceBuffer<float> output_float4;   // the same for bpt_resolve_float4
    DeviceBuffer<unsigned char> query_scratch; // rays in, hits out of bpt_intersect; grows, never shrinks
    // Pipelined frame read-back (bpt_resolve_half4_async): BPT_FRAME_SLOTS staging frames, a copy stream and per-slot events, so the
    // device -> host copy of frame k overlaps the rendering of frame k + 1.
    DeviceBuffer<uint16_t> frame_staging[BPT_FRAME_SLOTS];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t frame_resolved[BPT_FRAME_SLOTS] = {}, frame_copied[BPT_FRAME_SLOTS] = {};
    bool frame_in_flight[BPT_FRAME_SLOTS] = {};
    float half4_scale = 1.0f; // depth backend: the displayed value is depth / (far - near), SimpleRGPs.cu:247-258
    uint64_t material_version = 0;
    bool env_light_uploaded = false;
    void* wavefront = nullptr;         // integrator-owned state (bpt_render.cu)
    uint64_t scene_epoch = 0;          // bumped by every call that uploads or rebuilds scene data on the main stream

    // Multi-GPU (bpt_comm.cu): ncclComm_t of this rank, bound at run time.
    void* comm = nullptr;
    int comm_rank = 0, comm_rank_count = 1;

    bpt_counters counters = {};
    int launches_per_iteration = 5; // kernels per wavefront iteration; the iterations themselves are counted on the device
    uint64_t* device_counters = nullptr; // [extend, shadow]

    cudaEvent_t ev[8] = {};
    // Optional per-stage timing (bpt_set_profiling): events recorded around the stage kernels of one sample.
    bool profiling = false;
    std::vector<cudaEvent_t> stage_events;
    size_t stage_event(size_t index) {
        while (stage_events.size() <= index) { cudaEvent_t e; cudaEventCreate(&e); stage_events.push_back(e); }
        return index;
    }

    int fail(int status, const std::string& msg) { last_error = msg; return status; }
    int cuda_fail(cudaError_t e, const char* what) {
        last_error = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? BPT_ERROR_OUT_OF_MEMORY : BPT_ERROR_CUDA;
    }
};

#define BPT_CUDA_CHECK(ctx, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return (ctx)->cuda_fail(_e, #expr); } while (0)

inline Context* as_context(bpt_ctx* c) { return reinterpret_cast<Context*>(c); }
inline const Context* as_context(const bpt_ctx* c) { return reinterpret_cast<const Context*>(c); }

// Implemented in bpt_bvh.cu
int build_accel(Context* ctx);
int intersect_batch(Context* ctx, int64_t n, const float* origins, const float* directions, const float* tmin, const float* tmax,
                    int32_t* out_primitive, float* out_t, float* out_uv, uint8_t* out_occluded);
// Implemented in bpt_render.cu
int render(Context* ctx, const bpt_camera* camera, const bpt_settings* settings, int width, int height,
           uint32_t first_sample, uint32_t sample_count, int reset_accumulation);
int resolve_half4(Context* ctx, uint16_t* out, int on_device);
// Implemented in bpt_aov.cu
int render_aov(Context* ctx, const bpt_camera* camera, int kind, int width, int height, uint32_t first_sample, uint32_t sample_count, int reset_accumulation);
int resolve_float4(Context* ctx, float* out);
void release_wavefront(Context* ctx);
int tonemap_batch(Context* ctx, const bpt_tonemap_settings* settings, int64_t n, const float* rgb_in, float* rgb_out); // bpt_tonemap.cu
int resolve_half4_async(Context* ctx, uint16_t* out_host, int slot); // bpt_render.cu
int wait_frame(Context* ctx, int slot);
int resolve_tonemapped(Context* ctx, const bpt_tonemap_settings* settings, void* out, int output_format);
int sync_texture_table(Context* ctx); // uploads the texture id -> cudaTextureObject_t table when it changed (bpt_api.cu)

} // namespace bpt
