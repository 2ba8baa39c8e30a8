/* bpt_c_api.h - C ABI of the B200 path tracer that replaces Bifrost3D's OptiXRenderer hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types. The host side of
 * the reference (extensions/OptiXRenderer/OptiXRenderer/Renderer.cpp) flattens the Bifrost core
 * scene into an OptiX graph and calls `context->launch(...)`; a maintainer swapping in this library
 * flattens the same scene into the calls below instead (INTEGRATION.md shows the binding).
 * Every entry point cites the reference interface it replaces.
 *
 * Conventions: all functions return 0 on success and a negative bpt_status otherwise;
 * `bpt_last_error(ctx)` returns a static, human readable message for the last failure.
 * Unless a parameter says "device", pointers are HOST pointers and are copied during the call.
 * There is NO CPU fallback: without a CUDA device `bpt_create` fails.
 */
#ifndef BPT_C_API_H
#define BPT_C_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bpt_ctx bpt_ctx;

typedef enum bpt_status {
    BPT_OK = 0,
    BPT_ERROR_INVALID_ARGUMENT = -1,
    BPT_ERROR_CUDA = -2,
    BPT_ERROR_NO_DEVICE = -3,
    BPT_ERROR_OUT_OF_MEMORY = -4,
    BPT_ERROR_NOT_READY = -5
} bpt_status;

/* ---- PODs, layout identical to extensions/OptiXRenderer/OptiXRenderer/Types.h ------------------ */

/* Types.h:353-416 `Material` (64 bytes). A texture id is 0 (none) or the id of a texture uploaded with bpt_upload_texture. */
typedef struct bpt_material {
    uint16_t flags;              /* 1 = ThinWalled, 2 = Cutout */
    uint16_t shading_model;      /* 0 = Default, 1 = Diffuse, 2 = Transmissive (needs bpt_set_dielectric_tables) */
    float tint[3];
    float roughness;
    int32_t tint_roughness_texture_id;
    int32_t roughness_texture_id;
    float specularity;
    float metallic;
    int32_t metallic_texture_id;
    float coverage;
    int32_t coverage_texture_id;
    float emission[3];
    uint16_t coat;               /* UNorm16, Types.h:76-93 */
    uint16_t coat_roughness;     /* UNorm16 */
} bpt_material;

/* Types.h:290-312 `Light` (48 bytes): 11 payload floats + flags. */
enum { BPT_LIGHT_NONE = 0, BPT_LIGHT_SPHERE = 1, BPT_LIGHT_DIRECTIONAL = 2, BPT_LIGHT_ENVIRONMENT = 3,
       BPT_LIGHT_PRESAMPLED_ENVIRONMENT = 4, BPT_LIGHT_SPOT = 5, BPT_LIGHT_TYPE_MASK = 7 };
typedef struct bpt_light {
    /* sphere:      power[3], position[3], radius                       (Types.h:224-228)
     * spot:        power[3], position[3], radius, direction[3], cos_angle (Types.h:230-236)
     * directional: radiance[3], direction[3], pad                      (Types.h:238-242) */
    float data[11];
    uint32_t flags;
} bpt_light;

/* Types.h:210-222 `LightSample` (32 bytes). pdf < 0 encodes a delta light (Types.h:155-204). */
typedef struct bpt_light_sample {
    float radiance[3];
    float pdf;
    float direction_to_light[3];
    float distance;
} bpt_light_sample;

/* One MeshModel: mesh + material + world transform. Renderer.cpp:138-182 (one BLAS + Transform per
 * model there; flattened to world space here). `to_world` is a row-major 3x4 affine matrix; Bifrost
 * transforms are rigid + uniform scale (core/Bifrost/Bifrost/Math/Transform.h:28-35). */
typedef struct bpt_instance {
    int32_t mesh_id;
    int32_t material_id;
    float to_world[12];
} bpt_instance;

/* Types.h:486-501 `CameraStateGPU` minus the buffer ids. Matrices are row-major. */
typedef struct bpt_camera {
    float view_to_world_rotation[9];
    float inverse_projection[16];
    float inverse_view_projection[16];
} bpt_camera;

/* Renderer state read by one render call: Renderer.cpp:216 (max_bounce_count, default 4), :479
 * (next_event_sample_count, default 3), PublicTypes.h:40-45 (path regularisation PDF scale, default 0.5
 * and already multiplied by (1 + decay * accumulations) by the caller). */
typedef struct bpt_settings {
    uint32_t max_bounce_count;
    int32_t next_event_sample_count;
    float path_regularization_pdf_scale;
    /* Russian roulette: 0 = off, which is the reference's behaviour (it has none). n > 0: from the n-th surface interaction
     * of a path on, the path survives with probability clamp(max(throughput), 0.05, 1) and its throughput is divided by it;
     * the decision uses RNG dimension 3 of that bounce (Types.h:422-427 leaves dimensions 3..7 free). */
    uint32_t russian_roulette_start_bounce;
} bpt_settings;

typedef struct bpt_counters {
    uint64_t extend_rays;     /* closest-hit traversals            */
    uint64_t shadow_rays;     /* any-hit (visibility) traversals   */
    uint64_t samples;         /* pixel samples started             */
    uint64_t kernel_launches; /* CUDA kernels launched by the library */
    float extend_ms;          /* CUDA-event time spent in the closest-hit traversal kernel */
    float shadow_ms;
    float shade_ms;
    float other_ms;
    uint64_t extend_node_visits;     /* diagnostics, only filled by builds with -DBPT_TRAVERSAL_STATS */
    uint64_t extend_triangle_tests;
    uint64_t nonfinite_samples;      /* pixel samples whose radiance was NaN / inf: dropped from the accumulation, counted here */
    uint64_t traversal_stack_overflows; /* pushes beyond the traversal stack: the build guarantees 0; anything else is a wrong image */
    uint64_t iterations;             /* wavefront iterations (one closest-hit + one shadow traversal launch each), counted on the device */
} bpt_counters;

/* ---- context ----------------------------------------------------------------------------------- */

/* Renderer::initialize, Renderer.cpp:1365-1378 / Implementation ctor :273-574. */
int bpt_create(int cuda_device, bpt_ctx** out_ctx);
void bpt_destroy(bpt_ctx* ctx);
const char* bpt_last_error(const bpt_ctx* ctx);
/* CUDA stream (cudaStream_t) all work of this context is enqueued on; for event timing by callers. */
void* bpt_stream(bpt_ctx* ctx);

/* Rho / alpha tables that Renderer.cpp:400-466 uploads as textures. 32x32 floats each, row-major, from
 * Bifrost::Assets::Shading::{Rho::GGX_with_fresnel, Rho::GGX, Estimate_GGX_bounded_VNDF_alpha::alphas}
 * (core/Bifrost/Bifrost/Assets/Shading/Fittings.h:58-74). */
int bpt_set_tables(bpt_ctx* ctx, const float* ggx_with_fresnel_rho, const float* ggx_rho, const float* estimate_ggx_alpha);
/* Dielectric GGX rho tables that Renderer.cpp:436-466 uploads as one 3D texture. Each 16x16x16 {total_rho, reflected_rho}
 * pairs (8192 floats), [ior_i_over_o][roughness][cos_theta], from Bifrost::Assets::Shading::Rho::
 * {dielectric_GGX_into_light_medium, dielectric_GGX_into_dense_medium} (Fittings.h:36-46). Required before a
 * Transmissive material is rendered; sampled trilinearly like Rho::sample_dielectric_GGX (DielectricGGXRho.cpp:1087-1110). */
int bpt_set_dielectric_tables(bpt_ctx* ctx, const float* into_light_medium, const float* into_dense_medium);

/* ---- scene upload (Renderer::handle_updates, Renderer.cpp:578-1205) ---------------------------- */

/* Images and textures (Renderer.cpp:650-751). Pixel formats carry the values of Bifrost::Assets::PixelFormat (Image.h:27-38);
 * the ones the reference uploads are accepted. RGB24 / RGB_Float are widened to four channels (alpha 255 / 1) like
 * Renderer.cpp:683-694. Texel (x, y) is pixels[y * width + x]; texcoord (0, 0) is the corner of texel (0, 0).
 * Sampling follows the OptiX sampler the reference creates (:729-747): normalized coordinates, one mip level, hardware
 * bilinear (linear_filter != 0, from MagnificationFilter::Linear) or nearest filtering, sRGB decode of 8-bit texels before
 * filtering when is_srgb != 0 (RT_TEXTURE_READ_NORMALIZED_FLOAT_SRGB), wrap 0 = clamp to edge, 1 = repeat. */
enum { BPT_PIXEL_ALPHA8 = 1, BPT_PIXEL_RGB24 = 3, BPT_PIXEL_RGBA32 = 4, BPT_PIXEL_RGB_FLOAT = 6, BPT_PIXEL_RGBA_FLOAT = 7 };
enum { BPT_WRAP_CLAMP = 0, BPT_WRAP_REPEAT = 1 };
typedef struct bpt_texture_desc {
    int32_t width, height;
    int32_t pixel_format;   /* BPT_PIXEL_* */
    int32_t is_srgb;
    int32_t wrap_u, wrap_v; /* BPT_WRAP_* */
    int32_t linear_filter;
    int32_t reserved;
} bpt_texture_desc;
/* texture_id >= 1 (0 means "no texture" in bpt_material). Re-uploading an id replaces the texture. Upload textures before the
 * materials that reference them. */
int bpt_upload_texture(bpt_ctx* ctx, int texture_id, const bpt_texture_desc* desc, const void* pixels);
int bpt_destroy_texture(bpt_ctx* ctx, int texture_id);
/* rtTex2D<float4>(texture, u, v) for n texcoords (uv: 2n floats, out_rgba: 4n floats); unit entry point for the parity tests. */
int bpt_texture_sample(bpt_ctx* ctx, int texture_id, int64_t n, const float* uv, float* out_rgba);

/* load_mesh, Renderer.cpp:92-136. indices: 3*primitive_count uint32; positions: 3*vertex_count floats;
 * normals (nullable): 3*vertex_count floats, octahedral-encoded to short2 like OctahedralNormal::encode_precise;
 * texcoords (nullable): 2*vertex_count floats; tint_roughness (nullable): 4*vertex_count bytes. */
int bpt_upload_mesh(bpt_ctx* ctx, int mesh_id, const uint32_t* indices, int primitive_count,
                    const float* positions, const float* normals, const float* texcoords,
                    const uint8_t* tint_roughness, int vertex_count);
/* Per-vertex emission scale of a mesh (MeshFlag::Emissive, Renderer.cpp:114,131; TriangleAttributes.cu:78-83): 3*vertex_count
 * floats, interpolated over the triangle and multiplied with the material's emission. nullptr removes it (scale 1). */
int bpt_set_mesh_emission(bpt_ctx* ctx, int mesh_id, const float* emission, int vertex_count);
/* Meshes::Change::Destroyed, Renderer.cpp:628-640. Fails while an instance still references the mesh. */
int bpt_remove_mesh(bpt_ctx* ctx, int mesh_id);
/* Renderer.cpp:1043-1110 (mesh models) + :1010-1041 (transforms). Replaces all instances. Meshes stay resident on the
 * device, so a transform-only change costs bpt_set_instances + bpt_build_accel (device-side re-flatten and rebuild:
 * the counterpart of the reference's acceleration refit, Renderer.cpp:470-477) and no mesh traffic. */
int bpt_set_instances(bpt_ctx* ctx, const bpt_instance* instances, int count);
/* upload_material, Renderer.cpp:753-850. Index = MaterialID; index 0 is the invalid material. Materials are read at
 * shading time: changing them does not invalidate the acceleration structure (except when a first textured material makes
 * the per-primitive texcoords necessary). */
int bpt_set_materials(bpt_ctx* ctx, const bpt_material* materials, int count);
/* Renderer.cpp:852-1008. Sphere, spot and directional lights. */
int bpt_set_lights(bpt_ctx* ctx, const bpt_light* lights, int count);
/* Scene root environment, Renderer.cpp:1112-1200, PresampledEnvironmentMap.cpp:19-101.
 * texels: width*height RGBA float latlong map (nullable -> constant `tint`); per_pixel_pdf: pdf_width*pdf_height
 * floats (solid angle PDF sans sin theta, InfiniteAreaLight.cpp:140-157); samples: `sample_count` presampled
 * light samples (PresampledEnvironmentMap.cpp:60-96). */
int bpt_set_environment(bpt_ctx* ctx, const float tint[3], const float* texels, int width, int height,
                        const float* per_pixel_pdf, int pdf_width, int pdf_height,
                        const bpt_light_sample* samples, int sample_count);
/* The 2-D distribution of the environment map for importance sampling by CDF inversion ON THE DEVICE
 * (Shading/LightSources/EnvironmentLightImpl.h:22-83, the `EnvironmentLight` of Types.h:244-271): marginal_cdf holds
 * pdf_height + 1 floats, conditional_cdf pdf_height rows of pdf_width + 1 floats, normalised as Distribution2D::compute_CDFs
 * leaves them (core/Bifrost/Bifrost/Math/Distribution2D.h:172-207; the host reads them from InfiniteAreaLight::
 * get_image_marginal_CDF / get_image_conditional_CDF, InfiniteAreaLight.h:66-69). Call after bpt_set_environment (which
 * drops the CDFs of the previous map); pdf_width x pdf_height must equal the per pixel PDF's size. NULL, NULL removes them. */
int bpt_set_environment_cdfs(bpt_ctx* ctx, const float* marginal_cdf, const float* conditional_cdf, int pdf_width, int pdf_height);
/* How next event estimation samples the environment. PRESAMPLED (default) is what the reference's renderer does
 * (Renderer.cpp:1180-1195, PresampledEnvironmentLightImpl.h:22-27): pick one of the presampled lights. CDF inverts the 2-D
 * CDF per sample (Light::Environment, LightImpl.h:38-52) and needs bpt_set_environment_cdfs; without CDFs the environment
 * then yields no light samples, like EnvironmentLightImpl.h:70. BSDF-sampled rays that escape are evaluated the same way
 * (texel + per pixel PDF) in both modes. */
enum { BPT_ENVIRONMENT_NEE_PRESAMPLED = 0, BPT_ENVIRONMENT_NEE_CDF = 1 };
int bpt_set_environment_sampling(bpt_ctx* ctx, int mode);
/* Sorting of the surface hits before shading. From wavefront iteration `from_iteration` of a sample on (0 = from the camera rays'
 * hits on, 1 = from the first bounce on, < 0 = never) the hits are put in the order of a 12-bit key - the material's shading
 * class (shading model, coat) and the cell of the hit triangle in the Morton order of the acceleration structure - by a
 * counting sort on the device. Results do not depend on it (path state is indexed by pixel); it only changes which paths
 * share a warp in the shading kernel and, through the order the new rays are queued in, in the following traversal kernels. */
int bpt_set_hit_sorting(bpt_ctx* ctx, int from_iteration);
/* Acceleration structure build; replaces OptiX Trbvh (Renderer.cpp:161-182,470-477). Flattens the
 * instances to world space, builds the LBVH on the device. Must be called after meshes/instances change. */
int bpt_build_accel(bpt_ctx* ctx);
/* Number of triangles / BVH nodes of the last build and its device time in ms. */
int bpt_accel_info(bpt_ctx* ctx, int64_t* triangle_count, int64_t* node_count, float* build_ms);
/* Which node format the rays of the last build traverse: *kind = 8 compressed eight-wide nodes (80 bytes, quantised child
 * boxes), 4 = four-wide nodes (128 bytes), 2 = the binary nodes (64 bytes); their number and the depth of that tree. The
 * build takes the eight-wide nodes from 131 072 triangles on (below that the four-wide ones measure faster) and falls back
 * to the next narrower format when the tree is too deep for a format's traversal stack; BPT_CW=1 / BPT_CW=0 / BPT_WIDE=0 in
 * the environment force or rule out formats (A/B measurements, tests). OptiX keeps this choice to itself (rtAccelerationSetBuilder "Trbvh",
 * Renderer.cpp:161-182), so there is no reference counterpart. */
int bpt_accel_hierarchy(bpt_ctx* ctx, int* kind, int64_t* node_count, int* levels);

/* ---- rendering (Renderer::render, Renderer.cpp:1250-1265; SimpleRGPs.cu:74-140) ----------------- */

/* Renders `sample_count` progressive samples with accumulation indices first_sample .. first_sample+sample_count-1
 * for every pixel of a width x height frame into the context-owned accumulation buffer.
 * The buffer holds the per-pixel radiance SUM as double4 (w = number of samples); it is reset when
 * `reset_accumulation` != 0 or the frame size changes (Renderer.cpp:1207-1248). */
int bpt_render(bpt_ctx* ctx, const bpt_camera* camera, const bpt_settings* settings, int width, int height,
               uint32_t first_sample, uint32_t sample_count, int reset_accumulation);
/* First-hit feature ("AOV") backends; values mirror EntryPoints in Types.h:33-44. Replaces depth_RPG, albedo_RPG, tint_RPG,
 * roughness_RPG, shading_normal_RPG, primitive_id_RPG (SimpleRGPs.cu:227-340). Accumulates like bpt_render; the depth
 * backend accumulates the raw distance (the caller divides by far - near like SimpleRGPs.cu:247-258). */
enum { BPT_AOV_DEPTH = 3, BPT_AOV_ALBEDO = 4, BPT_AOV_TINT = 5, BPT_AOV_ROUGHNESS = 6, BPT_AOV_SHADING_NORMAL = 7, BPT_AOV_PRIMITIVE_ID = 8 };
int bpt_render_aov(bpt_ctx* ctx, const bpt_camera* camera, int aov_kind, int width, int height,
                   uint32_t first_sample, uint32_t sample_count, int reset_accumulation);
/* Accumulation targets. The reference keeps one accumulation buffer per camera (Renderer.cpp:199-222) and renders auxiliary
 * screenshots into scratch buffers (:1280). A context owns any number of targets, addressed by a caller-chosen slot >= 0
 * (the host shim uses the CameraID); bpt_render, bpt_render_aov, bpt_resolve_*, bpt_accumulation_device_ptr and
 * bpt_reduce_accumulation act on the SELECTED target. Slot 0 is selected after bpt_create; a slot is created (empty) the first
 * time it is selected. Selecting is host-side bookkeeping only: no device work, no synchronisation. */
int bpt_select_accumulation(bpt_ctx* ctx, int slot);
/* Frees the device memory of a target (waits for the stream). Releasing the selected slot leaves it selected and empty. */
int bpt_release_accumulation(bpt_ctx* ctx, int slot);
/* Device pointer to the double4[width*height] accumulation (sum) buffer, e.g. for an NCCL reduce. */
void* bpt_accumulation_device_ptr(bpt_ctx* ctx);
/* Checkpoint / resume of a progressive render. The resumable state of the reference is its accumulation buffer and the
 * `accumulations` count (Renderer.cpp:200-205,1262); here it is the selected target's double4 sums (xyz = radiance sum,
 * w = samples) and its size. A sample is a pure function of (pixel, accumulation index, scene) (Types.h:452-459), so a
 * render resumed from a saved state with first_sample = the number of samples it holds continues bit for bit as if it had
 * never stopped. bpt_read_accumulation waits for the rendering enqueued so far and copies 4 * width * height doubles to the
 * host (out_width / out_height report the size; pass sums = NULL to query it); bpt_write_accumulation replaces the selected
 * target with the given state. */
int bpt_read_accumulation(bpt_ctx* ctx, double* sums, int* out_width, int* out_height);
int bpt_write_accumulation(bpt_ctx* ctx, int width, int height, const double* sums);
/* ---- multi-GPU: sample-index sharding -------------------------------------------------------------
 * The reference renders on one device (Renderer.cpp:289-291). Here every rank (one process per GPU) uploads the same scene,
 * renders a disjoint range of accumulation indices (`first_sample`) into its own fp64 sum buffer, and ONE collective adds the
 * buffers: sum and sample count (w) both add, so bpt_resolve_* on the root yields the mean over all ranks' samples.
 * NCCL is loaded at run time (libnccl.so.2; BPT_NCCL_LIB overrides the path); single-GPU hosts never touch it.
 *   rank 0:    bpt_comm_unique_id(id); send `id` to the other ranks by any means (MPI, socket, file, torch.distributed)
 *   all ranks: bpt_comm_init(ctx, id, rank_count, rank); ... bpt_render(...) ...; bpt_reduce_accumulation(ctx, 0);
 *   root:      bpt_resolve_half4(...)                                                                                 */
enum { BPT_COMM_ID_BYTES = 128 }; /* sizeof(ncclUniqueId) */
int bpt_comm_unique_id(char out_id[BPT_COMM_ID_BYTES]);
int bpt_comm_init(bpt_ctx* ctx, const char id[BPT_COMM_ID_BYTES], int rank_count, int rank);
int bpt_comm_destroy(bpt_ctx* ctx);
/* Failure detection: BPT_OK, or the asynchronous error NCCL has recorded on the communicator (a peer that died, a link error;
 * ncclCommGetAsyncError). A collective is only enqueued by bpt_reduce_accumulation, so a host that wants to know whether the
 * combine went through calls this after bpt_synchronize. BPT_OK without a communicator. */
int bpt_comm_check(bpt_ctx* ctx);
/* ncclReduce(sum, fp64) of the selected accumulation target into rank `root`'s, in place, enqueued on the context's stream
 * after the samples already enqueued (no host synchronisation); root < 0: ncclAllReduce, every rank gets the total.
 * All ranks must call it with the same frame size. Call it once per job: the root's buffer then already holds the total. */
int bpt_reduce_accumulation(bpt_ctx* ctx, int root);
/* mean = sum / w, converted to half4 (alpha 1) exactly like SimpleRGPs.cu:39-42,106; written to `out` which is
 * a HOST pointer to width*height*4 uint16 (on_device == 0) or a DEVICE pointer (on_device != 0). */
int bpt_resolve_half4(bpt_ctx* ctx, uint16_t* out, int on_device);
/* Pipelined read-back for progressive display: enqueues the half4 resolve on the render stream and its device -> host copy
 * on a second stream, then returns; the copy of frame k overlaps the rendering of frame k + 1. `out_host` (width*height*4
 * uint16, ideally pinned) is complete after bpt_wait_frame(ctx, slot). BPT_FRAME_SLOTS slots rotate (as many frames can be on
 * their way to the host as samples can be in flight on the device); re-using a slot waits for its previous copy on the device,
 * not on the host. */
enum { BPT_FRAME_SLOTS = 4 };
int bpt_resolve_half4_async(bpt_ctx* ctx, uint16_t* out_host, int slot);
int bpt_wait_frame(bpt_ctx* ctx, int slot);
/* mean as float4 to a HOST buffer of width*height*4 floats. */
int bpt_resolve_float4(bpt_ctx* ctx, float* out);
/* Tonemapped resolve. Operators and parameters are the core's camera effects (core/Bifrost/Bifrost/Math/CameraEffects.h:
 * TonemappingMode :18, TonemappingSettings :21-32 with ACES() = {0, 0.53, 0.91, 0.23, 0.035}, filmic :161-224, agx :236-265,
 * khronos_neutral_tone_mapping :272-291); `exposure` is a linear scale applied first (ExposureMode::Fixed). The filmic
 * parameters are only read by BPT_TONEMAP_FILMIC. */
enum { BPT_TONEMAP_LINEAR = 0, BPT_TONEMAP_FILMIC = 1, BPT_TONEMAP_AGX = 2, BPT_TONEMAP_KHRONOS_NEUTRAL = 3 };
enum { BPT_OUTPUT_FLOAT4 = 0, BPT_OUTPUT_SRGB_RGBA8 = 1 };
typedef struct bpt_tonemap_settings {
    int32_t mode;
    float exposure;
    float black_clip, toe, slope, shoulder, white_clip;
    int32_t reserved;
} bpt_tonemap_settings;
/* mean radiance -> exposure -> operator, to a HOST buffer: width*height float4 (linear, alpha 1) or width*height RGBA8 with
 * the sRGB transfer function applied (Color.h:372-377). */
int bpt_resolve_tonemapped(bpt_ctx* ctx, const bpt_tonemap_settings* settings, void* out, int output_format);
/* the operator alone for n colours (rgb_in, rgb_out: 3n floats): unit entry point for the parity test */
int bpt_tonemap_colors(bpt_ctx* ctx, const bpt_tonemap_settings* settings, int64_t n, const float* rgb_in, float* rgb_out);
/* Image comparison, extensions/ImageOperations/ImageOperations/Compare.h: rms (:23-44: root mean square of the luminance of the
 * per channel absolute difference), ssim (:88-118: structural similarity of the whole images, luminance of the per channel
 * index) and mssim (:123-180: mean of the SSIM of the weighted window [p - support, p + support) around every pixel).
 * reference_rgba / target_rgba: width*height float4 on the HOST. Nullable outputs are skipped; out_rms_diff_rgba receives the
 * per channel absolute difference (the `diff` image of rms), out_mssim_diff_rgba 1 - SSIM per channel (the one of mssim). */
int bpt_compare_images(bpt_ctx* ctx, int width, int height, const float* reference_rgba, const float* target_rgba, int mssim_support,
                       float* out_rms, float* out_ssim, float* out_mssim, float* out_rms_diff_rgba, float* out_mssim_diff_rgba);
int bpt_synchronize(bpt_ctx* ctx);
/* enabled != 0: bpt_render brackets its stage kernels with CUDA events on the context's stream and accumulates their
 * durations into bpt_counters.{extend,shade,shadow}_ms (used by bench.py for the roofline figures). */
int bpt_set_profiling(bpt_ctx* ctx, int enabled);
int bpt_get_counters(bpt_ctx* ctx, bpt_counters* out, int reset);

/* ---- batched unit entry points (parity tests and the C1 workload) ------------------------------ */

enum { BPT_BSDF_DEFAULT_SHADING = 0, BPT_BSDF_GGX_R = 1, BPT_BSDF_OREN_NAYAR = 2, BPT_BSDF_BURLEY = 3,
       BPT_BSDF_TRANSMISSIVE_SHADING = 4, BPT_BSDF_GGX = 5 };
/* evaluate_with_PDF(wo, wi) and sample(wo, u) for n tuples. wo, wi, tint, u: 3n floats; rms: {roughness, metallic,
 * specularity} 3n floats; coat (nullable): {coat, coat_roughness} 2n floats.
 * BPT_BSDF_TRANSMISSIVE_SHADING (TransmissiveShading.h:22-98): rms = {roughness, cos_theta_o, specularity}, where the
 * sign of cos_theta_o says whether the path enters (>= 0) or leaves the medium, as in MonteCarlo.cu:190-191.
 * BPT_BSDF_GGX (combined reflection + transmission, GGX.h:258-443): tint = transmission tint,
 * rms = {roughness, ior_i_over_o, specularity}.
 * Outputs: eval_f 3n, eval_pdf n, sample_f 3n, sample_pdf n, sample_dir 3n.
 * on_device != 0: all pointers are device pointers and nothing is copied. */
int bpt_bsdf_eval_sample_pdf(bpt_ctx* ctx, int kind, int64_t n, const float* wo, const float* wi, const float* tint,
                             const float* rms, const float* coat, const float* u,
                             float* eval_f, float* eval_pdf, float* sample_f, float* sample_pdf, float* sample_dir,
                             int on_device);
/* DefaultShading as the renderer constructs it (per-vertex scale + path regularisation), for n elements. */
int bpt_default_shading_regularized(bpt_ctx* ctx, int64_t n, const bpt_material* materials, const float* tint_roughness_scale,
                                    const float* max_pdf_hint, const float* wo, const float* wi, const float* u,
                                    float* eval_f, float* eval_pdf, float* sample_f, float* sample_pdf, float* sample_dir);
/* sample_radiance / pdf / evaluate for n (light, position) pairs; light_stride 0 broadcasts lights[0]. */
int bpt_light_sample_pdf_evaluate(bpt_ctx* ctx, int64_t n, const bpt_light* lights, int light_stride, const float* position,
                                  const float* u2, const float* query_direction,
                                  bpt_light_sample* out_samples, float* out_pdf, float* out_radiance);
/* PracticalScrambledSobol::sample4ui / sample4f (RNG.h:280-292) for n (accumulation, pixel_hash, dimension) triples. */
int bpt_rng_sample4(bpt_ctx* ctx, int64_t n, const uint32_t* accumulation, const uint32_t* pixel_hash, const uint32_t* dimension,
                    uint32_t* out_ui4, float* out_f4);
/* Closest-hit and any-hit queries against the built acceleration structure for n rays.
 * origins/directions: 3n floats; tmin/tmax: n floats. out_primitive: global primitive index (instance-major) or -1;
 * out_t: hit distance; out_uv: barycentrics (2n); out_occluded: n bytes (any hit in [tmin, tmax]). Nullable outputs are skipped. */
int bpt_intersect(bpt_ctx* ctx, int64_t n, const float* origins, const float* directions, const float* tmin, const float* tmax,
                  int32_t* out_primitive, float* out_t, float* out_uv, uint8_t* out_occluded);

/* The device-wide primitives of the acceleration structure build (which replaces OptiX' Trbvh, Renderer.cpp:161-182) and of the
 * queue reordering, exposed for bit-exact tests against a host sort / cumulative sum. bpt_sort_pairs: stable least-significant-
 * digit radix sort of n (64-bit key, 32-bit value) pairs by key bits [begin_bit, end_bit), in place in the host arrays.
 * bpt_exclusive_scan: out[i] = in[0] + ... + in[i - 1] (wrapping), *out_total = sum of all n. */
int bpt_sort_pairs(bpt_ctx* ctx, int64_t n, uint64_t* keys, uint32_t* values, int begin_bit, int end_bit);
int bpt_exclusive_scan(bpt_ctx* ctx, int64_t n, const uint32_t* in, uint32_t* out, uint32_t* out_total);

#ifdef __cplusplus
}
#endif

#endif /* BPT_C_API_H */
