/* C entry points of the CPU oracle for traversal and the path integrator (oracle/pt_oracle.cpp).
 * TEST INFRASTRUCTURE: loaded only by tests/, __graft_entry__.smoke() and bench.py's CPU legs (tests/oracle_lib.py binds it).
 * Structs are layout-identical to include/bpt_c_api.h (bpt_instance, bpt_material, bpt_light, bpt_light_sample, bpt_camera, bpt_settings). */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
void* pto_scene_create(void);
void pto_scene_destroy(void* scene);
void pto_scene_add_mesh(void* scene, int mesh_id, const uint32_t* indices, int primitive_count, const float* positions, const float* normals,
                        const uint8_t* tint_roughness, int vertex_count);
void pto_scene_set_mesh_emission(void* scene, int mesh_id, const float* emission, int vertex_count);
void pto_scene_set_mesh_texcoords(void* scene, int mesh_id, const float* texcoords, int vertex_count);
/* pixel_format = BPT_PIXEL_* of include/bpt_c_api.h; the sampler restates the CUDA texture unit's documented arithmetic */
void pto_scene_add_texture(void* scene, int texture_id, int width, int height, int pixel_format, int is_srgb, int wrap_u, int wrap_v,
                           int linear_filter, const void* pixels);
void pto_texture_sample(void* scene, int texture_id, int64_t n, const float* uv, float* out_rgba);
void pto_scene_set_instances(void* scene, const void* instances, int count);
void pto_scene_set_materials(void* scene, const void* materials, int count);
void pto_scene_set_lights(void* scene, const void* lights, int count);
void pto_scene_set_environment(void* scene, const float* tint, const float* texels, int width, int height, const float* pdf, int pdf_width,
                               int pdf_height, const void* samples, int sample_count);
void pto_scene_set_environment_cdfs(void* scene, const float* marginal_cdf, const float* conditional_cdf, int pdf_width, int pdf_height,
                                    int sample_by_cdf);
void pto_scene_build(void* scene);                      /* flatten to world space + median split BVH */
int64_t pto_triangle_count(void* scene);
void pto_world_vertices(void* scene, float* out9);
/* brute != 0: loop over all triangles (THE traversal oracle); otherwise the BVH, which must agree bit for bit */
void pto_intersect(void* scene, int64_t n, const float* origins, const float* directions, const float* tmin, const float* tmax, int brute,
                   int32_t* out_primitive, float* out_t, float* out_uv, uint8_t* out_occluded);
/* restatement of path_tracing_RPG (SimpleRGPs.cu:131-140): adds samples first_sample.. to accum_sum (double4 per pixel) */
void pto_render(void* scene, const void* camera, const void* settings, int width, int height, uint32_t first_sample, uint32_t sample_count,
                int row_begin, int row_end, double* accum_sum, uint64_t* out_counters, int threads);
#ifdef __cplusplus
}
#endif
#endif
