// Host side of the drop-in: OptiXRenderer::Renderer over the Bifrost core scene handles, driving libbpt.so through the
// C ABI (include/bpt_c_api.h). Mirrors the behaviour of the reference's
// extensions/OptiXRenderer/OptiXRenderer/Renderer.cpp (initialize :1365-1378, handle_updates :578-1205,
// prepare_camera_state + render :1207-1265, setters :1389-1474) but flattens the scene into device arrays instead of an
// OptiX scene graph. Scene synchronisation is incremental at the granularity of the C ABI: meshes and textures are uploaded
// once and stay resident, a material edit uploads the material records only, and a transform or model change re-flattens the
// resident meshes and rebuilds the BVH on the device (the counterpart of the reference's acceleration refit).
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#include <OptiXRenderer/Renderer.h>

#include <optixu/optixpp_namespace.h>

#include <Bifrost/Assets/Image.h>
#include <Bifrost/Assets/InfiniteAreaLight.h>
#include <Bifrost/Assets/Material.h>
#include <Bifrost/Assets/Texture.h>
#include <Bifrost/Math/RNG.h>
#include <Bifrost/Assets/Mesh.h>
#include <Bifrost/Assets/MeshModel.h>
#include <Bifrost/Assets/Shading/Fittings.h>
#include <Bifrost/Core/Renderer.h>
#include <Bifrost/Math/Conversions.h>
#include <Bifrost/Math/FixedPointTypes.h>
#include <Bifrost/Math/half.h>
#include <Bifrost/Scene/Camera.h>
#include <Bifrost/Scene/LightSource.h>
#include <Bifrost/Scene/SceneNode.h>
#include <Bifrost/Scene/SceneRoot.h>

#include "../../include/bpt_c_api.h"

#include <cuda_runtime.h>

#include <climits>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

using namespace Bifrost;
using namespace Bifrost::Assets;
using namespace Bifrost::Math;
using namespace Bifrost::Scene;

// ---- optix facade ------------------------------------------------------------------------------------------------
namespace optix {

static int g_next_buffer_id = 1;

BufferObj::BufferObj(RTformat format, RTsize width, RTsize height)
    : m_format(format), m_width(width), m_height(height), m_element_size(format == RT_FORMAT_HALF4 ? 8 : 16), m_device(nullptr),
      m_owns_device(true), m_id(g_next_buffer_id++) {
    if (cudaMalloc(&m_device, m_width * m_height * m_element_size) != cudaSuccess)
        throw Exception("optix facade: cudaMalloc failed for a render target");
}
BufferObj::~BufferObj() { if (m_owns_device && m_device) cudaFree(m_device); }
void* BufferObj::map() {
    m_host.resize(m_width * m_height * m_element_size);
    cudaMemcpy(m_host.data(), m_device, m_host.size(), cudaMemcpyDeviceToHost);
    return m_host.data();
}
void BufferObj::setDevicePointer(int, void* pointer) {
    if (m_owns_device && m_device) cudaFree(m_device);
    m_device = pointer; m_owns_device = false;
}

} // namespace optix

namespace OptiXRenderer {

static const int MAX_RNG_SAMPLE_OFFSETS = 256; // Renderer.cpp:46

struct Renderer::Implementation {
    bpt_ctx* ctx = nullptr;
    optix::Context context;
    Core::RendererID owning_renderer_ID;

    struct CameraState {
        unsigned int accumulations = 0;
        unsigned int max_accumulation_count = UINT_MAX;
        unsigned int max_bounce_count = 4; // Renderer.cpp:216
        Vector2i frame_size = Vector2i(0, 0);
        Matrix4x4f inverse_view_projection_matrix = {};
        Backend backend = Backend::None;
        bool initialized = false;
    };
    std::vector<CameraState> per_camera_state;
    static constexpr int SCRATCH_ACCUMULATION_SLOT = 0x7fffffff; // request_auxiliary_buffers renders here

    int next_event_sample_count = 3;                         // Renderer.cpp:479
    PathRegularizationSettings path_regularization = { 0.5f, 0.0f }; // Renderer.cpp:482-483
    AIDenoiserFlags AI_denoiser_flags = AIDenoiserFlag::Default;
    bool scene_uploaded = false;

    Implementation(int cuda_device_ID, Core::RendererID renderer_ID) : owning_renderer_ID(renderer_ID) {
        // NB the reference ignores cuda_device_ID and always uses OptiX device 0 (Renderer.cpp:289-291); here it is honoured
        // so one process per GPU can shard samples.
        if (bpt_create(cuda_device_ID, &ctx) != BPT_OK)
            throw optix::Exception("no CUDA device available for the B200 path tracer");
        // Rho / alpha tables: the reference uploads the same arrays as textures (Renderer.cpp:400-466).
        using namespace Assets::Shading;
        if (bpt_set_tables(ctx, Rho::GGX_with_fresnel, Rho::GGX, Estimate_GGX_bounded_VNDF_alpha::alphas) != BPT_OK)
            throw optix::Exception(bpt_last_error(ctx));
        // Dielectric GGX rho (Renderer.cpp:436-466): Vector2f arrays are {total_rho, reflected_rho} float pairs.
        if (bpt_set_dielectric_tables(ctx, &Rho::dielectric_GGX_into_light_medium[0].x, &Rho::dielectric_GGX_into_dense_medium[0].x) != BPT_OK)
            throw optix::Exception(bpt_last_error(ctx));
        context = optix::Context(new optix::ContextObj());
    }
    ~Implementation() { bpt_destroy(ctx); }

    void conditional_per_camera_state_resize(CameraID camera_ID) {
        if (per_camera_state.size() <= camera_ID)
            per_camera_state.resize(Cameras::capacity());
    }

    // set_backend, Renderer.cpp:1411-1453: backend -> entry point.
    static int aov_of(Backend backend) {
        switch (backend) {
        case Backend::DepthVisualization: return BPT_AOV_DEPTH;
        case Backend::AlbedoVisualization: return BPT_AOV_ALBEDO;
        case Backend::TintVisualization: return BPT_AOV_TINT;
        case Backend::RoughnessVisualization: return BPT_AOV_ROUGHNESS;
        case Backend::ShadingNormalVisualization: return BPT_AOV_SHADING_NORMAL;
        case Backend::PrimitiveIdVisualization: return BPT_AOV_PRIMITIVE_ID;
        default: return 0;
        }
    }

    bpt_camera camera_of(CameraID camera_ID) {
        bpt_camera camera;
        Matrix4x4f inverse_projection_matrix = Cameras::get_inverse_projection_matrix(camera_ID);
        Matrix4x4f inverse_view_projection_matrix = Cameras::get_inverse_view_projection_matrix(camera_ID);
        Matrix3x3f view_to_world_rotation = to_matrix3x3(Cameras::get_inverse_view_transform(camera_ID).rotation);
        memcpy(camera.view_to_world_rotation, view_to_world_rotation.begin(), sizeof(camera.view_to_world_rotation));
        memcpy(camera.inverse_projection, inverse_projection_matrix.begin(), sizeof(camera.inverse_projection));
        memcpy(camera.inverse_view_projection, inverse_view_projection_matrix.begin(), sizeof(camera.inverse_view_projection));
        return camera;
    }

    // request_auxiliary_buffers, Renderer.cpp:1267-1358: re-render the requested features into scratch buffers.
    std::vector<Screenshot> request_auxiliary_buffers(CameraID camera_ID, Cameras::ScreenshotContent content_requested, Vector2i frame_size) {
        std::vector<Screenshot> screenshots;
        conditional_per_camera_state_resize(camera_ID);
        unsigned int accumulation_count = per_camera_state[camera_ID].accumulations > 1u ? per_camera_state[camera_ID].accumulations : 1u;
        int pixel_count = frame_size.x * frame_size.y;
        bpt_camera camera = camera_of(camera_ID);
        std::vector<float> mean(4 * size_t(pixel_count));
        // The features are rendered into a scratch accumulation target (Renderer.cpp:1280 allocates scratch buffers), so the
        // camera's own progressive accumulation is left alone.
        check(ctx, bpt_select_accumulation(ctx, SCRATCH_ACCUMULATION_SLOT), "bpt_select_accumulation");
        auto render_feature = [&](int aov) {
            check(ctx, bpt_render_aov(ctx, &camera, aov, frame_size.x, frame_size.y, 0, accumulation_count, 1), "bpt_render_aov");
            check(ctx, bpt_resolve_float4(ctx, mean.data()), "bpt_resolve_float4");
        };
        auto half_round = [](float v) { return float(half_float::half(v)); }; // the reference reads the half4 output buffer back
        if (content_requested.contains(Screenshot::Content::Depth)) {
            render_feature(BPT_AOV_DEPTH);
            float* pixels = new float[pixel_count];
            for (int i = 0; i < pixel_count; ++i) pixels[i] = mean[4 * i];
            screenshots.emplace_back(frame_size.x, frame_size.y, Screenshot::Content::Depth, PixelFormat::Intensity_Float, pixels);
        }
        auto rgb24 = [&](Screenshot::Content content, int aov) {
            render_feature(aov);
            RGB24* pixels = new RGB24[pixel_count];
            for (int i = 0; i < pixel_count; ++i) { pixels[i].r = half_round(mean[4 * i]); pixels[i].g = half_round(mean[4 * i + 1]); pixels[i].b = half_round(mean[4 * i + 2]); }
            screenshots.emplace_back(frame_size.x, frame_size.y, content, PixelFormat::RGB24, pixels);
        };
        if (content_requested.contains(Screenshot::Content::Albedo)) rgb24(Screenshot::Content::Albedo, BPT_AOV_ALBEDO);
        if (content_requested.contains(Screenshot::Content::Tint)) rgb24(Screenshot::Content::Tint, BPT_AOV_TINT);
        if (content_requested.contains(Screenshot::Content::Roughness)) {
            render_feature(BPT_AOV_ROUGHNESS);
            unsigned char* pixels = new unsigned char[pixel_count];
            for (int i = 0; i < pixel_count; ++i) pixels[i] = UNorm8::to_byte(half_round(mean[4 * i]));
            screenshots.emplace_back(frame_size.x, frame_size.y, Screenshot::Content::Roughness, PixelFormat::Intensity8, pixels);
        }
        check(ctx, bpt_release_accumulation(ctx, SCRATCH_ACCUMULATION_SLOT), "bpt_release_accumulation");
        return screenshots;
    }

    static void check(bpt_ctx* ctx, int status, const char* what) {
        if (status != BPT_OK) printf("OptiXRenderer(B200) error in %s: %s\n", what, bpt_last_error(ctx));
    }

    // Images and textures, Renderer.cpp:650-751: every 2D texture becomes a texture object keyed by its TextureID.
    // Returns true when something changed.
    std::map<unsigned int, bool> uploaded_textures;
    // Images and textures, Renderer.cpp:650-751. A texture travels when it is created, and again when the pixels of its image
    // change (Images::Change::PixelsUpdated, :657-669); destroyed textures are collected here and released on the device once
    // the materials that referenced them have been re-uploaded (release_destroyed_textures).
    std::map<unsigned int, int> uploaded_texture_channels; // channel count of the device texture by TextureID index
    std::vector<unsigned int> destroyed_textures;
    bool upload_textures() {
        bool changed = false;
        for (TextureID texture_ID : Textures::get_changed_textures())
            if (Textures::get_changes(texture_ID).contains(Textures::Change::Destroyed) && uploaded_textures[texture_ID]) {
                uploaded_textures[texture_ID] = false;
                uploaded_texture_channels.erase(texture_ID.get_index());
                destroyed_textures.push_back(texture_ID.get_index());
                changed = true; // materials are re-uploaded without it
            }
        for (TextureID texture_ID : Textures::get_iterable()) {
            Image image = Textures::get_image_ID(texture_ID);
            bool created = Textures::get_changes(texture_ID).contains(Textures::Change::Created);
            bool pixels_updated = image.exists() && Images::get_changes(image.get_ID()).contains(Images::Change::PixelsUpdated);
            if (uploaded_textures[texture_ID] && !created && !pixels_updated) continue;
            if (!image.exists() || image.get_depth() > 1) continue;
            bpt_texture_desc desc = {};
            desc.width = int(image.get_width()); desc.height = int(image.get_height());
            int channels = 4;
            switch (image.get_pixel_format()) {
            case PixelFormat::Alpha8: desc.pixel_format = BPT_PIXEL_ALPHA8; channels = 1; break;
            case PixelFormat::RGB24: desc.pixel_format = BPT_PIXEL_RGB24; break;
            case PixelFormat::RGBA32: desc.pixel_format = BPT_PIXEL_RGBA32; break;
            case PixelFormat::RGB_Float: desc.pixel_format = BPT_PIXEL_RGB_FLOAT; break;
            case PixelFormat::RGBA_Float: desc.pixel_format = BPT_PIXEL_RGBA_FLOAT; break;
            default:
                printf("OptiXRenderer(B200) warning: texture %u has a pixel format the renderer does not sample; ignored.\n", texture_ID.get_index());
                continue;
            }
            desc.is_srgb = image.is_sRGB() ? 1 : 0;
            desc.wrap_u = Textures::get_wrapmode_U(texture_ID) == WrapMode::Repeat ? BPT_WRAP_REPEAT : BPT_WRAP_CLAMP;
            desc.wrap_v = Textures::get_wrapmode_V(texture_ID) == WrapMode::Repeat ? BPT_WRAP_REPEAT : BPT_WRAP_CLAMP;
            // rtTex2D reads mip level 0 of a single-level sampler: the magnification filter applies (Renderer.cpp:742-744).
            desc.linear_filter = Textures::get_magnification_filter(texture_ID) == MagnificationFilter::Linear ? 1 : 0;
            int status = bpt_upload_texture(ctx, int(texture_ID.get_index()), &desc, image.get_pixels());
            check(ctx, status, "bpt_upload_texture");
            uploaded_textures[texture_ID] = status == BPT_OK;
            if (status == BPT_OK) uploaded_texture_channels[texture_ID.get_index()] = channels;
            changed = true;
        }
        return changed;
    }

    // After the materials were re-uploaded no device material references a destroyed texture any more.
    void release_destroyed_textures() {
        for (unsigned int texture_index : destroyed_textures)
            check(ctx, bpt_destroy_texture(ctx, int(texture_index)), "bpt_destroy_texture");
        destroyed_textures.clear();
    }

    // A material's texture id if the texture made it to the device, else 0 (untextured).
    // `required_channels`: the tint / roughness texture is sampled as RGBA, the roughness, metallic and coverage textures as
    // one channel (Types.h:388-414). A texture of the wrong kind is dropped from THIS material with a warning instead of
    // making bpt_set_materials refuse the whole array.
    int device_texture_id(TextureID texture_ID, int required_channels, const char* what, const std::string& material_name) {
        if (texture_ID == TextureID::invalid_UID()) return 0;
        auto it = uploaded_textures.find(texture_ID);
        if (it == uploaded_textures.end() || !it->second) return 0;
        if (uploaded_texture_channels[texture_ID.get_index()] != required_channels) {
            printf("OptiXRenderer(B200) warning: the %s texture of material '%s' has %d channel(s), %d needed; rendering the material without it.\n",
                   what, material_name.c_str(), uploaded_texture_channels[texture_ID.get_index()], required_channels);
            return 0;
        }
        return int(texture_ID.get_index());
    }

    // load_mesh, Renderer.cpp:92-136 / mesh updates :621-648: a mesh is uploaded once and stays on the device; only created
    // (or re-created) meshes travel again. Returns true when something changed.
    std::map<unsigned int, bool> uploaded_meshes;
    bool upload_meshes() {
        bool changed = false;
        for (MeshID mesh_ID : Meshes::get_iterable()) {
            if (uploaded_meshes[mesh_ID] && !Meshes::get_changes(mesh_ID).contains(Meshes::Change::Created)) continue;
            static_assert(sizeof(TintRoughness) == 4, "TintRoughness is uchar4");
            int status = bpt_upload_mesh(ctx, int(mesh_ID.get_index()), Meshes::get_indices(mesh_ID), int(Meshes::get_primitive_count(mesh_ID)),
                                         (const float*)Meshes::get_positions(mesh_ID), (const float*)Meshes::get_normals(mesh_ID),
                                         (const float*)Meshes::get_texcoords(mesh_ID), (const uint8_t*)Meshes::get_tint_and_roughness(mesh_ID),
                                         int(Meshes::get_vertex_count(mesh_ID)));
            check(ctx, status, "bpt_upload_mesh");
            if (status == BPT_OK && Meshes::get_emission(mesh_ID) != nullptr) // MeshFlag::Emissive, Renderer.cpp:114,131,154
                check(ctx, bpt_set_mesh_emission(ctx, int(mesh_ID.get_index()), (const float*)Meshes::get_emission(mesh_ID), int(Meshes::get_vertex_count(mesh_ID))),
                      "bpt_set_mesh_emission");
            uploaded_meshes[mesh_ID] = status == BPT_OK;
            changed = true;
        }
        return changed;
    }

    // Meshes destroyed in the core are dropped from the device once no instance references them any more.
    void remove_destroyed_meshes() {
        for (MeshID mesh_ID : Meshes::get_changed_meshes())
            if (Meshes::get_changes(mesh_ID) == Meshes::Change::Destroyed && uploaded_meshes[mesh_ID]) {
                check(ctx, bpt_remove_mesh(ctx, int(mesh_ID.get_index())), "bpt_remove_mesh");
                uploaded_meshes[mesh_ID] = false;
            }
    }

    // Transform + model, Renderer.cpp:1010-1110: object -> world from the node's global transform. The acceleration
    // structure is rebuilt on the device from the resident meshes (no mesh traffic).
    void upload_instances_and_build() {
        std::vector<bpt_instance> instances;
        for (MeshModelID model_ID : MeshModels::get_iterable()) {
            MeshID mesh_ID = MeshModels::get_mesh_ID(model_ID);
            if (!Meshes::has(mesh_ID) || !uploaded_meshes[mesh_ID]) continue;
            Matrix3x4f m = to_matrix3x4(SceneNodes::get_global_transform(MeshModels::get_scene_node_ID(model_ID)));
            bpt_instance inst = {};
            inst.mesh_id = int(mesh_ID.get_index());
            inst.material_id = int(MeshModels::get_material_ID(model_ID).get_index());
            memcpy(inst.to_world, m.begin(), sizeof(inst.to_world));
            instances.push_back(inst);
        }
        check(ctx, bpt_set_instances(ctx, instances.data(), int(instances.size())), "bpt_set_instances");
        check(ctx, bpt_build_accel(ctx), "bpt_build_accel");
    }

    void upload_materials() {
        // upload_material, Renderer.cpp:753-812; index = MaterialID, the invalid material 0 is uploaded as well (:821).
        std::vector<bpt_material> materials(Materials::capacity());
        for (auto& m : materials) { memset(&m, 0, sizeof(m)); m.coverage = 1.0f; }
        for (MaterialID material_ID : Materials::get_iterable()) {
            Assets::Material host = material_ID;
            bpt_material& d = materials[material_ID];
            d.flags = uint16_t(host.get_flags().raw());
            d.shading_model = uint16_t(int(host.get_shading_model()));
            RGB tint = host.get_tint();
            d.tint[0] = tint.r; d.tint[1] = tint.g; d.tint[2] = tint.b;
            d.roughness = host.get_roughness();
            d.specularity = host.get_specularity();
            d.metallic = host.get_metallic();
            d.coat = uint16_t(fminf(fmaxf(host.get_coat(), 0.0f), 1.0f) * 65535.0f + 0.5f);            // UNorm16, Types.h:86
            d.coat_roughness = uint16_t(fminf(fmaxf(host.get_coat_roughness(), 0.0f), 1.0f) * 65535.0f + 0.5f);
            d.coverage = host.is_cutout() ? host.get_cutout_threshold() : host.get_coverage();
            RGB emission = host.get_emission();
            d.emission[0] = emission.r; d.emission[1] = emission.g; d.emission[2] = emission.b;
            // Renderer.cpp:760-806: one texture carries tint (rgb) and roughness (a), or roughness alone.
            const std::string name = host.get_name();
            if (host.has_tint_texture())
                d.tint_roughness_texture_id = device_texture_id(host.get_tint_roughness_texture_ID(), 4, "tint / roughness", name);
            else if (host.has_roughness_texture())
                d.roughness_texture_id = device_texture_id(host.get_tint_roughness_texture_ID(), 1, "roughness", name);
            d.metallic_texture_id = device_texture_id(host.get_metallic_texture_ID(), 1, "metallic", name);
            d.coverage_texture_id = device_texture_id(host.get_coverage_texture_ID(), 1, "coverage", name);
        }
        check(ctx, bpt_set_materials(ctx, materials.data(), int(materials.size())), "bpt_set_materials");
    }

    // PresampledEnvironmentMap, PresampledEnvironmentMap.cpp:19-101: the latlong radiance map, the per pixel solid angle PDF
    // (sans sin theta) of the core's InfiniteAreaLight and a list of light samples drawn from it up front.
    bool upload_environment_map(const float tint[3], Image image, const InfiniteAreaLight& light, unsigned int sample_count = 8192u) {
        const int width = int(image.get_width()), height = int(image.get_height());
        std::vector<float> texels(4ull * width * height);
        for (int y = 0; y < height; ++y)
            for (int x = 0; x < width; ++x) {
                RGBA pixel = image.get_pixel(Vector2ui(x, y));
                float* texel = texels.data() + 4ull * (size_t(y) * width + x);
                texel[0] = pixel.r; texel[1] = pixel.g; texel[2] = pixel.b; texel[3] = pixel.a;
            }

        const bool importance_sample = light.image_integral() >= 0.00001f && sample_count > 0;
        std::vector<float> per_pixel_pdf(1, 0.0f);
        int pdf_width = 1, pdf_height = 1;
        std::vector<bpt_light_sample> samples(1); // a single invalid sample disables next event estimation (:63-64)
        memset(samples.data(), 0, sizeof(bpt_light_sample));
        if (importance_sample) {
            pdf_width = int(light.get_PDF_width()); pdf_height = int(light.get_PDF_height());
            per_pixel_pdf.resize(size_t(pdf_width) * pdf_height);
            InfiniteAreaLightUtils::reconstruct_solid_angle_PDF_sans_sin_theta(light, per_pixel_pdf.data());

            sample_count = max(2u, next_power_of_two(sample_count));
            const int exponent = int(log2(sample_count));
            std::vector<Vector2f> random_numbers(sample_count);
            RNG::fill_progressive_multijittered_bluenoise_samples(random_numbers.data(), random_numbers.data() + sample_count);
            samples.resize(sample_count);
            // Bit-reversed indexing keeps samples of one stratum next to each other in the list (:75-83).
            #pragma omp parallel for schedule(dynamic, 16)
            for (int i = 0; i < int(sample_count); ++i) {
                Assets::LightSample s = light.sample(random_numbers[reverse_bits(unsigned(i)) >> (32 - exponent)]);
                bpt_light_sample& d = samples[i];
                d.radiance[0] = s.radiance.r; d.radiance[1] = s.radiance.g; d.radiance[2] = s.radiance.b;
                d.pdf = s.PDF;
                d.direction_to_light[0] = s.direction_to_light.x; d.direction_to_light[1] = s.direction_to_light.y; d.direction_to_light[2] = s.direction_to_light.z;
                d.distance = s.distance;
            }
        }
        int status = bpt_set_environment(ctx, tint, texels.data(), width, height, per_pixel_pdf.data(), pdf_width, pdf_height, samples.data(), int(samples.size()));
        check(ctx, status, "bpt_set_environment");
        return status == BPT_OK;
    }

    void upload_lights() {
        // Renderer.cpp:852-1008: linearised light list; unknown types become the magenta warning sphere (:892-898).
        std::vector<bpt_light> lights;
        for (LightSourceID light_ID : LightSources::get_iterable()) {
            bpt_light l = {};
            switch (LightSources::get_type(light_ID)) {
            case LightSources::Type::Sphere: {
                Scene::SphereLight host = Scene::LightSource(light_ID);
                Vector3f p = host.get_node().get_global_transform().translation; RGB power = host.get_power();
                l.flags = BPT_LIGHT_SPHERE;
                l.data[0] = power.r; l.data[1] = power.g; l.data[2] = power.b; l.data[3] = p.x; l.data[4] = p.y; l.data[5] = p.z; l.data[6] = host.get_radius();
                break;
            }
            case LightSources::Type::Spot: {
                Scene::SpotLight host = Scene::LightSource(light_ID);
                Transform t = host.get_node().get_global_transform();
                Vector3f p = t.translation, d = t.rotation.forward(); RGB power = host.get_power();
                l.flags = BPT_LIGHT_SPOT;
                l.data[0] = power.r; l.data[1] = power.g; l.data[2] = power.b; l.data[3] = p.x; l.data[4] = p.y; l.data[5] = p.z; l.data[6] = host.get_radius();
                l.data[7] = d.x; l.data[8] = d.y; l.data[9] = d.z; l.data[10] = host.get_cos_angle();
                break;
            }
            case LightSources::Type::Directional: {
                Scene::DirectionalLight host = Scene::LightSource(light_ID);
                Vector3f d = host.get_node().get_global_transform().rotation.forward(); RGB radiance = host.get_radiance();
                l.flags = BPT_LIGHT_DIRECTIONAL;
                l.data[0] = radiance.r; l.data[1] = radiance.g; l.data[2] = radiance.b; l.data[3] = d.x; l.data[4] = d.y; l.data[5] = d.z;
                break;
            }
            default:
                printf("OptiXRenderer warning: Unknown light source type %u on light %u\n", unsigned(LightSources::get_type(light_ID)), light_ID.get_index());
                l.flags = BPT_LIGHT_SPHERE;
                l.data[0] = 100000; l.data[2] = 100000; l.data[6] = 5;
            }
            lights.push_back(l);
        }
        check(ctx, bpt_set_lights(ctx, lights.data(), int(lights.size())), "bpt_set_lights");
    }

    void handle_updates() {
        bool should_reset_accumulations = false;

        // Cameras, Renderer.cpp:581-619
        for (CameraID cam_ID : Cameras::get_changed_cameras()) {
            auto changes = Cameras::get_changes(cam_ID);
            if (changes.contains(Cameras::Change::Destroyed)) {
                if (cam_ID < per_camera_state.size()) per_camera_state[cam_ID] = CameraState();
                check(ctx, bpt_release_accumulation(ctx, int((unsigned int)cam_ID)), "bpt_release_accumulation");
                continue;
            }
            bool uses_this_renderer = owning_renderer_ID == Cameras::get_renderer_ID(cam_ID);
            bool created = uses_this_renderer && changes.is_set(Cameras::Change::Created);
            bool switched = uses_this_renderer && changes.is_set(Cameras::Change::Renderer);
            conditional_per_camera_state_resize(cam_ID);
            CameraState& state = per_camera_state[cam_ID];
            if (!state.initialized && (created || switched)) {
                state.accumulations = 0u;
                state.frame_size = Vector2i(1, 1);
                state.inverse_view_projection_matrix = {};
                if (state.backend == Backend::None) state.backend = Backend::PathTracing;
                state.initialized = true;
            }
        }

        // Incremental scene synchronisation (Renderer.cpp:621-1110): textures and meshes travel only when created, a material
        // edit uploads the 64-byte material records and nothing else, and a transform or model change re-flattens the resident
        // meshes and rebuilds the BVH on the device.
        bool textures_changed = upload_textures();
        bool meshes_changed = upload_meshes();
        bool materials_changed = !scene_uploaded || textures_changed || !Materials::get_changed_materials().is_empty();
        if (materials_changed) upload_materials();
        release_destroyed_textures();
        bool instances_changed = !scene_uploaded || meshes_changed || !Meshes::get_changed_meshes().is_empty();
        instances_changed |= !MeshModels::get_changed_models().is_empty();
        for (SceneNodeID node_ID : SceneNodes::get_changed_nodes())
            if (SceneNodes::get_changes(node_ID).contains(SceneNodes::Change::Transform)) instances_changed = true;
        int64_t triangle_count = 0;
        bool accel_lost = bpt_accel_info(ctx, &triangle_count, nullptr, nullptr) != BPT_OK; // e.g. a first textured material
        if (instances_changed || accel_lost) upload_instances_and_build();
        remove_destroyed_meshes();
        bool geometry_changed = materials_changed || instances_changed || accel_lost;
        if (geometry_changed) should_reset_accumulations = true;

        bool lights_changed = !scene_uploaded || !LightSources::get_changed_lights().is_empty() || geometry_changed;
        if (lights_changed) {
            upload_lights();
            should_reset_accumulations = true;
        }

        // Scene roots, Renderer.cpp:1112-1200: environment tint and map.
        for (SceneRoot scene_data : SceneRoots::get_changed_scenes()) {
            float tint[3] = { 0, 0, 0 };
            bool has_map = false;
            if (!scene_data.get_changes().contains(SceneRoots::Change::Destroyed)) {
                RGB env_tint = scene_data.get_environment_tint();
                tint[0] = env_tint.r; tint[1] = env_tint.g; tint[2] = env_tint.b;
                Texture environment_map = scene_data.get_environment_map();
                if (environment_map.exists()) {
                    Image image = environment_map.get_image();
                    if (channel_count(image.get_pixel_format()) == 4 && scene_data.get_environment_light() != nullptr)
                        has_map = upload_environment_map(tint, image, *scene_data.get_environment_light());
                    else // Renderer.cpp:1151-1160
                        printf("OptiXRenderer only supports environments with 4 channels. '%s' has %u.\n", image.get_name().c_str(), channel_count(image.get_pixel_format()));
                }
            }
            if (!has_map)
                check(ctx, bpt_set_environment(ctx, tint, nullptr, 0, 0, nullptr, 0, 0, nullptr, 0), "bpt_set_environment");
            should_reset_accumulations = true;
        }
        scene_uploaded = true;

        if (should_reset_accumulations)
            for (auto& camera_state : per_camera_state) camera_state.accumulations = 0u;
    }

    unsigned int render(CameraID camera_ID, optix::Buffer buffer, Vector2i frame_size, unsigned int first_sample) {
        conditional_per_camera_state_resize(camera_ID);
        CameraState& state = per_camera_state[camera_ID];

        // prepare_camera_state, Renderer.cpp:1207-1248
        if (frame_size != state.frame_size) { state.frame_size = frame_size; state.accumulations = 0u; }
        Matrix4x4f inverse_projection_matrix = Cameras::get_inverse_projection_matrix(camera_ID);
        Matrix4x4f inverse_view_projection_matrix = Cameras::get_inverse_view_projection_matrix(camera_ID);
        if (state.inverse_view_projection_matrix != inverse_view_projection_matrix) state.accumulations = 0u;
        state.inverse_view_projection_matrix = inverse_view_projection_matrix;
        Matrix3x3f view_to_world_rotation = to_matrix3x3(Cameras::get_inverse_view_transform(camera_ID).rotation);

        if (state.accumulations >= state.max_accumulation_count)
            return state.accumulations;

        bpt_camera camera;
        memcpy(camera.view_to_world_rotation, view_to_world_rotation.begin(), sizeof(camera.view_to_world_rotation));
        memcpy(camera.inverse_projection, inverse_projection_matrix.begin(), sizeof(camera.inverse_projection));
        memcpy(camera.inverse_view_projection, inverse_view_projection_matrix.begin(), sizeof(camera.inverse_view_projection));
        bpt_settings settings = {};
        settings.max_bounce_count = state.max_bounce_count;
        settings.next_event_sample_count = next_event_sample_count;
        settings.path_regularization_pdf_scale = path_regularization.PDF_scale_at_accumulation(int(state.accumulations));

        // One accumulation target per camera, Renderer.cpp:199-222.
        check(ctx, bpt_select_accumulation(ctx, int((unsigned int)camera_ID)), "bpt_select_accumulation");
        int status;
        int aov = aov_of(state.backend);
        if (state.backend == Backend::AIDenoisedPathTracing) {
            printf("OptiXRenderer(B200): Backend %u not supported (proprietary OptiX denoiser).\n", unsigned(state.backend));
            return state.accumulations;
        }
        if (aov == 0)
            status = bpt_render(ctx, &camera, &settings, frame_size.x, frame_size.y, first_sample + state.accumulations, 1, state.accumulations == 0 ? 1 : 0);
        else
            status = bpt_render_aov(ctx, &camera, aov, frame_size.x, frame_size.y, first_sample + state.accumulations, 1, state.accumulations == 0 ? 1 : 0);
        check(ctx, status, "bpt_render");
        if (status == BPT_OK) {
            check(ctx, bpt_resolve_half4(ctx, (uint16_t*)buffer->getDevicePointer(0), 1), "bpt_resolve_half4");
            check(ctx, bpt_synchronize(ctx), "bpt_synchronize");
            ++state.accumulations;
        }
        return state.accumulations;
    }
};

Renderer* Renderer::initialize(int cuda_device_ID, const std::filesystem::path& data_directory) {
    try {
        return new Renderer(cuda_device_ID, data_directory);
    } catch (optix::Exception e) {
        printf("OptiXRenderer failed to initialize:\n%s\n", e.getErrorString().c_str());
        return nullptr;
    }
}

Renderer::Renderer(int cuda_device_ID, const std::filesystem::path&)
    : m_renderer_ID(Core::Renderers::create("OptiXRenderer")), m_impl(nullptr) {
    try {
        m_impl = new Implementation(cuda_device_ID, m_renderer_ID);
    } catch (...) {
        Core::Renderers::destroy(m_renderer_ID);
        throw;
    }
}

Renderer::~Renderer() {
    Core::Renderers::destroy(m_renderer_ID);
    delete m_impl;
}

Backend Renderer::get_backend(CameraID camera_ID) const { m_impl->conditional_per_camera_state_resize(camera_ID); return m_impl->per_camera_state[camera_ID].backend; }
void Renderer::set_backend(CameraID camera_ID, Backend backend) {
    if (backend == Backend::None) return;
    m_impl->conditional_per_camera_state_resize(camera_ID);
    auto& state = m_impl->per_camera_state[camera_ID];
    state.backend = backend;
    state.accumulations = 0u;
}
unsigned int Renderer::get_max_bounce_count(CameraID camera_ID) const { m_impl->conditional_per_camera_state_resize(camera_ID); return m_impl->per_camera_state[camera_ID].max_bounce_count; }
void Renderer::set_max_bounce_count(CameraID camera_ID, unsigned int bounce_count) { m_impl->conditional_per_camera_state_resize(camera_ID); m_impl->per_camera_state[camera_ID].max_bounce_count = bounce_count; }
unsigned int Renderer::get_max_accumulation_count(CameraID camera_ID) const { m_impl->conditional_per_camera_state_resize(camera_ID); return m_impl->per_camera_state[camera_ID].max_accumulation_count; }
void Renderer::set_max_accumulation_count(CameraID camera_ID, unsigned int accumulation_count) { m_impl->conditional_per_camera_state_resize(camera_ID); m_impl->per_camera_state[camera_ID].max_accumulation_count = accumulation_count; }
int Renderer::get_next_event_sample_count(SceneRootID) const { return m_impl->next_event_sample_count; }
void Renderer::set_next_event_sample_count(SceneRootID, int sample_count) { m_impl->next_event_sample_count = sample_count < MAX_RNG_SAMPLE_OFFSETS ? sample_count : MAX_RNG_SAMPLE_OFFSETS; }
PathRegularizationSettings Renderer::get_path_regularization_settings() const { return m_impl->path_regularization; }
void Renderer::set_path_regularization_settings(PathRegularizationSettings settings) { m_impl->path_regularization = settings; }
AIDenoiserFlags Renderer::get_AI_denoiser_flags() const { return m_impl->AI_denoiser_flags; }
void Renderer::set_AI_denoiser_flags(AIDenoiserFlags flags) { m_impl->AI_denoiser_flags = flags; }
void Renderer::handle_updates() { m_impl->handle_updates(); }
unsigned int Renderer::render(CameraID camera_ID, optix::Buffer buffer, Vector2i frame_size) { return m_impl->render(camera_ID, buffer, frame_size, m_first_sample); }
std::vector<Screenshot> Renderer::request_auxiliary_buffers(CameraID camera_ID, Cameras::ScreenshotContent content_requested, Vector2i frame_size) {
    return m_impl->request_auxiliary_buffers(camera_ID, content_requested, frame_size);
}
optix::Context& Renderer::get_context() { return m_impl->context; }

bool Renderer::create_communicator_id(char id[128]) { return bpt_comm_unique_id(id) == BPT_OK; }
bool Renderer::join_communicator(const char id[128], int rank_count, int rank) {
    int status = bpt_comm_init(m_impl->ctx, id, rank_count, rank);
    Implementation::check(m_impl->ctx, status, "bpt_comm_init");
    return status == BPT_OK;
}
bool Renderer::reduce_accumulation(CameraID camera_ID, int root) {
    Implementation::check(m_impl->ctx, bpt_select_accumulation(m_impl->ctx, int((unsigned int)camera_ID)), "bpt_select_accumulation");
    int status = bpt_reduce_accumulation(m_impl->ctx, root);
    Implementation::check(m_impl->ctx, status, "bpt_reduce_accumulation");
    return status == BPT_OK;
}
bool Renderer::resolve_accumulation(CameraID camera_ID, optix::Buffer target) {
    Implementation::check(m_impl->ctx, bpt_select_accumulation(m_impl->ctx, int((unsigned int)camera_ID)), "bpt_select_accumulation");
    int status = bpt_resolve_half4(m_impl->ctx, (uint16_t*)target->getDevicePointer(0), 1);
    Implementation::check(m_impl->ctx, status, "bpt_resolve_half4");
    if (status == BPT_OK) status = bpt_synchronize(m_impl->ctx);
    return status == BPT_OK;
}


namespace {
struct CheckpointHeader { char magic[8]; int32_t width, height; uint32_t accumulations; uint32_t reserved; };
const char CHECKPOINT_MAGIC[8] = { 'B', 'P', 'T', 'A', 'C', 'C', '1', 0 };
}

bool Renderer::save_accumulation(CameraID camera_ID, const std::filesystem::path& file) {
    m_impl->conditional_per_camera_state_resize(camera_ID);
    const Implementation::CameraState& state = m_impl->per_camera_state[camera_ID];
    Implementation::check(m_impl->ctx, bpt_select_accumulation(m_impl->ctx, int((unsigned int)camera_ID)), "bpt_select_accumulation");
    CheckpointHeader header = {};
    memcpy(header.magic, CHECKPOINT_MAGIC, sizeof(header.magic));
    if (bpt_read_accumulation(m_impl->ctx, nullptr, &header.width, &header.height) != BPT_OK || header.width <= 0 || header.height <= 0) return false;
    header.accumulations = state.accumulations;
    std::vector<double> sums(4ull * header.width * header.height);
    int status = bpt_read_accumulation(m_impl->ctx, sums.data(), nullptr, nullptr);
    Implementation::check(m_impl->ctx, status, "bpt_read_accumulation");
    if (status != BPT_OK) return false;
    FILE* f = fopen(file.string().c_str(), "wb");
    if (!f) return false;
    bool ok = fwrite(&header, sizeof(header), 1, f) == 1 && fwrite(sums.data(), sizeof(double), sums.size(), f) == sums.size();
    return fclose(f) == 0 && ok;
}

bool Renderer::load_accumulation(CameraID camera_ID, const std::filesystem::path& file) {
    FILE* f = fopen(file.string().c_str(), "rb");
    if (!f) return false;
    CheckpointHeader header = {};
    std::vector<double> sums;
    bool ok = fread(&header, sizeof(header), 1, f) == 1 && memcmp(header.magic, CHECKPOINT_MAGIC, sizeof(header.magic)) == 0 &&
              header.width > 0 && header.height > 0 && (int64_t)header.width * header.height <= 0x7fffffffll;
    if (ok) {
        sums.resize(4ull * header.width * header.height);
        ok = fread(sums.data(), sizeof(double), sums.size(), f) == sums.size();
    }
    fclose(f);
    if (!ok) return false;
    Implementation::check(m_impl->ctx, bpt_select_accumulation(m_impl->ctx, int((unsigned int)camera_ID)), "bpt_select_accumulation");
    int status = bpt_write_accumulation(m_impl->ctx, header.width, header.height, sums.data());
    Implementation::check(m_impl->ctx, status, "bpt_write_accumulation");
    if (status != BPT_OK) return false;
    // the camera continues from the restored count: frame size and view as they are now, or render() would start over
    m_impl->conditional_per_camera_state_resize(camera_ID);
    Implementation::CameraState& state = m_impl->per_camera_state[camera_ID];
    state.frame_size = Vector2i(header.width, header.height);
    state.inverse_view_projection_matrix = Cameras::get_inverse_view_projection_matrix(camera_ID);
    state.accumulations = header.accumulations;
    return true;
}

} // namespace OptiXRenderer
