// Runs the REFERENCE's own renderer integration tests (tests/OptiXRendererTests/RendererTest.h, staged unmodified in
// baseline/_ref) against the B200 drop-in: the fixture creates the scene through the Bifrost core handles, calls
// OptiXRenderer::Renderer::initialize / handle_updates / render and maps the half4 render target.
#include <gtest/gtest.h>

#include <Utils.h>

#include <Bifrost/Assets/Image.h>
#include <Bifrost/Assets/Material.h>
#include <Bifrost/Assets/Mesh.h>
#include <Bifrost/Assets/MeshModel.h>
#include <Bifrost/Assets/Texture.h>
#include <Bifrost/Core/Renderer.h>
#include <Bifrost/Scene/Camera.h>
#include <Bifrost/Scene/LightSource.h>
#include <Bifrost/Scene/SceneNode.h>
#include <Bifrost/Scene/SceneRoot.h>

#include <cuda_fp16.h>
#include <functional>

#include <RendererTest.h>

// ---- additional cases for the rows the reference's fixture does not reach: textures and environment maps ----
namespace OptiXRenderer {

// A quad that fills the orthographic frame with texcoords (0,0)..(1,1) and a nearest-filtered RGBA32 tint texture with one
// texel per pixel: the TintVisualization backend must show the texels (Material::get_tint_roughness, Types.h:388-396).
TEST_F(RendererFixture, b200_render_tint_texture) {
    using namespace Bifrost;
    using namespace Bifrost::Assets;
    using namespace Bifrost::Scene;

    auto size = Math::Vector2i(4, 3);
    CameraID camera_ID = create_ortho_camera(size);
    SceneRoot root = Cameras::get_scene_ID(camera_ID);

    Mesh mesh = Mesh("Quad", 2, 4, { MeshFlag::Position, MeshFlag::Texcoord });
    mesh.get_primitives()[0] = { 0, 1, 2 };
    mesh.get_primitives()[1] = { 1, 2, 3 };
    mesh.get_positions()[0] = { -0.5f * size.x, -0.5f * size.y, 1.0f };
    mesh.get_positions()[1] = { -0.5f * size.x, 0.5f * size.y, 1.0f };
    mesh.get_positions()[2] = { 0.5f * size.x, -0.5f * size.y, 1.0f };
    mesh.get_positions()[3] = { 0.5f * size.x, 0.5f * size.y, 1.0f };
    mesh.get_texcoords()[0] = { 0.0f, 0.0f };
    mesh.get_texcoords()[1] = { 0.0f, 1.0f };
    mesh.get_texcoords()[2] = { 1.0f, 0.0f };
    mesh.get_texcoords()[3] = { 1.0f, 1.0f };

    Image image = Image::create2D("Tint", PixelFormat::RGBA32, false, Math::Vector2ui(size.x, size.y));
    for (int y = 0; y < size.y; ++y)
        for (int x = 0; x < size.x; ++x)
            image.set_pixel(Math::RGBA(x / float(size.x - 1), y / float(size.y - 1), 0.5f, 1.0f), Math::Vector2ui(x, y));
    Texture texture = Texture::create2D(image, MagnificationFilter::None, MinificationFilter::None, WrapMode::Clamp, WrapMode::Clamp);

    auto material = Bifrost::Assets::Material::create_dielectric("Material", Math::RGB::white(), 0.0f);
    material.set_flags(MaterialFlag::ThinWalled);
    material.set_shading_model(ShadingModel::Diffuse);
    material.set_tint_roughness_texture(texture);

    SceneNode node = SceneNode("Node");
    node.set_parent(root.get_root_node());
    MeshModel(node, mesh, material);

    render(camera_ID, size, Backend::TintVisualization, [=](half4* pixels) {
        for (int y = 0; y < size.y; ++y)
            for (int x = 0; x < size.x; ++x) {
                half4 pixel = pixels[x + y * size.x];
                EXPECT_FLOAT_EQ_EPS(x / float(size.x - 1), float(pixel.r), 0.003f) << " at pixel (" << x << ", " << y << ")";
                EXPECT_FLOAT_EQ_EPS(y / float(size.y - 1), float(pixel.g), 0.003f) << " at pixel (" << x << ", " << y << ")";
                EXPECT_FLOAT_EQ_EPS(0.5f, float(pixel.b), 0.003f) << " at pixel (" << x << ", " << y << ")";
            }
    });
}

// An empty scene under a constant RGBA_Float environment map: every pixel shows map * tint (miss program,
// SimpleRGPs.cu:349-362), with the map going Texture -> InfiniteAreaLight -> presampled light list on the way.
TEST_F(RendererFixture, b200_render_environment_map) {
    using namespace Bifrost;
    using namespace Bifrost::Assets;
    using namespace Bifrost::Scene;

    auto size = Math::Vector2i(8, 6);
    Image image = Image::create2D("Environment", PixelFormat::RGBA_Float, false, Math::Vector2ui(16, 8));
    for (unsigned int i = 0; i < 16 * 8; ++i)
        image.set_pixel(Math::RGBA(0.5f, 1.0f, 2.0f, 1.0f), i);
    Texture environment = Texture::create2D(image, MagnificationFilter::Linear, MinificationFilter::Linear, WrapMode::Repeat, WrapMode::Clamp);
    SceneRoot scene = SceneRoot("Test", environment, Math::RGB(0.5f, 0.5f, 0.25f));

    Math::Matrix4x4f orthographic_matrix, inverse_orthographic_matrix;
    CameraUtils::compute_orthographic_projection(float(size.x), float(size.y), 1000.0f, orthographic_matrix, inverse_orthographic_matrix);
    CameraID camera_ID = Cameras::create("Test", scene.get_ID(), orthographic_matrix, inverse_orthographic_matrix);
    Cameras::set_renderer_ID(camera_ID, renderer->get_renderer_ID());

    render(camera_ID, size, [=](half4* pixels) {
        for (int i = 0; i < size.x * size.y; ++i) {
            EXPECT_FLOAT_EQ_EPS(0.25f, float(pixels[i].r), 1e-3f);
            EXPECT_FLOAT_EQ_EPS(0.5f, float(pixels[i].g), 1e-3f);
            EXPECT_FLOAT_EQ_EPS(0.5f, float(pixels[i].b), 1e-3f);
        }
    });
}

// Incremental updates: moving the quad's scene node out of the view re-flattens the resident mesh (no mesh upload) and the
// next frame shows the background; a material edit alone changes the tint without touching the geometry.
TEST_F(RendererFixture, b200_incremental_transform_and_material_updates) {
    using namespace Bifrost;
    using namespace Bifrost::Assets;
    using namespace Bifrost::Scene;

    auto size = Math::Vector2i(4, 3);
    CameraID camera_ID = create_ortho_camera_with_quad_scene(size, optix::make_float3(0.25f, 0.5f, 0.75f));
    render(camera_ID, size, Backend::TintVisualization, [=](half4* pixels) {
        EXPECT_FLOAT_EQ_EPS(0.125f, float(pixels[0].r), 0.003f); // vertex tints, as in render_tint
    });

    auto reset_changes = []() {
        Images::reset_change_notifications(); Textures::reset_change_notifications(); Materials::reset_change_notifications();
        Meshes::reset_change_notifications(); MeshModels::reset_change_notifications(); SceneNodes::reset_change_notifications();
        SceneRoots::reset_change_notifications(); Cameras::reset_change_notifications(); LightSources::reset_change_notifications();
    };
    reset_changes();

    // material edit only
    for (MaterialID material_ID : Materials::get_iterable())
        Materials::set_tint(material_ID, Math::RGB(0.5f, 0.5f, 0.5f));
    render(camera_ID, size, Backend::TintVisualization, [=](half4* pixels) {
        EXPECT_FLOAT_EQ_EPS(0.0625f, float(pixels[0].r), 0.003f); // material tint 0.5 x vertex tint 0.125
    });
    reset_changes();

    // transform edit only: the quad leaves the frustum
    for (MeshModelID model_ID : MeshModels::get_iterable()) {
        SceneNode node = MeshModels::get_scene_node_ID(model_ID);
        Math::Transform t = node.get_global_transform();
        t.translation.x += 100.0f;
        node.set_global_transform(t);
    }
    render(camera_ID, size, [=](half4* pixels) {
        for (int i = 0; i < size.x * size.y; ++i) {
            EXPECT_FLOAT_EQ_EPS(0.25f, float(pixels[i].r), 1e-3f);
            EXPECT_FLOAT_EQ_EPS(0.75f, float(pixels[i].b), 1e-3f);
        }
    });
}

// One accumulation buffer per camera (Renderer.cpp:199-222): two cameras of different sizes on the same scene render
// alternately through one renderer; each keeps its own progressive accumulation (the count grows, neither frame size wipes
// the other), and an auxiliary screenshot (rendered into scratch, Renderer.cpp:1280) does not restart a camera's accumulation.
TEST_F(RendererFixture, b200_two_cameras_keep_their_own_accumulation) {
    using namespace Bifrost;
    using namespace Bifrost::Scene;

    auto size_a = Math::Vector2i(8, 6), size_b = Math::Vector2i(5, 4);
    CameraID camera_a = create_ortho_camera(size_a, optix::make_float3(0.25f, 0.5f, 0.75f));
    Math::Matrix4x4f orthographic_matrix, inverse_orthographic_matrix;
    CameraUtils::compute_orthographic_projection(float(size_b.x), float(size_b.y), 1000.0f, orthographic_matrix, inverse_orthographic_matrix);
    CameraID camera_b = Cameras::create("Second", Cameras::get_scene_ID(camera_a), orthographic_matrix, inverse_orthographic_matrix);
    Cameras::set_renderer_ID(camera_b, renderer->get_renderer_ID());
    renderer->handle_updates();
    auto target_a = create_render_target(renderer, size_a), target_b = create_render_target(renderer, size_b);
    for (unsigned int frame = 1; frame <= 3; ++frame) {
        EXPECT_EQ(frame, renderer->render(camera_a, target_a, size_a));
        EXPECT_EQ(frame, renderer->render(camera_b, target_b, size_b));
    }
    half4* pixels_a = (half4*)target_a->map();
    for (int i = 0; i < size_a.x * size_a.y; ++i) {
        EXPECT_FLOAT_EQ_EPS(0.25f, float(pixels_a[i].r), 1e-3f); EXPECT_FLOAT_EQ_EPS(0.75f, float(pixels_a[i].b), 1e-3f);
    }
    target_a->unmap();
    half4* pixels_b = (half4*)target_b->map();
    for (int i = 0; i < size_b.x * size_b.y; ++i) {
        EXPECT_FLOAT_EQ_EPS(0.25f, float(pixels_b[i].r), 1e-3f); EXPECT_FLOAT_EQ_EPS(0.5f, float(pixels_b[i].g), 1e-3f);
    }
    target_b->unmap();

    auto screenshots = renderer->request_auxiliary_buffers(camera_a, Screenshot::Content::Depth, size_a);
    EXPECT_EQ(1u, screenshots.size());
    EXPECT_EQ(4u, renderer->render(camera_a, target_a, size_a)); // the progressive render continues
    EXPECT_EQ(4u, renderer->render(camera_b, target_b, size_b));
}

// Checkpoint / resume: a render that is saved after three samples, carried on, rolled back to the saved state and carried on
// again arrives at the same pixels (a sample is a pure function of pixel, accumulation index and scene).
TEST_F(RendererFixture, b200_checkpoint_and_resume) {
    using namespace Bifrost;
    using namespace Bifrost::Scene;

    auto size = Math::Vector2i(6, 5);
    CameraID camera_ID = create_ortho_camera_with_quad_scene(size, optix::make_float3(0.25f, 0.5f, 0.75f));
    renderer->handle_updates();
    auto target = create_render_target(renderer, size);
    for (unsigned int frame = 1; frame <= 3; ++frame) EXPECT_EQ(frame, renderer->render(camera_ID, target, size));
    const std::filesystem::path file = std::filesystem::temp_directory_path() / "bpt_checkpoint_gtest.bin";
    ASSERT_TRUE(renderer->save_accumulation(camera_ID, file));
    EXPECT_EQ(4u, renderer->render(camera_ID, target, size));
    EXPECT_EQ(5u, renderer->render(camera_ID, target, size));
    std::vector<unsigned short> uninterrupted(4 * size.x * size.y);
    memcpy(uninterrupted.data(), target->map(), uninterrupted.size() * sizeof(unsigned short)); target->unmap();

    ASSERT_TRUE(renderer->load_accumulation(camera_ID, file)); // back to three samples
    EXPECT_EQ(4u, renderer->render(camera_ID, target, size));
    EXPECT_EQ(5u, renderer->render(camera_ID, target, size));
    EXPECT_EQ(0, memcmp(uninterrupted.data(), target->map(), uninterrupted.size() * sizeof(unsigned short))); target->unmap();
    EXPECT_FALSE(renderer->load_accumulation(camera_ID, file.string() + ".missing"));
    std::filesystem::remove(file);
}

} // namespace OptiXRenderer

// tests/OptiXRendererTests/Utils.cpp uses windows.h; the data directory is unused by this implementation.
std::filesystem::path get_data_directory() { return std::filesystem::path("."); }

int main(int argc, char** argv) {
    testing::InitGoogleTest(&argc, argv);
    return RUN_ALL_TESTS();
}
