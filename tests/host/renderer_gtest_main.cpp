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

// tests/OptiXRendererTests/Utils.cpp uses windows.h; the data directory is unused by this implementation.
std::filesystem::path get_data_directory() { return std::filesystem::path("."); }

int main(int argc, char** argv) {
    testing::InitGoogleTest(&argc, argv);
    return RUN_ALL_TESTS();
}
