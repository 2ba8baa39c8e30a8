// OptiXRenderer::Renderer re-created on top of libbpt.so (include/bpt_c_api.h): same class, same public methods and
// semantics as extensions/OptiXRenderer/OptiXRenderer/Renderer.h:40-86, so code written against the reference
// (DX11OptiXAdaptor, OptiXRendererTests/RendererTest.h) compiles and runs unchanged. No OptiX underneath.
#ifndef _OPTIXRENDERER_RENDERER_H_
#define _OPTIXRENDERER_RENDERER_H_

#include <OptiXRenderer/PublicTypes.h>

#include <Bifrost/Scene/Camera.h>
#include <Bifrost/Utils/IdDeclarations.h>

#include <filesystem>
#include <vector>

namespace optix {
template <class T> class Handle;
class BufferObj;
typedef Handle<BufferObj> Buffer;
class ContextObj;
typedef Handle<ContextObj> Context;
}

namespace OptiXRenderer {

class Renderer final {
public:
    // Returns nullptr (and prints why) when no CUDA device / context can be created; never throws. Renderer.cpp:1365-1378.
    static Renderer* initialize(int cuda_device_ID, const std::filesystem::path& data_directory);
    ~Renderer();

    Bifrost::Core::RendererID get_renderer_ID() const { return m_renderer_ID; }

    Backend get_backend(Bifrost::Scene::CameraID camera_ID) const;
    void set_backend(Bifrost::Scene::CameraID camera_ID, Backend backend);
    unsigned int get_max_bounce_count(Bifrost::Scene::CameraID camera_ID) const;
    void set_max_bounce_count(Bifrost::Scene::CameraID camera_ID, unsigned int bounce_count);
    unsigned int get_max_accumulation_count(Bifrost::Scene::CameraID camera_ID) const;
    void set_max_accumulation_count(Bifrost::Scene::CameraID camera_ID, unsigned int accumulation_count);
    int get_next_event_sample_count(Bifrost::Scene::SceneRootID scene_root_ID) const;
    void set_next_event_sample_count(Bifrost::Scene::SceneRootID scene_root_ID, int sample_count);
    PathRegularizationSettings get_path_regularization_settings() const;
    void set_path_regularization_settings(PathRegularizationSettings settings);
    AIDenoiserFlags get_AI_denoiser_flags() const;
    void set_AI_denoiser_flags(AIDenoiserFlags flags);

    // Reads the change lists of the Bifrost core managers and mirrors the scene on the device.
    void handle_updates();

    // Renders one more progressive sample into the camera's accumulation buffer and writes the running mean as half4
    // into `buffer`. Returns the accumulation count.
    unsigned int render(Bifrost::Scene::CameraID camera_ID, optix::Buffer buffer, Bifrost::Math::Vector2i frame_size);

    std::vector<Bifrost::Scene::Screenshot> request_auxiliary_buffers(Bifrost::Scene::CameraID camera_ID,
        Bifrost::Scene::Cameras::ScreenshotContent content_requested, Bifrost::Math::Vector2i frame_size);

    optix::Context& get_context();

    // Additions for sample-sharded multi-GPU rendering: render accumulation indices [first, first + count).
    void set_sample_range(unsigned int first_sample) { m_first_sample = first_sample; }

private:
    Renderer(int cuda_device_ID, const std::filesystem::path& data_directory);
    Renderer(Renderer& other) = delete;
    Renderer& operator=(const Renderer& rhs) = delete;

    Bifrost::Core::RendererID m_renderer_ID;
    unsigned int m_first_sample = 0;
    struct Implementation;
    Implementation* m_impl;
};

} // namespace OptiXRenderer

#endif
