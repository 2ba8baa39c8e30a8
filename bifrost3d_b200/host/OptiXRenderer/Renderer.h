// Drop-in replacement of the reference's renderer class (extensions/OptiXRenderer/OptiXRenderer/Renderer.h:40-86) on top of
// libbpt.so (include/bpt_c_api.h). Class name, method names, argument types and behaviour follow the reference so that code
// written against it (DX11OptiXAdaptor, tests/OptiXRendererTests/RendererTest.h) compiles and links unchanged; there is no
// OptiX underneath, `optix::Buffer` / `optix::Context` are the small facade types of host/optix_facade.
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#pragma once

#include <OptiXRenderer/PublicTypes.h>

#include <Bifrost/Scene/Camera.h>
#include <Bifrost/Utils/IdDeclarations.h>

#include <filesystem>
#include <vector>

namespace optix {
template <class T> class Handle;
class BufferObj;
class ContextObj;
using Buffer = Handle<BufferObj>;
using Context = Handle<ContextObj>;
} // namespace optix

namespace OptiXRenderer {

class Renderer final {
    using CameraID = Bifrost::Scene::CameraID;
    using SceneRootID = Bifrost::Scene::SceneRootID;
    using Vector2i = Bifrost::Math::Vector2i;

public:
    // ---- life time -------------------------------------------------------------------------------------------------
    // nullptr (with the reason printed) when no CUDA device is usable; never throws (Renderer.cpp:1365-1378).
    static Renderer* initialize(int cuda_device_ID, const std::filesystem::path& data_directory);
    ~Renderer();
    Renderer(Renderer&) = delete;
    Renderer& operator=(const Renderer&) = delete;

    Bifrost::Core::RendererID get_renderer_ID() const { return m_renderer_ID; }
    optix::Context& get_context();

    // ---- per-camera, per-scene and global settings (Renderer.cpp:1389-1474) ------------------------------------------
    Backend get_backend(CameraID camera) const;
    void set_backend(CameraID camera, Backend backend);

    unsigned int get_max_bounce_count(CameraID camera) const;
    void set_max_bounce_count(CameraID camera, unsigned int bounces);

    unsigned int get_max_accumulation_count(CameraID camera) const;
    void set_max_accumulation_count(CameraID camera, unsigned int accumulations);

    int get_next_event_sample_count(SceneRootID scene_root) const;
    void set_next_event_sample_count(SceneRootID scene_root, int samples);

    PathRegularizationSettings get_path_regularization_settings() const;
    void set_path_regularization_settings(PathRegularizationSettings settings);

    AIDenoiserFlags get_AI_denoiser_flags() const; // accepted and stored; the OptiX denoiser backend is out of scope
    void set_AI_denoiser_flags(AIDenoiserFlags flags);

    // ---- per frame ---------------------------------------------------------------------------------------------------
    // Mirrors what changed in the Bifrost core managers since the last call onto the device (incrementally).
    void handle_updates();

    // One more progressive sample of every pixel; the running mean goes to `target` as half4. Returns the accumulation count.
    unsigned int render(CameraID camera, optix::Buffer target, Vector2i frame_size);

    // Depth / albedo / tint / roughness screenshots of the current scene (Renderer.cpp:1267-1358).
    std::vector<Bifrost::Scene::Screenshot> request_auxiliary_buffers(CameraID camera, Bifrost::Scene::Cameras::ScreenshotContent content,
                                                                      Vector2i frame_size);

    // ---- addition: sample-sharded multi-GPU rendering ------------------------------------------------------------------
    // This renderer's accumulation indices start at `first_sample` (rank r of R renders [r * spp / R, (r + 1) * spp / R)).
    void set_sample_range(unsigned int first_sample) { m_first_sample = first_sample; }
    // The collective that combines the ranks' accumulation buffers (bpt_comm_* / bpt_reduce_accumulation of the C ABI, NCCL
    // underneath). Rank 0 creates the 128-byte id and hands it to the other ranks by whatever means the host has; every rank
    // joins; after the last sample every rank calls reduce_accumulation, and `root` then resolves the mean of all samples
    // with the next render-less resolve (resolve_accumulation).
    static bool create_communicator_id(char id[128]);
    bool join_communicator(const char id[128], int rank_count, int rank);
    bool reduce_accumulation(CameraID camera, int root);
    // Writes the current mean of `camera`'s accumulation to `target` without rendering another sample.
    bool resolve_accumulation(CameraID camera, optix::Buffer target);

    // ---- addition: checkpoint / resume ---------------------------------------------------------------------------------
    // The reference's resumable state - a camera's accumulation buffer and its `accumulations` count (Renderer.cpp:200-205,
    // 1262) - written to / restored from a file. A sample is a pure function of (pixel, accumulation index, scene), so the
    // render continues after load_accumulation exactly as if it had never stopped (same scene and camera assumed).
    bool save_accumulation(CameraID camera, const std::filesystem::path& file);
    bool load_accumulation(CameraID camera, const std::filesystem::path& file);

private:
    Renderer(int cuda_device_ID, const std::filesystem::path& data_directory);

    Bifrost::Core::RendererID m_renderer_ID;
    unsigned int m_first_sample = 0;
    struct Implementation; // everything CUDA / C ABI related lives behind this pointer
    Implementation* m_impl;
};

} // namespace OptiXRenderer
