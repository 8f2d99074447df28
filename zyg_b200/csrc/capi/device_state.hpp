// Shared state behind the opaque `zygpu_device` handle (include/zygpu.h) and the error helpers of the C ABI.
#pragma once

#include "../../../include/zygpu.h"

#include "../device/render.cuh"
#include "../device/trace.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

inline std::string& zygpuError() {
    thread_local std::string error;
    return error;
}

inline int fail(const char* fmt, ...) {
    char    buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    zygpuError() = buf;
    return -1;
}

#define CUDA_OK(expr)                                                                    \
    do {                                                                                 \
        const cudaError_t e_ = (expr);                                                   \
        if (cudaSuccess != e_) return fail("%s: %s", #expr, cudaGetErrorString(e_));     \
    } while (0)

struct DeviceMesh {
    zygpu::MeshDevice  view{};
    zygpu::MeshShading shading{};
    uint64_t           source     = 0;  // zyg_mesh::serial
    void*              buffers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

// Everything the render entry points keep on the device (capi/zygpu_render.cu).
struct RenderState {
    // Device buffers of the uploaded scene, reused by the next upload in call order when they are large enough: a frame of the
    // su_* API re-uploads the compiled scene, and a cudaFree / cudaMalloc cycle per array costs 0.1 - 1.5 s on some frames.
    std::vector<void*>  scene_buffers;
    std::vector<size_t> scene_buffer_bytes;
    size_t              scene_buffer_cursor = 0;
    zygpu::SceneDevice scene{};
    bool               has_scene  = false;
    bool               has_meshes = false;
    bool               has_image_area_lights = false;  // a finite PROP_IMAGE light: shadow records carry the sample's uv
    bool               has_textures = false;  // a material reads an image per vertex: shade_a hands its stochastic_r to shade_b
    bool               deferred_lights = false;  // light selection / sampling in the persistent light kernels
    bool               can_split  = false;  // a material can split paths: 4 vertex records per camera sample
    ZygpuView          view{};
    bool               has_view = false;

    zygpu::PathState   paths{};
    std::vector<void*> path_buffers;
    uint32_t           max_light_samples   = 1;  // shadow records reserved per path
    uint32_t           worst_light_samples = 1;  // what a path vertex can need at most (Tree.potentialMaxLights)

    float4*  film        = nullptr;
    float4*  resolved    = nullptr;
    uint32_t film_pixels = 0;
    float*   film_alpha  = nullptr;  // views with alpha_transparency: the alpha lane of the Transparent buffer (sum of weight * alpha)
    zygpu::AovFilm aov{};  // one layer per class of ZygpuView.aov_slots (aov.Buffer)

    cudaStream_t stream = nullptr;

    unsigned long long* tally    = nullptr;  // zygpu_set_counting: fetch / step counters of the instrumented traversal kernels
    bool                counting = false;

    ZygpuRenderStats stats{};
    uint64_t         stats_carry[2] = {0, 0};  // closest / shadow rays counted by path buffers that were replaced since the clear

    // progress: passes enqueued by zygpu_render / passes the device has finished (a host function on the stream counts them)
    uint32_t              passes_enqueued = 0;
    std::atomic<uint32_t> passes_completed{0};
};

struct zygpu_device {
    int                     ordinal = 0;
    std::vector<DeviceMesh> meshes;

    // staging for the host-buffer entry point
    static constexpr int      kStreams    = 3;
    static constexpr uint64_t kChunkRays  = 1u << 20;
    cudaStream_t              streams[kStreams] = {};
    void*                     d_rays[kStreams]  = {};
    void*                     d_out[kStreams]   = {};
    zygpu::TraceCounters*     d_counters        = nullptr;

    // work counters of the persistent kernels: one per in-flight launch, handed out round-robin
    static constexpr int kWorkCounters = 64;
    uint32_t*            d_work        = nullptr;
    int                  next_work     = 0;
    uint32_t*            workCounter() { return d_work + (next_work++ % kWorkCounters); }

    // traversal-stack scratch of the ray-pool kernels (device/trace.cu): [main region][one region per staging stream]
    void*  d_stacks           = nullptr;
    size_t stack_bytes_main   = 0;
    size_t stack_bytes_stream = 0;

    RenderState render;
};

void zygpuReleaseRender(zygpu_device* dev);

