// zyg's C API (src/capi/capi.zig) over the host scene model and the device ABI (include/zyg_su.h).
// One process-global engine like the reference (capi.zig:55). No compute happens here: render calls
// compile the scene on the host and launch the device passes through zygpu_*.
#include "../../../include/zyg_su.h"
#include "../../../include/zygpu.h"

#include "../host/image_writer.hpp"
#include "../host/mesh_handle.hpp"
#include "../host/scene_model.hpp"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <memory>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Engine {
    zyg::SceneModel scene;

    std::vector<zyg_mesh*> meshes;  // owned

    zygpu_device* device         = nullptr;
    int           device_ordinal = 0;

    uint32_t frame     = 0;
    uint32_t iteration = 0;

    int mesh_builder = 0;  // zyg_su_set_mesh_builder: 0 host (reference-order SAH tree), 1 device (LBVH)

    // Scene.compile + upload are skipped while nothing was edited since the last frame (driver.zig:154-180 recompiles every
    // frame; the products are the same, so the device keeps them)
    uint64_t uploaded_revision = ~0ull;

    // su_triangle_mesh_create(async = true): the BVH is built on a worker thread and joined by the next call that needs a
    // compiled scene or starts another build (one outstanding build, pool.zig:155-162; commitAsync, capi.zig:550,583)
    std::thread async_build;
    uint32_t    async_shape  = 0;
    zyg_mesh*   async_mesh   = nullptr;
    int         async_result = 0;
    std::string async_error;

    std::vector<float> target;  // Driver.target: resolved RGBA of the last su_resolve_frame

    // take.exporters (take.zig:303-331): the `Image` sinks of su_exporters_create
    struct Exporter {
        enum Format { PNG, EXR, RGBE } format = PNG;
        bool half            = true;   // EXR "bitdepth": 16
        bool error_diffusion = false;  // PNG
    };
    std::vector<Exporter> exporters;

    void (*log_post)(uint32_t, const char*) = nullptr;
    void (*progress_start)(uint32_t)        = nullptr;
    void (*progress_tick)()                 = nullptr;

    ~Engine() {
        if (async_build.joinable()) async_build.join();
        if (async_mesh) zyg_mesh_free(async_mesh);
        if (device) {
            zygpu_set_light_tree_builder(nullptr, 0);  // the builder hook must not outlive the device it builds on
            zygpu_destroy(device);
        }
        for (zyg_mesh* m : meshes) zyg_mesh_free(m);
    }
};

std::unique_ptr<Engine> g_engine;

enum LogLevel : uint32_t { Info = 0, Warning = 1, Error = 2 };  // log.zig:5-7

constexpr uint32_t kNumAovClasses = 9;  // AovValue.NumClasses, rendering/sensor/aov/aov_value.zig:10-44; capi.zig:615, 633

void logf(uint32_t level, const char* fmt, ...) {
    char    buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (g_engine && g_engine->log_post) {
        g_engine->log_post(level, buf);
    } else if (level >= Warning) {
        std::fprintf(stderr, "%s: %s\n", Warning == level ? "Warning" : "Error", buf);
    }
}

bool parse(const char* text, zyg::json::Value& out) {
    if (!text) return false;
    zyg::json::Parser p(text);
    return p.parse(out);
}

// commitAsync, capi.zig:550,583 / shape_provider.zig:127-144: the finished tree of an asynchronous build becomes the shape
int commitAsync(Engine& e) {
    if (!e.async_build.joinable()) return 0;
    e.async_build.join();
    if (0 != e.async_result || !e.async_mesh) {
        logf(Error, "%s", e.async_error.c_str());  // the reference swallows the failure (shape_provider.zig:326-328); the shape stays empty
        return -1;
    }
    e.meshes[e.async_shape - 7] = e.async_mesh;
    e.scene.setMesh(e.async_shape, e.async_mesh);
    e.async_mesh = nullptr;
    return 0;
}

int compileScene(Engine& e) {
    if (0 != commitAsync(e)) return -1;
    std::string error;
    const bool  ok = e.scene.compile(error);
    for (const std::string& w : e.scene.warnings()) logf(Warning, "%s", w.c_str());
    e.scene.warnings().clear();
    if (!ok) logf(Error, "%s", error.c_str());
    return ok ? 0 : -1;
}

// Scene.compile + upload + view, the host half of Driver.startFrame (driver.zig:154-180). The reference recompiles every frame;
// the products only change when the scene was edited, so an unchanged scene keeps what the device already holds.
int prepareFrame(Engine& e) {
    if (0 != commitAsync(e)) return -1;
    if (e.device && e.uploaded_revision == e.scene.revision()) return 0;
    if (0 != compileScene(e)) return -1;
    if (!e.device) {
        if (0 != zygpu_create(e.device_ordinal, &e.device)) {
            e.device = nullptr;
            logf(Error, "%s", zygpu_last_error());
            return -1;
        }
    }
    if (0 != zygpu_upload_scene(e.device, &e.scene.scene()) || 0 != zygpu_set_view(e.device, &e.scene.view())) {
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    e.uploaded_revision = e.scene.revision();
    return 0;
}

// Progressor.tick (progress.zig:3-25): the reference ticks once per finished tile (driver.zig:305). The device path has no
// tiles; it finishes passes (all pixels x some samples), so the `tiles` ticks are spread over the passes as they complete.
void tickWhileRendering(Engine& e, uint32_t tiles, uint32_t first_pass) {
    const uint32_t total = zygpu_passes_enqueued(e.device) - first_pass;
    uint32_t       ticked = 0;
    for (;;) {
        const uint32_t finished = zygpu_passes_completed(e.device) - first_pass;
        const uint32_t due      = 0 == total ? tiles : uint32_t(uint64_t(tiles) * std::min(finished, total) / total);
        for (; ticked < due; ++ticked) e.progress_tick();
        if (finished >= total) break;
        std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
}

int renderRange(Engine& e, uint32_t frame, uint32_t iteration, uint32_t num_samples) {
    if (0 != prepareFrame(e)) return -1;
    e.frame = frame;

    const uint32_t spp = num_samples > 0 ? num_samples : e.scene.samplesPerPixel();  // driver.zig:141-142

    // renderFrameForward, driver.zig:309-336: clear, then every tile; the progressor ticks once per tile
    const uint32_t tiles = ((e.scene.width() + 31) / 32) * ((e.scene.height() + 31) / 32);
    if (e.progress_start) e.progress_start(tiles);

    const uint32_t first_pass = zygpu_passes_enqueued(e.device);
    if (0 != zygpu_clear_film(e.device) || 0 != zygpu_render(e.device, iteration, spp)) {
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    if (e.progress_tick) tickWhileRendering(e, tiles, first_pass);
    if (0 != zygpu_synchronize(e.device)) {
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    return 0;
}

}  // namespace

extern "C" {

int32_t su_init(void) {
    if (g_engine) return -1;
    g_engine.reset(new Engine);
    g_engine->scene.setSamplesPerPixel(1);  // capi.zig:104
    return 0;
}

int32_t su_release(void) {
    if (!g_engine) return -1;
    g_engine.reset();
    return 0;
}

int32_t su_mount(const char*) { return g_engine ? 0 : -1; }

int32_t su_perspective_camera_create(uint32_t width, uint32_t height) {
    if (!g_engine) return -1;
    g_engine->scene.setCamera(width, height);
    return int32_t(g_engine->scene.cameraEntity());
}

int32_t su_camera_set_fov(float fov) {
    if (!g_engine) return -1;
    g_engine->scene.setFov(fov);
    return 0;
}

int32_t su_camera_sensor_dimensions(int32_t* dimensions) {
    if (!g_engine || !dimensions) return -1;
    dimensions[0] = int32_t(g_engine->scene.width());
    dimensions[1] = int32_t(g_engine->scene.height());
    return 0;
}

// Take.loadExporters, take.zig:303-331 (capi.zig:189-200). `Video` (an ffmpeg pipe) is not an image codec: warned about and skipped.
int32_t su_exporters_create(const char* json) {
    if (!g_engine) return -1;
    zyg::json::Value v;
    if (!parse(json, v) || zyg::json::Value::Object != v.kind) return -1;
    Engine& e = *g_engine;
    e.exporters.clear();
    for (const auto& entry : v.object) {
        if ("Image" == entry.first) {
            Engine::Exporter x;
            const zyg::json::Value* fm = entry.second.get("format");
            const std::string format = fm && zyg::json::Value::String == fm->kind ? fm->string : "PNG";
            if ("EXR" == format) {
                x.format = Engine::Exporter::EXR;
                x.half   = 16 == zyg::json::readUIntMember(entry.second, "bitdepth", 16);
            } else if ("RGBE" == format) {
                x.format = Engine::Exporter::RGBE;
            } else {
                x.format          = Engine::Exporter::PNG;
                x.error_diffusion = zyg::json::readBoolMember(entry.second, "error_diffusion", false);
            }
            e.exporters.push_back(x);
        } else if ("Video" == entry.first) {
            logf(Warning, "Video exporter (ffmpeg pipe) is not supported: skipped");
        }
    }
    return 0;
}

// capi.zig:202-213 -> View.loadAOV (take.zig:106-129)
int32_t su_aovs_create(const char* json) {
    if (!g_engine) return -1;
    zyg::json::Value v;
    if (!parse(json, v)) return -1;
    g_engine->scene.loadAovs(v);
    return 0;
}

int32_t su_sampler_create(uint32_t num_samples) {
    if (g_engine) g_engine->scene.setSamplesPerPixel(num_samples);
    return -1;  // capi.zig:215-221 returns -1 on every path
}

int32_t su_integrators_create(const char* json) {
    if (!g_engine) return -1;
    zyg::json::Value v;
    if (!parse(json, v)) return -1;
    g_engine->scene.loadIntegrators(v);
    return 0;
}

int32_t su_image_create(uint32_t id, uint32_t format, uint32_t num_channels, uint32_t width, uint32_t height, uint32_t depth,
                        uint32_t pixel_stride, const uint8_t* data) {
    if (!g_engine) return -1;
    return g_engine->scene.createImage(id, format, num_channels, width, height, depth, pixel_stride, data);
}

int32_t su_image_update(uint32_t id, uint32_t pixel_stride, const uint8_t* data) {
    if (!g_engine) return -1;
    return g_engine->scene.updateImage(id, pixel_stride, data);
}

int32_t su_material_create(uint32_t /*id*/, const char* json) {
    if (!g_engine) return -1;
    zyg::json::Value v;
    if (!parse(json, v)) return -1;
    return g_engine->scene.createMaterial(v);
}

int32_t su_material_update(uint32_t id, const char* json) {
    if (!g_engine) return -1;
    zyg::json::Value v;
    if (!parse(json, v)) return -1;
    if (id >= g_engine->scene.numMaterials()) return -3;
    return g_engine->scene.updateMaterial(id, v) ? 0 : -4;
}

int32_t su_triangle_mesh_create(uint32_t /*id*/, uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles,
                                const uint32_t* indices, uint32_t num_vertices, const float* positions,
                                uint32_t positions_stride, const float* normals, uint32_t normals_stride,
                                const float* /*tangents*/, uint32_t /*tangents_stride*/, const float* uvs,
                                uint32_t uvs_stride, bool async) {
    if (!g_engine) return -1;
    if (num_triangles < 1 || num_vertices < 1 || !positions) return -1;  // capi.zig:398-408

    Engine& e = *g_engine;
    commitAsync(e);  // one outstanding build (pool.zig:155-162): a second request waits for the first
    if (1 == e.mesh_builder && num_triangles >= 4) {
        // the device build takes milliseconds: nothing to gain from the async thread
        if (!e.device && 0 != zygpu_create(e.device_ordinal, &e.device)) {
            e.device = nullptr;
            logf(Error, "%s", zygpu_last_error());
            return -1;
        }
        zyg_mesh* mesh = nullptr;
        if (0 != zygpu_mesh_build(e.device, num_parts, parts, num_triangles, indices, num_vertices, positions, positions_stride, normals,
                                  normals_stride, uvs, uvs_stride, &mesh, nullptr)) {
            logf(Error, "%s", zygpu_last_error());
            return -1;
        }
        e.meshes.push_back(mesh);
        return int32_t(e.scene.addMesh(mesh, num_parts > 0 ? num_parts : 1));
    }
    if (async) {
        // shape_provider.zig:299-303, 847-913: the build reads the caller's buffers on the async thread; they must outlive the
        // next call that commits (su_render_frame / su_start_frame / another mesh). The shape id is handed out at once.
        e.meshes.push_back(nullptr);
        e.async_shape  = e.scene.addMesh(nullptr, num_parts > 0 ? num_parts : 1);
        e.async_result = 0;
        e.async_build  = std::thread([=, &e] {
            e.async_result = zyg_mesh_build(num_parts, parts, num_triangles, indices, num_vertices, positions, positions_stride, normals,
                                            normals_stride, uvs, uvs_stride, 0, &e.async_mesh);
            if (0 != e.async_result) e.async_error = zygpu_last_error();  // the error slot is per thread
        });
        return int32_t(e.async_shape);
    }

    zyg_mesh* mesh = nullptr;
    if (0 != zyg_mesh_build(num_parts, parts, num_triangles, indices, num_vertices, positions, positions_stride, normals,
                            normals_stride, uvs, uvs_stride, 0, &mesh)) {
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    g_engine->meshes.push_back(mesh);
    return int32_t(g_engine->scene.addMesh(mesh, num_parts > 0 ? num_parts : 1));
}

static int32_t propCreate(uint32_t shape, uint32_t num_materials, const uint32_t* materials, bool unoccluding) {
    if (!g_engine) return -1;
    zyg::SceneModel& scene = g_engine->scene;
    if (shape >= scene.numShapes()) return -1;

    // capi.zig:431-449: out-of-range ids and missing entries fall back to the debug material
    std::vector<uint32_t> mats;
    for (uint32_t i = 0; i < num_materials; ++i) {
        mats.push_back(materials[i] >= scene.numMaterials() ? scene.fallbackMaterial() : materials[i]);
    }
    if (mats.empty()) mats.push_back(scene.fallbackMaterial());
    return int32_t(scene.createPropShape(shape, mats.data(), uint32_t(mats.size()), unoccluding));
}

int32_t su_prop_create(uint32_t shape, uint32_t num_materials, const uint32_t* materials) {
    return propCreate(shape, num_materials, materials, false);
}

int32_t zyg_su_prop_create_unoccluding(uint32_t shape, uint32_t num_materials, const uint32_t* materials) {
    return propCreate(shape, num_materials, materials, true);
}

int32_t zyg_su_instancer_create(uint32_t num_prototypes, const uint32_t* prototypes, uint32_t num_instances,
                                const uint32_t* prototype_indices, const float* transformations) {
    if (!g_engine || (num_instances > 0 && !transformations)) return -1;
    std::vector<zyg::Transformation> trafos(num_instances);
    for (uint32_t i = 0; i < num_instances; ++i) zyg::decomposeMatrix(transformations + size_t(i) * 16, trafos[i]);
    return g_engine->scene.createInstancer(prototypes, num_prototypes, prototype_indices, trafos.data(), num_instances);
}

int32_t su_prop_create_instance(uint32_t entity) {
    if (!g_engine) return -1;
    return g_engine->scene.createPropInstance(entity);
}

int32_t su_light_create(uint32_t prop) {
    if (!g_engine) return -1;
    return g_engine->scene.createLight(prop) ? 0 : -1;
}

int32_t su_prop_set_transformation(uint32_t prop, const float* trafo) {
    if (!g_engine || !trafo) return -1;
    zyg::Transformation t;
    zyg::decomposeMatrix(trafo, t);
    return g_engine->scene.setWorldTransformation(prop, t) ? 0 : -1;
}

int32_t su_prop_set_transformation_frame(uint32_t prop, uint32_t frame, const float* trafo) {
    if (0 != frame) return -1;  // static scenes only
    return su_prop_set_transformation(prop, trafo);
}

int32_t su_prop_set_visibility(uint32_t prop, uint32_t in_camera, uint32_t in_reflection, uint32_t in_sss) {
    if (!g_engine) return -1;
    return g_engine->scene.setVisibility(prop, in_camera > 0, in_reflection > 0, in_sss > 0) ? 0 : -1;
}

int32_t su_render_frame(uint32_t frame) {
    if (!g_engine) return -1;
    return renderRange(*g_engine, frame, 0, 0);  // driver.render(0, frame, 0, 0), capi.zig:559
}

int32_t zyg_su_render_frame_range(uint32_t frame, uint32_t iteration, uint32_t num_samples) {
    if (!g_engine) return -1;
    return renderRange(*g_engine, frame, iteration, num_samples);
}

// Driver.exportFrame + ImageSequence.write, driver.zig:224-253, exporting/image_sequence.zig:24-56: resolve the beauty, then one
// file "image_<camera:02>_<frame:06>.<ext>" per exporter in the working directory; with the Transparent sensor ("alpha_transparency")
// PNG and EXR carry the alpha channel. Every recorded AOV class follows as image_<..>_<frame><_albedo|_depth|_mat|_ng|_n|_r|_emission|
// _direct|_indirect>.<ext> in the encoding of its class (aov_value.zig:32-40).
int32_t su_export_frame(void) {
    if (!g_engine || !g_engine->device) return -1;
    Engine&        e = *g_engine;
    const uint32_t n = e.scene.width() * e.scene.height();
    e.target.resize(size_t(n) * 4);
    if (0 != zygpu_resolve(e.device, e.target.data(), n)) {
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    const int32_t* crop = e.scene.view().crop;
    const bool     alpha = e.scene.alphaTransparency();  // ImageSequence.alpha = sensor.class.alphaTransparency(), take.zig:311
    // the beauty, then every recorded AOV class under its own name and encoding (driver.zig:231-250, image_sequence.zig:36-78)
    static const char* const kAovExtension[kNumAovClasses] = {"_albedo", "_depth", "_mat", "_ng", "_n", "_r", "_emission", "_direct", "_indirect"};
    static const zyg::Encoding kAovEncoding[kNumAovClasses] = {zyg::Encoding::Color, zyg::Encoding::Depth,  zyg::Encoding::Id,
                                                               zyg::Encoding::Normal, zyg::Encoding::Normal, zyg::Encoding::Float,
                                                               zyg::Encoding::Color, zyg::Encoding::Color,  zyg::Encoding::Color};
    for (uint32_t aov = 0; aov < kNumAovClasses; ++aov) {
        if (0 == (e.scene.aovSlots() & (1u << aov))) continue;
        std::vector<float> layer(size_t(n) * 4);
        if (0 != zygpu_resolve_aov(e.device, aov, layer.data(), n, 0)) {
            logf(Error, "%s", zygpu_last_error());
            return -1;
        }
        for (const Engine::Exporter& x : e.exporters) {
            std::vector<uint8_t> bytes;
            const char*          ext = "png";
            bool                 ok  = false;
            switch (x.format) {
                case Engine::Exporter::PNG:
                    ok = zyg::encodePngAs(bytes, layer.data(), int32_t(e.scene.width()), int32_t(e.scene.height()), crop, kAovEncoding[aov], x.error_diffusion);
                    break;
                case Engine::Exporter::EXR:
                    ext = "exr";
                    ok  = zyg::encodeExrAs(bytes, layer.data(), int32_t(e.scene.width()), int32_t(e.scene.height()), crop, kAovEncoding[aov], x.half);
                    break;
                case Engine::Exporter::RGBE:  // the RGBE writer has no notion of an encoding (rgbe_writer.zig): the resolved values as they are
                    ext = "hdr";
                    ok  = zyg::encodeRgbe(bytes, layer.data(), int32_t(e.scene.width()), int32_t(e.scene.height()), crop);
                    break;
            }
            char name[64];
            std::snprintf(name, sizeof(name), "image_%02u_%06u%s.%s", 0u, e.frame, kAovExtension[aov], ext);
            if (!ok || !zyg::writeFile(name, bytes)) {
                logf(Error, "Exporting frame %u to %s failed", e.frame, name);
                return -1;
            }
        }
    }
    for (const Engine::Exporter& x : e.exporters) {
        std::vector<uint8_t> bytes;
        const char*          ext = "png";
        bool                 ok  = false;
        switch (x.format) {
            case Engine::Exporter::PNG:
                ok = zyg::encodePng(bytes, e.target.data(), int32_t(e.scene.width()), int32_t(e.scene.height()), crop, alpha, x.error_diffusion);
                break;
            case Engine::Exporter::EXR:
                ext = "exr";
                ok  = zyg::encodeExr(bytes, e.target.data(), int32_t(e.scene.width()), int32_t(e.scene.height()), crop, alpha, x.half);
                break;
            case Engine::Exporter::RGBE:
                ext = "hdr";
                ok  = zyg::encodeRgbe(bytes, e.target.data(), int32_t(e.scene.width()), int32_t(e.scene.height()), crop);
                break;
        }
        char name[64];
        std::snprintf(name, sizeof(name), "image_%02u_%06u.%s", 0u, e.frame, ext);
        if (!ok || !zyg::writeFile(name, bytes)) {
            logf(Error, "Exporting frame %u to %s failed", e.frame, name);
            return -1;
        }
    }
    return 0;
}

int32_t su_start_frame(uint32_t frame) {
    if (!g_engine) return -1;
    Engine& e = *g_engine;
    if (0 != prepareFrame(e)) return -1;
    e.frame     = frame;
    e.iteration = 0;
    return 0 == zygpu_clear_film(e.device) ? 0 : -1;  // startFrame(progressive = true), driver.zig:175-179
}

int32_t su_render_iterations(uint32_t num_steps) {
    if (!g_engine || !g_engine->device) return -1;
    Engine& e = *g_engine;
    if (0 != zygpu_render(e.device, e.iteration, num_steps)) {
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    e.iteration += num_steps;
    return 0;
}

int32_t su_resolve_frame(uint32_t aov) {
    if (!g_engine || !g_engine->device) return -1;
    Engine& e = *g_engine;
    const uint32_t n = e.scene.width() * e.scene.height();
    e.target.resize(size_t(n) * 4);
    if (aov < kNumAovClasses) {  // Driver.resolveAov, driver.zig:219-222: -2 when the class is not recorded
        const int rc = zygpu_resolve_aov(e.device, aov, e.target.data(), n, 0);
        return 0 == rc ? 0 : (-2 == rc ? -2 : -1);
    }
    return 0 == zygpu_resolve(e.device, e.target.data(), n) ? 0 : -1;  // anything above the classes resolves the beauty
}

int32_t su_resolve_frame_to_buffer(uint32_t aov, uint32_t width, uint32_t height, float* buffer) {
    if (!g_engine || !g_engine->device || !buffer) return -1;
    Engine& e = *g_engine;
    const uint32_t n = std::min(width * height, e.scene.width() * e.scene.height());  // capi.zig:628
    if (aov < kNumAovClasses) {
        const int rc = zygpu_resolve_aov(e.device, aov, buffer, n, 0);
        return 0 == rc ? 0 : (-2 == rc ? -2 : -1);
    }
    return 0 == zygpu_resolve(e.device, buffer, n) ? 0 : -1;
}

// CopyFramebufferContext, capi.zig:662-724: the resolved target as UInt8 sRGB-gamma or Float32, 3 or 4 channels
int32_t su_copy_framebuffer(uint32_t format, uint32_t num_channels, uint32_t width, uint32_t height, uint8_t* destination) {
    if (!g_engine || !destination) return -1;
    Engine&        e = *g_engine;
    const uint32_t w = std::min(width, e.scene.width());
    const uint32_t h = std::min(height, e.scene.height());
    if (e.target.size() < size_t(e.scene.width()) * e.scene.height() * 4) return -1;

    auto linearToGamma = [](float c) -> float {  // spectrum/srgb.zig linearToGamma
        if (c <= 0.f) return 0.f;
        if (c < 0.0031308f) return 12.92f * c;
        if (c < 1.f) return 1.055f * std::pow(c, 1.f / 2.4f) - 0.055f;
        return 1.f;
    };

    for (uint32_t y = 0; y < h; ++y) {
        for (uint32_t x = 0; x < w; ++x) {
            const float* src = e.target.data() + (size_t(y) * e.scene.width() + x) * 4;
            const size_t o   = (size_t(y) * width + x) * num_channels;
            if (0 == format) {  // UInt8
                for (uint32_t c = 0; c < num_channels; ++c) {
                    const float v      = c < 3 ? linearToGamma(src[c]) : src[3];
                    destination[o + c] = uint8_t(v * 255.f + 0.5f);
                }
            } else if (4 == format) {  // Float32
                float* dst = reinterpret_cast<float*>(destination) + o;
                for (uint32_t c = 0; c < num_channels; ++c) dst[c] = src[c];
            } else {
                return -1;
            }
        }
    }
    return 0;
}

int32_t su_register_log(void (*post)(uint32_t, const char*)) {
    if (!g_engine) return -1;
    g_engine->log_post = post;
    return 0;
}

int32_t su_register_progress(void (*start)(uint32_t), void (*tick)(void)) {
    if (!g_engine) return -1;
    g_engine->progress_start = start;
    g_engine->progress_tick  = tick;
    return 0;
}

int32_t zyg_su_sensor_create(const char* json) {
    if (!g_engine) return -1;
    zyg::json::Value v;
    if (!parse(json, v)) return -1;
    g_engine->scene.loadSensor(v);
    return 0;
}

int32_t zyg_su_camera_set_lens(float aperture_radius, float focus_distance) {
    if (!g_engine) return -1;
    g_engine->scene.setLens(aperture_radius, focus_distance);
    return 0;
}

int32_t zyg_su_set_device(int32_t ordinal) {
    if (!g_engine || g_engine->device) return -1;
    g_engine->device_ordinal = ordinal;
    return 0;
}

int32_t zyg_su_set_mesh_builder(int32_t builder) {
    if (!g_engine || builder < 0 || builder > 1) return -1;
    g_engine->mesh_builder = builder;
    return 0;
}

int32_t zyg_su_set_light_tree_builder(int32_t builder, uint32_t min_lights) {
    if (!g_engine || builder < 0 || builder > 1) return -1;
    Engine& e = *g_engine;
    if (0 == builder) return zygpu_set_light_tree_builder(nullptr, 0);
    if (!e.device && 0 != zygpu_create(e.device_ordinal, &e.device)) {
        e.device = nullptr;
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    return zygpu_set_light_tree_builder(e.device, min_lights);
}

int32_t zyg_su_triangle_mesh_refit(uint32_t shape, const float* positions, uint32_t positions_stride, const float* normals,
                                   uint32_t normals_stride) {
    if (!g_engine || shape < 7 || shape - 7 >= g_engine->meshes.size() || !positions) return -1;
    Engine& e = *g_engine;
    if (0 != commitAsync(e) || !e.meshes[shape - 7]) return -1;
    if (!e.device && 0 != zygpu_create(e.device_ordinal, &e.device)) {
        e.device = nullptr;
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    if (0 != zygpu_mesh_refit(e.device, e.meshes[shape - 7], positions, positions_stride, normals, normals_stride, nullptr)) {
        logf(Error, "%s", zygpu_last_error());
        return -1;
    }
    e.scene.setMesh(shape, e.meshes[shape - 7]);  // the bounds and areas changed: the next frame compiles and uploads again
    return 0;
}

int32_t zyg_su_camera_set_crop(int32_t x0, int32_t y0, int32_t x1, int32_t y1) {
    if (!g_engine) return -1;
    g_engine->scene.setCrop(x0, y0, x1, y1);
    return 0;
}

int32_t zyg_su_compile(const ZygpuScene** scene, const ZygpuView** view) {
    if (!g_engine) return -1;
    if (0 != compileScene(*g_engine)) return -1;
    g_engine->uploaded_revision = ~0ull;  // compile() rebuilt the arrays the last upload was made from
    if (scene) *scene = &g_engine->scene.scene();
    if (view) *view = &g_engine->scene.view();
    return 0;
}

void* zyg_su_device(void) { return g_engine ? g_engine->device : nullptr; }

const zyg_mesh* zyg_su_mesh(uint32_t shape) {
    if (!g_engine || shape < 7 || shape - 7 >= g_engine->meshes.size()) return nullptr;
    return g_engine->meshes[shape - 7];
}

int32_t zyg_su_denoise_frame_to_buffer(float sigma, uint32_t width, uint32_t height, float* buffer) {
    if (!g_engine || !g_engine->device || !buffer) return -1;
    Engine&        e  = *g_engine;
    const uint32_t n  = std::min(width * height, e.scene.width() * e.scene.height());
    const int      rc = zygpu_denoise(e.device, sigma, buffer, n);
    if (-1 == rc) logf(Error, "%s", zygpu_last_error());
    return 0 == rc ? 0 : (-2 == rc ? -2 : -1);
}

int32_t zyg_su_write_image(const char* path, uint32_t format, uint32_t flags, const float* rgba, int32_t width, int32_t height,
                           const int32_t* crop) {
    if (!path || !rgba || width < 1 || height < 1 || format > 2) return -1;
    const int32_t full[4] = {0, 0, width, height};
    const int32_t* c      = crop ? crop : full;
    std::vector<uint8_t> bytes;
    bool                 ok = false;
    const uint32_t      e        = (flags >> 8) & 7u;  // Writer.Encoding when not 0: Depth 2, Id 3, Normal 4, Float 5
    const zyg::Encoding encoding = e > 1u && e <= 5u ? zyg::Encoding(e) : (0 != (flags & 1u) ? zyg::Encoding::ColorAlpha : zyg::Encoding::Color);
    if (0 == format) ok = zyg::encodePngAs(bytes, rgba, width, height, c, encoding, 0 != (flags & 4u));
    if (1 == format) ok = zyg::encodeExrAs(bytes, rgba, width, height, c, encoding, 0 != (flags & 2u));
    if (2 == format) ok = zyg::encodeRgbe(bytes, rgba, width, height, c);
    return ok && zyg::writeFile(path, bytes) ? 0 : -1;
}

}  // extern "C"
