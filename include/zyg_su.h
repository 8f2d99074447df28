/* zyg_su.h — zyg's own C API (libzyg's `su_*` functions, src/capi/capi.zig:57-738) served by the B200 backend.
 *
 * A client of libzyg.so (src/capi-test/test.py, src/blender-plugin/engine.py) can load libzyg_b200.so
 * instead: same symbols, same argument meaning, same return convention (0 or a non-negative id on
 * success, -1 failure / not initialised, -3 bad material id, -4 material update failed). One
 * process-global engine, single caller thread (capi.zig:55).
 *
 * Scope: static scenes, perspective camera (pinhole / thin lens), PTMIS, the built-in shapes Rectangle / Cube / Disk /
 * Sphere / Canopy / Distant and triangle meshes (props, prop instances, instancers), materials Substitute / Glass / Light with
 * uniform parameters, image maps for colour / roughness / metallic / normal and emission (su_image_create: Float32 or UInt8 x 1, 2, 3),
 * the nine AOV classes next to the beauty. A scene that
 * uses a Dome prop, a Disk light with an emission map, thin or dispersive Glass is refused by the render calls (-1 and a log message) rather than rendered
 * wrongly; unsupported material parameters are ignored with a warning through the log callback.
 * Entry points outside that scope exist and return -1 (animation frames other than 0).
 */
#ifndef ZYG_SU_H
#define ZYG_SU_H

#include <stdbool.h>
#include <stdint.h>

#include "zygpu_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

int32_t su_init(void);                                                    /* capi.zig:57  */
int32_t su_release(void);                                                 /* :116 */
int32_t su_mount(const char* folder);                                     /* :131 (accepted, unused) */
int32_t su_perspective_camera_create(uint32_t width, uint32_t height);    /* :143 -> camera entity id */
int32_t su_camera_set_fov(float fov);                                     /* :169 (radians) */
int32_t su_camera_sensor_dimensions(int32_t* dimensions);                 /* :178 */
int32_t su_exporters_create(const char* json);                            /* :189 {"Image":{"format":"PNG"|"EXR"|"RGBE","bitdepth":16|32,
                                                                             "error_diffusion":bool}} (take.zig:303-331; "Video" skipped) */
int32_t su_aovs_create(const char* json);                                 /* :202 {"Albedo":bool,"Depth":..,"MaterialId":..,"GeometricNormal":..,
                                                                             "ShadingNormal":..,"Roughness":..,"Emission":..,"Direct":..,
                                                                             "Indirect":..} sets / clears the class bits (take.zig:106-129) */
int32_t su_sampler_create(uint32_t num_samples);                          /* :215 (returns -1 even on success, like the reference) */
int32_t su_integrators_create(const char* json);                          /* :223 */
int32_t su_image_create(uint32_t id, uint32_t format, uint32_t num_channels, uint32_t width, uint32_t height,
                        uint32_t depth, uint32_t pixel_stride, const uint8_t* data); /* :236 -> image id (the pixels are copied) */
int32_t su_image_update(uint32_t id, uint32_t pixel_stride, const uint8_t* data);   /* :300 */
int32_t su_material_create(uint32_t id, const char* json);                /* :342 -> material id */
int32_t su_material_update(uint32_t id, const char* json);                /* :355 */
int32_t su_triangle_mesh_create(uint32_t id, uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles,
                                const uint32_t* indices, uint32_t num_vertices, const float* positions,
                                uint32_t positions_stride, const float* normals, uint32_t normals_stride,
                                const float* tangents, uint32_t tangents_stride, const float* uvs, uint32_t uvs_stride,
                                bool async);                              /* :379 -> shape id. async: the BVH is built on a worker
                                                                             thread that reads the caller's buffers until the next
                                                                             su_render_frame / su_start_frame / mesh call joins it
                                                                             (shape_provider.zig:299-303; commitAsync capi.zig:550) */
int32_t su_prop_create(uint32_t shape, uint32_t num_materials, const uint32_t* materials); /* :425 -> prop id */
int32_t su_prop_create_instance(uint32_t entity);                         /* :457 -> prop id sharing the shape and materials */
int32_t su_light_create(uint32_t prop);                                   /* :471 */
int32_t su_prop_set_transformation(uint32_t prop, const float* trafo);    /* :485 (row-major 4x4, rows = basis vectors, last row = position) */
int32_t su_prop_set_transformation_frame(uint32_t prop, uint32_t frame, const float* trafo); /* :506 (frame 0 only) */
int32_t su_prop_set_visibility(uint32_t prop, uint32_t in_camera, uint32_t in_reflection, uint32_t in_sss); /* :535 */
int32_t su_render_frame(uint32_t frame);                                  /* :548 */
int32_t su_export_frame(void);                                            /* :569 resolves the beauty and writes image_00_<frame:06>.<ext>
                                                                             per exporter into the working directory
                                                                             (exporting/image_sequence.zig:24-56) */
int32_t su_start_frame(uint32_t frame);                                   /* :581 */
int32_t su_render_iterations(uint32_t num_steps);                         /* :602 */
int32_t su_resolve_frame(uint32_t aov);                                   /* :613 aov < 9 = AovValue.NumClasses: that class
                                                                             (aov_buffer.zig:51-82), -2 when it is not recorded;
                                                                             anything else resolves the beauty */
int32_t su_resolve_frame_to_buffer(uint32_t aov, uint32_t width, uint32_t height, float* buffer); /* :626 */
int32_t su_copy_framebuffer(uint32_t format, uint32_t num_channels, uint32_t width, uint32_t height,
                            uint8_t* destination);                        /* :643 */
int32_t su_register_log(void (*post)(uint32_t level, const char* text));  /* :726 */
int32_t su_register_progress(void (*start)(uint32_t resolution), void (*tick)(void)); /* :731 (start(number of 32 x 32 tiles), then
                                                                             that many ticks spread over the passes as they finish) */

/* ---- extensions (not in libzyg): what a take / scene file sets and the C API cannot ---------------- */

/* The take's "sensor" block (src/cli/take_loader.zig:186-232): {"clamp":{...},"filter":{"Mitchell":{}}}.
 * Without this call the sensor is libzyg's default: Mitchell radius 2, no clamp (take.zig:59-64). */
int32_t zyg_su_sensor_create(const char* json);
/* su_prop_create for an entity of type "Light", which the scene loader creates un-occluding unless told
 * otherwise (src/util/scene_loader.zig:249,365-366). */
int32_t zyg_su_prop_create_unoccluding(uint32_t shape, uint32_t num_materials, const uint32_t* materials);
/* An entity of type "Instancer" (src/util/scene_loader.zig:401-508, src/core/scene/prop/instancer.zig): `prototypes` are
 * entities made with su_prop_create (they leave the scene's own prop tree, like is_prototype entities of a scene file);
 * instance i places prototypes[prototype_indices[i]] with the 4x4 `transformations + 16 * i` (the layout of
 * su_prop_set_transformation) relative to the returned entity, whose own transformation su_prop_set_transformation sets.
 * The instancer is flattened at compile time: every instance becomes a prop of the two-level device layout. */
int32_t zyg_su_instancer_create(uint32_t num_prototypes, const uint32_t* prototypes, uint32_t num_instances,
                                const uint32_t* prototype_indices, const float* transformations);
/* Thin-lens parameters of the take's camera block (camera_perspective.zig:200-240). */
int32_t zyg_su_camera_set_lens(float aperture_radius, float focus_distance);
/* The take's camera "crop": x0, y0, x1, y1 (x1 / y1 exclusive), clamped like Base.setResolution (camera_base.zig:32-41). Pixel
 * ids and sampler seeds still run over the full resolution (worker.zig:127-141). x1 < 0 restores the full frame. */
int32_t zyg_su_camera_set_crop(int32_t x0, int32_t y0, int32_t x1, int32_t y1);
/* Which builder su_triangle_mesh_create uses: 0 (default) the host's restatement of the reference's SAH /
 * spatial-split build (reference-order tree, identical `primitive` ids), 1 the device LBVH build (zygpu_mesh_build: milliseconds
 * instead of seconds per million triangles; meshes of fewer than 4 triangles still use the host). */
int32_t zyg_su_set_mesh_builder(int32_t builder);
/* Moved vertices of a registered mesh (same topology): refits its trees on the device (zygpu_mesh_refit). `normals` may be NULL.
 * The next su_render_frame / su_start_frame compiles and uploads the scene again. */
int32_t zyg_su_triangle_mesh_refit(uint32_t shape, const float* positions, uint32_t positions_stride, const float* normals,
                                   uint32_t normals_stride);
/* Which builder makes the light trees at compile time: 0 (default) the host's restatement of the reference's
 * builder, 1 the device builder (zygpu_set_light_tree_builder) for trees of at least `min_lights` lights / emissive triangles. */
int32_t zyg_su_set_light_tree_builder(int32_t builder, uint32_t min_lights);
/* CUDA device used by the render calls (default 0). */
int32_t zyg_su_set_device(int32_t ordinal);
/* su_render_frame for a sample range: Driver.render(camera, frame, iteration, num_samples), the CLI's
 * --sample / --num-samples (src/cli/options.zig:88-91). Clears the film first like renderFrameForward. */
int32_t zyg_su_render_frame_range(uint32_t frame, uint32_t iteration, uint32_t num_samples);
/* Runs Scene.compile + camera.update and returns the flattened records (valid until the next su_* call
 * that edits the scene). Needs no GPU. */
int32_t zyg_su_compile(const struct ZygpuScene** scene, const struct ZygpuView** view);
/* The device the engine renders on (a zygpu_device*), NULL before the first render call. */
void* zyg_su_device(void);
/* Read access to the registered meshes by shape id (>= 7). */
const struct zyg_mesh* zyg_su_mesh(uint32_t shape);
/* The image codecs behind su_export_frame on a caller-owned RGBA float image (image/image_writer.zig:15-66). format: 0 PNG,
 * 1 EXR, 2 RGBE; flags: bit 0 alpha channel (PNG / EXR), bit 1 EXR half floats, bit 2 PNG error diffusion; crop = x0, y0, x1, y1
 * (exclusive) or NULL for the full frame. Host only, needs no engine and no GPU. */
/* `it --denoise sigma` (src/it/denoise.zig) on the frame that was just rendered, from the device's own buffers: width * height RGBA
 * like su_resolve_frame_to_buffer. -2 unless su_aovs_create switched on ShadingNormal and Albedo. */
int32_t zyg_su_denoise_frame_to_buffer(float sigma, uint32_t width, uint32_t height, float* buffer);
/* (flags: 1 alpha, 2 half floats (EXR), 4 error diffusion (PNG), bits 8-10 a Writer.Encoding other than colour: 2 Depth, 3 Id,
 * 4 Normal, 5 Float - image_writer.zig:17-24, what su_export_frame uses for the AOV layers) */
int32_t zyg_su_write_image(const char* path, uint32_t format, uint32_t flags, const float* rgba, int32_t width, int32_t height,
                           const int32_t* crop);

#ifdef __cplusplus
}
#endif

#endif /* ZYG_SU_H */
