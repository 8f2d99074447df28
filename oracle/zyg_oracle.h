/* ORACLE — test infrastructure only. CPU restatement of the reference (Opioid/zyg) hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. The product (zyg_b200/) never does.
 *
 * PARITY UNPINNED for geometry and shading: the reference has no automated tests, golden vectors or
 * fixtures and cannot be built here (Zig, no toolchain). Pins that do exist:
 * the published PCG32 known-answer vector (pcg32.cpp) and structural self-checks in tests/.
 */
#ifndef ZYG_ORACLE_H
#define ZYG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ZoRay {
    float origin[3];
    float min_t;
    float direction[3];
    float max_t;
} ZoRay;

typedef struct ZoHit {
    float    t, u, v;
    uint32_t primitive;
} ZoHit;

/* TriangleTree.intersect (triangle_tree.zig:46-109), identity transformation. threads = 0: all cores. */
void zo_trace_closest(const void* nodes, const uint32_t* triangles, const float* positions, const ZoRay* rays,
                      uint64_t n, ZoHit* out, uint32_t threads, uint64_t* visited_nodes, uint64_t* tested_tris);

/* TriangleTree.intersectP (triangle_tree.zig:197-242). out[i] = 1 if occluded. */
void zo_trace_any(const void* nodes, const uint32_t* triangles, const float* positions, const ZoRay* rays, uint64_t n,
                  uint32_t* out, uint32_t threads);

/* O(N) closest hit over all triangles with the reference's accept rule; no tree involved. */
void zo_brute_closest(const uint32_t* triangles, uint32_t num_triangles, const float* positions, const ZoRay* rays,
                      uint64_t n, ZoHit* out, uint32_t* num_ties, uint32_t threads);

/* rnd.Generator (src/base/random/generator.zig:1-47): start(state, sequence) then n draws. */
void zo_pcg32_uints(uint64_t state, uint64_t sequence, uint32_t n, uint32_t* out);
void zo_pcg32_floats(uint64_t state, uint64_t sequence, uint32_t n, float* out);

/* ---- forward surface-integration pass (render.cpp) ---- */
struct ZygpuScene;
struct ZygpuView;

/* Adds samples [iteration, iteration + num_samples) of every pixel to `film` (Pack4f per pixel of the full
 * resolution, weight sum in w). per_sample_iterations != 0: num_samples calls of (iteration + k, 1), the
 * progressive API's schedule, which is the one the device implements. threads = 0: all cores. */
/* Host arrays of one compiled triangle mesh in the reference's layout (what zyg_mesh_data returns). */
typedef struct ZoMesh {
    const void*     nodes;     /* 32-byte bvh.Node */
    const uint32_t* triangles; /* 3 per tree-order triangle */
    const float*    positions; /* 3 per vertex */
    const uint16_t* normals;   /* 2 per vertex, oct snorm16 */
    const float*    uvs;       /* 2 per vertex */
    const uint16_t* parts;     /* per tree-order triangle */
} ZoMesh;

/* `meshes` is indexed by ZygpuProp.mesh (NULL when the scene has none). */
void zo_render(const struct ZygpuScene* scene, const struct ZygpuView* view, const ZoMesh* meshes, uint32_t iteration,
               uint32_t num_samples, int per_sample_iterations, float* film, uint32_t threads);
/* zo_render plus the AOV layers of view->aov_slots (Worker.commonAOV, worker.zig:209-242; Sensor.addSample's AOV half,
 * sensor.zig:197-377): aov_layers[c] = Pack4f image of class c (ZYG_AOV_*), cleared by the caller to the class default
 * (aov.Buffer.clear), null for inactive classes. */
void zo_render_aov(const struct ZygpuScene* scene, const struct ZygpuView* view, const ZoMesh* meshes, uint32_t iteration,
                   uint32_t num_samples, int per_sample_iterations, float* film, float* const* aov_layers, uint32_t threads);
/* ... and the alpha lane of the Transparent sensor buffer (buffer_transparent.zig; Pool.transparency, vertex.zig:243-268): `alpha` = one
 * float per pixel, sum of weight * alpha, not cleared; null = Opaque. */
void zo_render_layers(const struct ZygpuScene* scene, const struct ZygpuView* view, const ZoMesh* meshes, uint32_t iteration,
                      uint32_t num_samples, int per_sample_iterations, float* film, float* const* aov_layers, float* alpha,
                      uint32_t threads);
/* Transparent.resolveTonemap (buffer_transparent.zig:82-93): zo_resolve with alpha = |alpha sum / weight|. */
void zo_resolve_transparent(const struct ZygpuView* view, const float* film, const float* alpha, uint32_t num_pixels, float* rgba);
/* aov.Buffer.resolve (aov_buffer.zig:51-82) of one class. */
void zo_resolve_aov(uint32_t aov_class, const float* layer, uint32_t num_pixels, float* rgba);
/* The `it` tool's denoise operator (src/it/denoise.zig:137-246, 375-451) over the unresolved film and the unresolved ShadingNormal and
 * Albedo layers of a view (the colour in AP1, as the tool's loader leaves it): rgba = sRGB primaries, alpha 1. */
void zo_denoise(const struct ZygpuView* view, const float* film, const float* normal_layer, const float* albedo_layer, float sigma, float* rgba);
/* 0 (default): sampler draws in the reference's order. 1: the draws of PathtracerMIS.sampleLights regrouped the way the
 * device takes them (all light samples of a vertex, then one draw per visible sample); identical when a vertex takes one
 * light sample. */
void zo_set_wavefront_light_order(int on);
/* Tree.randomLight / Tree.pdf (light_tree.zig:346-517) over the compiled scene's light tree. picks = (light id, pdf) pairs. */
uint32_t zo_light_tree_random(const struct ZygpuScene* scene, const struct ZygpuView* view, const float p[3], const float n[3],
                              int total_sphere, float random, float split_threshold, float* picks);
float    zo_light_tree_pdf(const struct ZygpuScene* scene, const struct ZygpuView* view, const float p[3], const float n[3],
                           int total_sphere, float split_threshold, uint32_t light);
/* shape_sampler.ImageImpl of ZygpuScene.image_samplers[index]: sample (r2 pairs -> u, v, pdf triples), pdf (uv pairs),
 * and the texture lookup ts.sample2D_3 (u, v, stochastic_r triples -> rgb). */
void zo_image_sample(const struct ZygpuScene* scene, uint32_t index, uint32_t n, const float* r2, float* uv_pdf);
void zo_image_pdf(const struct ZygpuScene* scene, uint32_t index, uint32_t n, const float* uv, float* pdf);
void zo_image_texel(const struct ZygpuScene* scene, uint32_t index, uint32_t n, const float* uvr, float* rgb);
/* Opaque.resolveTonemap, Linear tonemapper. */
void zo_resolve(const struct ZygpuView* view, const float* film, uint32_t num_pixels, float* rgba);

float zo_ggx_micro_directional_albedo(float alpha, float n_dot_wo, uint32_t num_samples);
float zo_ggx_f_s_ss(float alpha, float f0, float ior_t, float n_dot_wo, uint32_t num_samples);
/* integrate_directional_albedo / integrate_average_albedo of the same generator (ggx_integrate.zig:89-132) over the shipped
 * tables (`luts` = ZygpuScene.ggx_luts). */
float zo_ggx_directional_albedo(const float* luts, float alpha, float f0, float n_dot_wo, uint32_t num_samples);
float zo_ggx_average_albedo(const float* luts, float alpha, float f0, uint32_t num_samples);
float zo_ggx_micro_average_albedo(const float* luts, float alpha, uint32_t num_samples); /* :59-73 */
void  zo_sobol_stream(uint32_t sample, uint32_t seed, uint32_t n, uint32_t pad_every, float* out);
void  zo_sobol_directions(uint32_t* out160);

/* ---- the oracle's own scene-compile builders (builders.cpp) -------------------------------------------------------
 * Independent restatement of builder_base.zig / split_candidate.zig / triangle_tree_builder.zig / prop_tree_builder.zig /
 * light_tree_builder.zig / Part.configure: tests compare every array with what the product's host builds (byte for byte),
 * bench.py --impl reference builds its tree here. Handles are freed with zo_build_free; the *_data views are valid until then. */
struct ZygpuAabb;
void        zo_build_free(void* handle);
void*       zo_mesh_build(uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles, const uint32_t* indices, uint32_t num_vertices,
                          const float* positions, uint32_t positions_stride, const float* normals, uint32_t normals_stride,
                          const float* uvs, uint32_t uvs_stride, uint32_t threads);
const void* zo_mesh_data(const void* handle, int which, uint64_t* num_bytes);
void*       zo_prop_tree_build(const uint32_t* indices, uint32_t num_indices, const struct ZygpuAabb* aabbs, uint32_t threads);
const void* zo_prop_tree_data(const void* handle, int which, uint64_t* num_bytes);
void*       zo_light_tree_build(uint32_t num_lights, const struct ZygpuAabb* aabbs, const float* cones, const uint8_t* two_sided,
                                const uint8_t* finite);
const void* zo_light_tree_data(const void* handle, int sampler, int which, uint64_t* num_bytes);
void*       zo_mesh_sampler_build(const ZoMesh* mesh, uint32_t num_tree_triangles, uint32_t num_parts, uint32_t part, int two_sided);
const void* zo_mesh_sampler_data(const void* handle, int which, uint64_t* num_bytes);

#ifdef __cplusplus
}
#endif

#endif
