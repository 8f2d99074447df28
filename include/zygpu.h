/* zygpu.h — device ABI of the B200 surface-integration backend for zyg.
 *
 * This is the boundary a Zig host binds with `extern fn` / @cImport: plain C,
 * pointers and sizes only. zyg has no device API of its own; each entry point names the reference
 * code whose role it takes over.
 *
 * Conventions (same as src/capi/capi.zig): every function returns 0 on success (or a non-negative
 * id) and -1 on failure; zygpu_last_error() holds the message. One caller thread per device.
 * Buffers are caller-owned; the library copies what it needs.
 */
#ifndef ZYGPU_H
#define ZYGPU_H

#include <stdint.h>

#include "zygpu_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- host-side scene compile ------------------------------------------------------------- */

/* A compiled triangle mesh: reference-order binary BVH + the derived device layout.
 * Replaces shape_provider.zig:847-924 (buildDescAsync/buildBVH) + triangle_tree_builder.zig:33-65. */
typedef struct zyg_mesh zyg_mesh;

/* Argument meaning follows su_triangle_mesh_create (src/capi/capi.zig:379-423): `parts` holds
 * num_parts x {start_index, num_indices, material}; strides are in floats; indices may be NULL
 * (then triangle i uses vertices 3i..3i+2); normals / uvs may be NULL. num_threads = 0 uses all
 * host cores (the result never depends on it). */
int zyg_mesh_build(uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles, const uint32_t* indices,
                   uint32_t num_vertices, const float* positions, uint32_t positions_stride, const float* normals,
                   uint32_t normals_stride, const float* uvs, uint32_t uvs_stride, uint32_t num_threads,
                   zyg_mesh** out);
void zyg_mesh_free(zyg_mesh* mesh);

enum {
    ZYG_MESH_BINARY_NODES = 0, /* 32-byte nodes, src/core/scene/bvh/node.zig:9-20 */
    ZYG_MESH_TRIANGLES    = 1, /* u32[3] per BVH-order triangle, triangle.zig:6-10 */
    ZYG_MESH_ORIGINAL     = 2, /* u32 per BVH-order triangle: index in the caller's triangle list */
    ZYG_MESH_POSITIONS    = 3, /* f32[3] per vertex + 1 pad float, triangle_data.zig:49,54 */
    ZYG_MESH_NORMALS      = 4, /* u16[2] per vertex, oct-encoded snorm16 */
    ZYG_MESH_UVS          = 5, /* f32[2] per vertex */
    ZYG_MESH_PARTS        = 6, /* u16 per BVH-order triangle */
    ZYG_MESH_WIDE_NODES   = 7, /* 96-byte device nodes (80 bytes + padding to three 32-byte loads) */
    ZYG_MESH_WIDE_TRIS    = 8  /* 64-byte device triangle records */
};

/* Read-only view of one of the arrays above; valid until zyg_mesh_free. */
const void* zyg_mesh_data(const zyg_mesh* mesh, int which, uint64_t* num_bytes);

typedef struct ZygMeshInfo {
    uint32_t num_source_triangles;
    uint32_t num_tree_triangles; /* >= source: spatial splits duplicate references */
    uint32_t num_vertices;
    uint32_t num_binary_nodes;
    uint32_t num_wide_nodes;
    uint32_t wide_max_depth;
    uint32_t num_degenerate_leaves;
    uint32_t num_leaf_order_fixups;
    float    aabb_min[3];
    float    aabb_max[3];
} ZygMeshInfo;

int zyg_mesh_info(const zyg_mesh* mesh, ZygMeshInfo* info);

/* ---- device -------------------------------------------------------------------------------- */

typedef struct zygpu_device zygpu_device;

int  zygpu_create(int device_ordinal, zygpu_device** out);
void zygpu_destroy(zygpu_device* dev);

/* Uploads the compiled mesh (wide layout + reference layout). Returns a mesh id >= 0. */
int zygpu_upload_mesh(zygpu_device* dev, const zyg_mesh* mesh);

/* ---- device BVH build / refit ------------------------------------------------
 * zyg_mesh_build on the device: replaces builder_base.zig:65-283 + triangle_tree_builder.zig:33-207 for callers that want the
 * first pixel sooner than the SAH / spatial-split build allows. The tree is an LBVH (60-bit Morton order, Karras' binary radix
 * tree, leaves of at most three triangles) collapsed into the same 8-wide quantised nodes; triangles are not duplicated, so
 * num_tree_triangles == num_source_triangles. The handle is an ordinary zyg_mesh (arrays copied back to the host): closest hits
 * are the reference's up to equal-t ties and the leaf-box gate (a hit whose ray misses the box of its leaf is not reported, by
 * either tree; the boxes differ). Needs at least 4 triangles. `device_ms` (may be NULL): CUDA-event time of the build. */
int zygpu_mesh_build(zygpu_device* dev, uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles, const uint32_t* indices,
                     uint32_t num_vertices, const float* positions, uint32_t positions_stride, const float* normals,
                     uint32_t normals_stride, const float* uvs, uint32_t uvs_stride, zyg_mesh** out, float* device_ms);
/* Same topology, moved vertices: refits the wide tree level by level from the leaves (triangle records, leaf gates, quantised
 * child boxes), the boxes of the binary tree and the bounding sphere, on the device copy of the mesh (uploaded if it was not)
 * and in the handle. `normals` may be NULL (kept). A scene that uses the mesh has to be compiled and uploaded again. */
int zygpu_mesh_refit(zygpu_device* dev, zyg_mesh* mesh, const float* positions, uint32_t positions_stride, const float* normals,
                     uint32_t normals_stride, float* device_ms);

/* ---- device light-tree build --------------------------------------------------
 * Replaces light_tree_builder.zig:281-428 (Builder.build for the scene's lights, Builder.buildPrimitive for the emissive triangles
 * of a mesh part, i.e. the expensive half of Scene.propPrepareSampling, scene.zig:402-497) for trees of at least `min_lights`
 * lights: while a device is installed, the host scene compile (zyg_su_compile / su_render_frame) and the mesh-light sampler build
 * hand those trees to the GPU: Morton order over the light centres, Karras' radix tree, bounds / cone / power / variance / two-
 * sidedness aggregated bottom-up, serialised into the reference's 32-byte light_tree.Node. The estimator is unchanged, the pdfs are
 * those of another valid tree than the reference's cost-driven one. dev = NULL uninstalls. Process-global, like the su_* engine. */
int zygpu_set_light_tree_builder(zygpu_device* dev, uint32_t min_lights);
/* CUDA-event time spent in device light-tree builds since the last reset. */
float zygpu_light_tree_build_ms(int reset);

/* 32-byte ray: origin, min_t, direction, max_t (object space of the mesh; src/base/math/ray.zig). */
typedef struct ZygpuRay {
    float origin[3];
    float min_t;
    float direction[3];
    float max_t;
} ZygpuRay;

/* 16-byte closest-hit record (src/core/scene/shape/intersection.zig:46-61). primitive is the
 * BVH-order triangle index (0xFFFFFFFF = miss, then t = max_t). */
typedef struct ZygpuHit {
    float    t, u, v;
    uint32_t primitive;
} ZygpuHit;

enum {
    ZYGPU_CLOSEST        = 0, /* TriangleTree.intersect,  triangle_tree.zig:46-109 — wide BVH */
    ZYGPU_ANY            = 1, /* TriangleTree.intersectP, triangle_tree.zig:197-242 — wide BVH */
    ZYGPU_CLOSEST_BINARY = 2, /* same, visiting the 32-byte reference nodes in reference order */
    ZYGPU_ANY_BINARY     = 3
};

typedef struct ZygpuTraceCounters {
    uint64_t nodes;     /* node records fetched (80 B wide / 32 B binary) */
    uint64_t triangles; /* triangle tests */
    uint64_t rays;
    uint64_t max_stack;
} ZygpuTraceCounters;

/* Host buffers: copies rays in, traces, copies results out (chunked and overlapped on internal
 * streams). `out` is ZygpuHit[n] for the closest modes and uint32_t[n] (1 = occluded) for any-hit. */
int zygpu_trace_batch(zygpu_device* dev, int mesh, int mode, const ZygpuRay* rays, uint64_t n, void* out);

/* Device buffers, asynchronous on `stream` (a CUstream / cudaStream_t; NULL = default stream). Calls on one device share a
 * scratch region for the traversal stacks: launches of different calls must not overlap on the device (use one stream).
 * If `counters` is non-NULL the instrumented kernel runs, the call synchronises the stream and
 * writes fetch counts for this batch. */
int zygpu_trace_batch_device(zygpu_device* dev, int mesh, int mode, const void* d_rays, uint64_t n, void* d_out,
                             void* stream, ZygpuTraceCounters* counters);

/* ---- forward surface-integration pass --------------------------------------------------------
 * Replaces Driver.renderFrameForward / renderFrameIterationForward (src/core/rendering/driver.zig:309-348):
 * the host compiles its scene as before (Scene.compile, camera.update), hands the flattened result over,
 * and the per-pixel PathtracerMIS recursion runs as wavefront stages on the device. */

/* Copies the compiled scene to the device (meshes referenced by the scene are uploaded on first use). */
int zygpu_upload_scene(zygpu_device* dev, const ZygpuScene* scene);
/* Camera, integrator, sampler and sensor settings; (re)allocates the film when the resolution changes. */
int zygpu_set_view(zygpu_device* dev, const ZygpuView* view);
/* Opaque.clear(0), buffer_opaque.zig:23-27. */
int zygpu_clear_film(zygpu_device* dev);
/* Adds samples [iteration, iteration + num_samples) of every pixel to the film: Driver.renderIterations
 * (driver.zig:182-187) called once per sample, which is the progressive API's schedule and reseeds the PCG
 * stream of the depth >= 3 sampler per sample (worker.zig:143). Asynchronous on the device's render stream. */
int zygpu_render(zygpu_device* dev, uint32_t iteration, uint32_t num_samples);
/* Sensor.resolveTonemap (Linear) into a host RGBA fp32 buffer of num_pixels pixels; synchronises. */
int zygpu_resolve(zygpu_device* dev, float* rgba, uint32_t num_pixels);
/* aov.Buffer.resolve (rendering/sensor/aov/aov_buffer.zig:51-82) of one AOV class (ZYG_AOV_*) the view records (ZygpuView.aov_slots):
 * colours are |rgb| / weight in sRGB primaries, normals rgb / weight, Roughness x / weight, Depth and MaterialId the stored value;
 * alpha 1. Returns -2 when the class is not active (Driver.resolveAovToBuffer, driver.zig:209-217). `download_layer` != 0 copies
 * the unresolved Pack4f layer instead. The AOV layers stay on the device that rendered them: zygpu_reduce_film sums the beauty only. */
int zygpu_resolve_aov(zygpu_device* dev, uint32_t aov_class, float* rgba, uint32_t num_pixels, int download_layer);
/* The `it` tool's denoise operator (src/it/denoise.zig; `it --denoise sigma`, options.zig:98-100) as a post kernel on the device: the
 * beauty filtered with the normalised Gaussian of sigma (radius ceil(3 sigma)), each tap blended in by normal agreement, albedo
 * distance and the local noise estimate; sRGB primaries, alpha 1, like the tool's output. Needs the ShadingNormal and Albedo classes
 * in ZygpuView.aov_slots (the "_n" and "_albedo" files the tool looks for, operator.zig:70-93): -2 otherwise. Synchronises. */
int zygpu_denoise(zygpu_device* dev, float sigma, float* rgba, uint32_t num_pixels);
/* The weighted-sum film itself: Pack4f per pixel (sum w*rgb, sum w), buffer_opaque.zig:12. */
int   zygpu_download_film(zygpu_device* dev, float* film, uint32_t num_pixels);
int   zygpu_upload_film(zygpu_device* dev, const float* film, uint32_t num_pixels);
/* Device pointer of the film for the multi-GPU reduce (one ncclReduce(sum, fp32) per frame). */
void* zygpu_film_device(zygpu_device* dev, uint64_t* num_floats);
int   zygpu_synchronize(zygpu_device* dev);
/* The film combine of the multi-GPU split: one ncclReduce(sum, fp32) of the W*H*4 film to `root`, enqueued
 * on the render stream behind the passes. `nccl_comm` is the caller's ncclComm_t; the library resolves ncclReduce from the NCCL
 * already loaded in the process, it does not link one. */
int   zygpu_reduce_film(zygpu_device* dev, void* nccl_comm, int root);
/* CUDA ordinal the device was created on. */
int   zygpu_device_ordinal(zygpu_device* dev);
/* Progress of the asynchronous zygpu_render calls since zygpu_create: passes enqueued and passes the device has finished
 * (Progressor.tick granularity of the device path; driver.zig:305 ticks per tile). */
uint32_t zygpu_passes_enqueued(zygpu_device* dev);
uint32_t zygpu_passes_completed(zygpu_device* dev);
/* The CUDA stream (cudaStream_t) the render calls are enqueued on, for events and for ordering a collective after a
 * pass. NULL before the first zygpu_upload_scene / zygpu_set_view. */
void* zygpu_render_stream(zygpu_device* dev);

typedef struct ZygpuRenderStats {
    uint64_t camera_samples;  /* path samples started */
    uint64_t closest_rays;    /* Scene.intersect calls */
    uint64_t shadow_rays;     /* Scene.visibility calls */
    uint64_t kernel_launches; /* launches of this library's kernels */
    uint64_t passes;
    uint64_t overflow_retries; /* passes run again with a larger shadow-record reservation */
} ZygpuRenderStats;
/* Totals since the last zygpu_clear_film; synchronises the render stream. */
int zygpu_render_stats(zygpu_device* dev, ZygpuRenderStats* stats);

/* Instrumented render traversal (the counted fetches behind the roofline of the render path, SURVEY.md §8d "B_sample"): while
 * counting is on the closest-hit and shadow traversal kernels count their 80-byte node, 64-byte triangle-record and 32-byte
 * prop-record fetches and their warp-level steps. zygpu_set_counting(dev, on) also zeroes the counters. */
typedef struct ZygpuTraversalCounts {
    uint64_t nodes, triangles, props;                 /* records fetched (per lane) */
    uint64_t node_steps, triangle_steps, prop_steps;  /* warp-level lock-step steps of each kind */
} ZygpuTraversalCounts;
int zygpu_set_counting(zygpu_device* dev, int on);
int zygpu_traversal_counts(zygpu_device* dev, ZygpuTraversalCounts* closest, ZygpuTraversalCounts* shadow);

const char* zygpu_last_error(void);

#ifdef __cplusplus
}
#endif

#endif /* ZYGPU_H */
