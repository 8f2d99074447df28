// Wavefront stages of the forward surface-integration pass (PathtracerMIS) for sm_100a.
//
// The per-pixel recursion of the reference (Worker.render -> PathtracerMIS.li, src/core/rendering/worker.zig:
// 104-168, integrator/surface/pathtracer_mis.zig:37-172) is cut into stages that each run over a compacted
// queue of path slots; path state lives in SoA arrays of 16-byte words in HBM:
//
//   generate   camera sample + ray for `samples_in_pass` samples of every (padded) pixel      -> queue A
//   extend     closest hit: prop tree -> analytic shape / 8-wide mesh BVH                      A
//   shade_a    emission at the hit + un-occluding emitter gather, termination, Russian roulette,
//              material setup, light picks and light samples -> shadow-ray records             A -> queue B
//   shadow     any-hit visibility of every shadow-ray record
//   shade_b    NEE contributions of the visible records, BSDF sample, next ray                 B -> queue A
//   film       per-pixel weighted sum over the pass's samples (gather, no atomics)
//
// shade_a / shade_b are split at the shadow ray because the reference draws the light's stochastic
// number only after the shadow test passed (light.zig:127, pathtracer_mis.zig:252-260): keeping that
// order keeps every path on the sampler dimensions the CPU path uses.
#pragma once

#include "../../../include/zygpu_scene.h"
#include "trace.cuh"

#include <cstdint>
#include <cuda_runtime.h>

namespace zygpu {

struct MeshShading {  // hit-point reconstruction data of one mesh
    const uint32_t* triangles;  // 3 per BVH-order triangle
    const float*    positions;  // 3 per vertex
    const uint16_t* normals;    // 2 per vertex, oct snorm16
    const float*    uvs;        // 2 per vertex
    const uint16_t* parts;      // per BVH-order triangle
};

struct MeshSamplerDevice {  // ZygpuMeshSampler with device pointers
    float4                bounds_min, bounds_max;
    uint32_t              num_triangles, num_nodes, two_sided, mesh;
    const ZygpuLightNode* nodes;
    const uint32_t*       node_middles;
    const uint32_t*       light_orders;
    const uint32_t*       light_mapping;
    const uint32_t*       triangle_mapping;
    const float*          triangle_pdfs;
    const uint32_t*       primitive_mapping;
    const float4*         triangle_props;  // per light triangle {centre, radius}, {normal, pdf}: MeshImpl.lightProperties, filled at upload
};

// ZygpuImageSampler on the device: the emission image and the cdf rows of its Distribution2D.
struct ImageSamplerDevice {
    uint32_t     width, height, address_u, address_v, filter;
    float        total_weight, scale_u, scale_v;
    const float* pixels;
    const float* marginal_cdf;
    const float* conditional_cdf;
};

struct SceneDevice {
    const ZygpuProp*     props;
    const float4*        trafos;  // 4 per prop
    const float4*        aabbs;   // 2 per prop
    const uint32_t*      material_ids;
    const uint32_t*      light_ids;
    const ZygpuMaterial* materials;

    const ZygpuLight* lights;
    const float4*     light_aabbs;  // 2 per light
    const float4*     light_cones;

    // light tree
    const ZygpuLightNode* lt_nodes;
    const uint32_t*       lt_middles;
    const uint32_t*       lt_orders;
    const uint32_t*       lt_mapping;
    float4                lt_bounds_min, lt_bounds_max;
    float                 lt_infinite_weight, lt_infinite_guard;
    uint32_t              lt_infinite_end, lt_max_split_depth, lt_num_infinite, lt_num_nodes;
    const float*          lt_infinite_cdf;  // Tree.infinite_light_distribution.cdf

    const float4*   solid_nodes;  // 2 per node
    const uint32_t* solid_indices;
    uint32_t        num_solid_nodes;
    const float4*   tlas_nodes;  // the solid prop tree in the 8-wide device layout (host/wide_bvh.hpp): 5 per node
    const float4*   tlas_recs;   // prop records of its leaf slots: 2 per record
    const float4*   unocc_nodes;
    const uint32_t* unocc_indices;
    uint32_t        num_unocc_nodes;

    const MeshSamplerDevice* mesh_samplers;    // by ZygpuLight.sampler
    uint32_t                 num_mesh_samplers;
    const float*             mesh_part_areas;  // Part.area per part entry

    const ImageSamplerDevice* image_samplers;  // by ZygpuMaterial.emission_map / ZygpuLight.sampler (PROP_IMAGE lights)

    const uint32_t* infinite_props;  // Scene.infinite_props (Distant, Canopy): met only by rays that leave the scene
    uint32_t        num_infinite_props;

    const MeshDevice*  meshes;
    const MeshShading* mesh_shading;

    const float* luts;  // ZygpuScene.ggx_luts

    // What a ray's traversal stack can hold at most in the fused kernels: two entries (node group, postponed leaf group) per level of the
    // prop tree and of the deepest mesh tree, the parked world ray (5), one spare. The one-ray-per-lane kernel keeps kWideStack (48)
    // entries in local memory, the ray-pool kernel kScenePoolStack in global scratch; a scene beyond either takes the next variant.
    uint32_t trace_stack_bound;

    float4 world_lo;     // lower corner of the box around the finite props and, per axis, cells per unit length: the ray-sort grid
    float4 world_cells;  // (device/render_trace.cu, sortKeyKernel)
};

constexpr uint32_t kSortBins = 1u << 16;
constexpr uint32_t kScenePoolStack      = 96;  // traversal stack entries per pooled ray (global scratch: depth is cheap there)
constexpr uint32_t kScenePoolStackWords = 4u * 64u * kScenePoolStack;  // per block: 4 warps x 64 pooled rays  // keys of the ray sort (device/render_trace.cu)

struct PathState {
    // Per path vertex. A camera sample ("slot") owns `lanes` vertex records: 1 when no material of the scene can split a
    // path, 4 (= Pool.NumVertices, vertex.zig:216) when one can (Glass). Vertex id = lane * capacity + slot.
    float4* ray_o;   // origin xyz | flags: state bits 0-7, probe depth 8-15, vertex.depth 16-23, log2(path_count) 24-25, media 26-27
    float4* ray_d;   // direction xyz | max_t (after extend: hit t)
    float4* thr;     // throughput rgb | bxdf_pdf
    float4* prev_p;  // vertex.origin xyz | reg_alpha
    float4* prev_n;  // vertex.geo_n xyz | split_weight
    float4* hit;     // u, v | primitive | prop
    uint4*  med;     // medium stack (prop/medium.zig:30-153): props of the up to 3 entries | their parts, 8 bits each; null when lanes == 1

    // Per camera sample.
    float4* acc_e;   // emission rgb | pixel_uv.x
    float4* acc_d;   // direct rgb | pixel_uv.y
    float4* acc_i;   // indirect rgb | alpha of the sample (Pool.transparency[3], vertex.zig:243-268; views with alpha_transparency)
    uint4*  smp;     // Sobol: block seed, run seed, dimension | vertex-pool word (lanes of the current / next generation)
    uint2*  rng;     // PCG state

    // shadow-ray records, written by shade_a: path `slot` owns records [slot * shadow_stride, +sh_n[slot])
    float4*   sh_o;   // origin xyz | light pdf (sample pdf * pick pdf)
    float4*   sh_p;   // offset light position xyz | light id (bit 31: infinite light, the ray runs along wi to RayMaxT)
    float4*   sh_wi;  // light_sample.wi xyz | visible (written by shadow)
    uint32_t* sh_n;   // per path: number of records
    float2*   sh_uv;  // per record: uvw of a light sample taken through the light's emission image (Rectangle.sampleMaterialTo), read
                      // by Light.evaluateTo in shade_b; null unless the scene has an image-mapped finite light
    // AOV values of the camera sample (Worker.commonAOV, worker.zig:209-242; aov.Value): null unless the view records an AOV class
    float4* aov_albedo;  // throughput * mat_sample.aovAlbedo() of the last primary-ray vertex
    float4* aov_gn;      // geometric normal of the first hit
    float4* aov_sn;      // shading normal of the first hit
    float4* aov_misc;    // roughness | depth (FLT_MAX: nothing hit) | 1 + material id | -

    float*    stoch;  // per vertex: rs.stochastic_r of Vertex.sample (vertex.zig:165), drawn by shade_a and read again by shade_b
                      // for the material's image lookups; null when no material of the scene reads an image per vertex

    // deferred light sampling (scenes with many lights; null otherwise): shade_a leaves the request, the persistent
    // light kernels turn it into picks and shadow records
    float4*   ls_p;     // shading point xyz | the light-selection random number
    float4*   ls_g;     // Fragment.geo_n xyz | bit 0 translucent, bit 1 the material sample's geo_n is -geo_n, bit 2 low split
                        // threshold, bits 8-15 total depth
    uint2*    picks;    // 64 per slot: light id, pdf
    uint32_t* pick_n;   // per slot
    uint32_t* queue_l;  // slots with a request

    // mesh candidates collected by the top kernels, per trace item (closest: path slot, shadow: record)
    uint32_t* ml_props;  // 8 per item
    uint32_t* ml_count;
    uint32_t* queue_m;  // items with candidates (fused traversal kernel: the trace items in ray-sort order)
    uint32_t* sort_bins; // ray sort: one counter per key (kSortBins + 1), null when the pass does not sort
    uint2*    trace_stacks;        // ray-pool traversal kernel: kScenePoolStackWords entries per resident block (device/render_trace.cu)
    uint32_t  trace_stack_blocks;  // blocks the scratch array is sized for

    uint32_t* queue_a;   // slots with at least one vertex in the current generation
    uint32_t* queue_b;   // slots whose vertex of the current round survived shade_a
    uint32_t* queue_t;   // lanes > 1: vertex ids to extend (lanes == 1: the slots of queue_a are the vertex ids)
    uint32_t* queue_s;   // lanes > 1: slots with more than one vertex in the current generation
    uint32_t* queue_r;   // shadow_stride > 1: the shadow records written by shade_a, compacted (null: every slot has one record)
    unsigned long long* tally;  // instrumented passes only (else null), 12 words: node / triangle / prop-record fetches and warp-level
                                // NODE / TRIANGLE / PROP steps of the closest-hit ([0..5]) and shadow ([6..11]) traversal — the
                                // counted bytes behind the render-path roofline and the lanes-per-step of the lock-step loop
    uint32_t* counters;  // [0] |A|, [1] |B|, [2] |mesh queue|, [3] shadow overflow flag, [4] |next A|, [5] closest rays, [6] shadow rays,
                         // [7] |T|, [8] work counter of the persistent mesh kernel, [9] |S|, [10] |R|, [11] |L|, [12] / [13] work counters of the
                         // persistent light kernels, [15] sorted trace items (16 words in all)

    uint32_t capacity;       // path slots
    uint32_t shadow_stride;  // shadow records reserved per path
    uint32_t lanes;          // vertex records per slot: 1 or 4
};

struct PassParams {
    uint32_t iteration;        // first sample of the pass
    uint32_t samples_in_pass;  // k
    uint32_t padded_w, padded_h;
    uint32_t num_paths;  // k * padded_w * padded_h
    uint32_t debug_slot; // ZYGPU_DEBUG_SLOT: the shade stages print the vertices of this slot (diagnostics); 0xFFFFFFFF = off
};

int         numSmsOfCurrentDevice();
cudaError_t uploadSobolDirections();
uint32_t    sceneTraceLaunches(bool has_meshes, uint32_t num_solid_nodes);  // kernels one extend / shadow stage launches (for the launch statistics)

cudaError_t launchGenerate(const ZygpuView& view, const PathState& st, const PassParams& pass, cudaStream_t stream);
// The queue lengths live on the device; the grids are sized for `max_items` and exit early.
// `bounce`: the path depth of the stage (the ray sort skips the camera rays, which arrive in pixel order).
cudaError_t launchExtend(const SceneDevice& scene, const PathState& st, uint32_t max_items, bool has_meshes, uint32_t bounce, cudaStream_t stream);
// Diagnostics (ZYGPU_VERIFY_TRACE=1): the closest-hit stage once more with one thread per ray in the reference's prop order, and a
// comparison of two result sets (prints the rays whose hits differ).
cudaError_t launchExtendReference(const SceneDevice& scene, const PathState& st, uint32_t max_items, cudaStream_t stream);
cudaError_t launchCompareHits(const PathState& st, const float4* ray_d_before, const float4* ray_d_a, const float4* hit_a, uint32_t max_items,
                              uint32_t bounce, cudaStream_t stream);
// `round` = which vertex of each slot's current generation the shade stages work on: the vertices of one camera sample share
// its sampler and are processed in the reference's order (VertexPool.consume, vertex.zig:232-283), so rounds are sequential.
cudaError_t launchBeginGeneration(const PathState& st, cudaStream_t stream);
cudaError_t launchBeginRound(const PathState& st, cudaStream_t stream);
cudaError_t launchEndGeneration(const PathState& st, cudaStream_t stream);
cudaError_t launchShadeA(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass,
                         uint32_t max_items, uint32_t round, cudaStream_t stream);
// Deferred light sampling between shade_a and the shadow stage (PathState.queue_l non-null).
cudaError_t launchLightStages(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass, uint32_t max_items,
                              cudaStream_t stream);
cudaError_t launchShadow(const SceneDevice& scene, const PathState& st, uint32_t max_items, bool has_meshes, uint32_t bounce, cudaStream_t stream);
cudaError_t launchShadeB(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass,
                         uint32_t max_items, uint32_t round, cudaStream_t stream);
// `film_alpha`: the alpha lane of the Transparent buffer (sum of weight * alpha per pixel, buffer_transparent.zig:45-54), or null
cudaError_t launchFilm(const ZygpuView& view, const PathState& st, const PassParams& pass, float4* film, float* film_alpha, cudaStream_t stream);
// The AOV layers of the sensor (aov.Buffer, rendering/sensor/aov/aov_buffer.zig): one Pack4f image per active class.
struct AovFilm {
    float4* layers[9];  // by aov.Value.Class; null = inactive
};
cudaError_t launchAovClear(const AovFilm& aov, uint32_t num_pixels, cudaStream_t stream);  // aov.Buffer.clear
// MeshImpl.lightProperties (shape_sampler.zig:198-226) of every light triangle of `sampler` into props[2 * num_triangles]
cudaError_t launchMeshLightProps(const MeshDevice& mesh, const MeshSamplerDevice& sampler, float4* props, cudaStream_t stream);
cudaError_t launchAovFilm(const ZygpuView& view, const PathState& st, const PassParams& pass, const AovFilm& aov, cudaStream_t stream);
cudaError_t launchResolveAov(uint32_t aov_class, const float4* layer, float4* rgba, uint32_t num_pixels, cudaStream_t stream);
// The `it` tool's denoise operator (src/it/denoise.zig:137-246, 375-451) as a post kernel over the film and the ShadingNormal / Albedo
// layers still on the device. `weights`: the normalised (2 radius + 1)^2 Gaussian of Denoise.init (:34-72).
cudaError_t launchDenoise(const ZygpuView& view, const float4* film, const float4* normal, const float4* albedo, const float* weights, int32_t radius,
                          float4* rgba, cudaStream_t stream);
cudaError_t launchResolve(const ZygpuView& view, const float4* film, const float* film_alpha, float4* rgba, uint32_t num_pixels, cudaStream_t stream);

}  // namespace zygpu
