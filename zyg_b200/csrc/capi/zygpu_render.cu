// Render entry points of the device ABI (include/zygpu.h): scene / view upload and the pass loop that
// strings the wavefront stages of device/render.cu together. Replaces Driver.renderFrameForward /
// renderFrameIterationForward (src/core/rendering/driver.zig:309-348).
#include "device_state.hpp"

#include "../host/mesh_handle.hpp"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <dlfcn.h>

namespace {

template <typename T>
int uploadArray(RenderState& r, const T* host, size_t count, const T** device) {
    const size_t bytes = std::max<size_t>(count * sizeof(T), 128);
    const size_t slot  = r.scene_buffer_cursor++;
    if (slot == r.scene_buffers.size()) {
        r.scene_buffers.push_back(nullptr);
        r.scene_buffer_bytes.push_back(0);
    }
    if (r.scene_buffer_bytes[slot] < bytes) {  // grow with some slack so that a scene that changes a little keeps its buffers
        cudaFree(r.scene_buffers[slot]);
        r.scene_buffers[slot]      = nullptr;
        r.scene_buffer_bytes[slot] = 0;
        const size_t capacity      = bytes + bytes / 8;
        CUDA_OK(cudaMalloc(&r.scene_buffers[slot], capacity));
        r.scene_buffer_bytes[slot] = capacity;
    }
    void* d = r.scene_buffers[slot];
    if (count > 0 && host) CUDA_OK(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));  // no host array: the caller fills it
    *device = static_cast<const T*>(d);
    return 0;
}

void freeAll(std::vector<void*>& buffers) {
    for (void* b : buffers) cudaFree(b);
    buffers.clear();
}

template <typename T>
int allocPath(RenderState& r, T** out, size_t count) {
    void* d = nullptr;
    CUDA_OK(cudaMalloc(&d, std::max<size_t>(count * sizeof(T), 16)));
    r.path_buffers.push_back(d);
    *out = static_cast<T*>(d);
    return 0;
}

// Path slots per pass: the late bounces of a pass run on what is left of it, so a pass should fill the machine many times over
// (measured, profiles/r02_sweeps.md: 4 Mi instead of 16 Mi slots costs 6 - 28 % of a frame; the 66 M paths of a 4K x 8 spp frame in
// passes of 32 Mi instead of 16 Mi: 1300 -> 1256 ms) while the path state stays a modest share of the 180 GB.
uint32_t targetPathsPerPass() {
    const char* v = getenv("ZYGPU_PATHS_PER_PASS");
    return v ? uint32_t(std::max(1, atoi(v))) : (32u << 20);
}

// `lanes` vertex records per slot (1, or 4 when a material of the scene can split a path); trace items are vertex ids for
// closest-hit rays and shadow records for any-hit rays, so the per-item arrays hold max(lanes, shadow_stride) per slot.
int ensurePaths(RenderState& r, uint32_t capacity, uint32_t shadow_stride, uint32_t lanes, bool deferred_lights, bool textures, bool record_uvs) {
    // the per-sample AOV values are written by the shade_a instances that read image maps (one run-time test there, none in the others)
    const bool aovs = 0 != r.view.aov_slots;
    textures        = textures || aovs;
    if (r.paths.capacity >= capacity && r.paths.shadow_stride == shadow_stride && r.paths.lanes == lanes &&
        (nullptr != r.paths.queue_l) == deferred_lights && (nullptr != r.paths.stoch) == textures &&
        (nullptr != r.paths.sh_uv) == record_uvs && (nullptr != r.paths.aov_misc) == aovs) {
        return 0;
    }
    freeAll(r.path_buffers);
    zygpu::PathState& p = r.paths;
    p                   = zygpu::PathState{};
    const size_t vertices = size_t(capacity) * lanes;
    const size_t items    = size_t(capacity) * std::max(shadow_stride, lanes);
    if (0 != allocPath(r, &p.ray_o, vertices) || 0 != allocPath(r, &p.ray_d, vertices) || 0 != allocPath(r, &p.thr, vertices) ||
        0 != allocPath(r, &p.prev_p, vertices) || 0 != allocPath(r, &p.prev_n, vertices) || 0 != allocPath(r, &p.hit, vertices) ||
        0 != allocPath(r, &p.acc_e, capacity) || 0 != allocPath(r, &p.acc_d, capacity) || 0 != allocPath(r, &p.acc_i, capacity) ||
        0 != allocPath(r, &p.smp, capacity) || 0 != allocPath(r, &p.rng, capacity) ||
        0 != allocPath(r, &p.sh_o, size_t(capacity) * shadow_stride) || 0 != allocPath(r, &p.sh_p, size_t(capacity) * shadow_stride) ||
        0 != allocPath(r, &p.sh_wi, size_t(capacity) * shadow_stride) || 0 != allocPath(r, &p.sh_n, capacity) ||
        0 != allocPath(r, &p.ml_props, items * 8) || 0 != allocPath(r, &p.ml_count, items) ||
        0 != allocPath(r, &p.queue_m, items) || 0 != allocPath(r, &p.sort_bins, zygpu::kSortBins + 1) ||
        0 != allocPath(r, &p.trace_stacks, size_t(zygpu::numSmsOfCurrentDevice()) * 8 * zygpu::kScenePoolStackWords) || 0 != allocPath(r, &p.queue_a, capacity) || 0 != allocPath(r, &p.queue_b, capacity) || 0 != allocPath(r, &p.counters, 16)) {
        return -1;
    }
    if (lanes > 1 && (0 != allocPath(r, &p.med, vertices) || 0 != allocPath(r, &p.queue_t, vertices) || 0 != allocPath(r, &p.queue_s, capacity))) {
        return -1;
    }
    if (shadow_stride > 1 && 0 != allocPath(r, &p.queue_r, size_t(capacity) * shadow_stride)) return -1;
    if (textures && 0 != allocPath(r, &p.stoch, vertices)) return -1;
    if (aovs && (0 != allocPath(r, &p.aov_albedo, capacity) || 0 != allocPath(r, &p.aov_gn, capacity) || 0 != allocPath(r, &p.aov_sn, capacity) ||
                 0 != allocPath(r, &p.aov_misc, capacity))) {
        return -1;
    }
    if (record_uvs && 0 != allocPath(r, &p.sh_uv, size_t(capacity) * shadow_stride)) return -1;
    if (deferred_lights && (0 != allocPath(r, &p.ls_p, capacity) || 0 != allocPath(r, &p.ls_g, capacity) ||
                            0 != allocPath(r, &p.picks, size_t(capacity) * 64) || 0 != allocPath(r, &p.pick_n, capacity) ||
                            0 != allocPath(r, &p.queue_l, capacity))) {
        return -1;
    }
    CUDA_OK(cudaMemsetAsync(p.counters, 0, 16 * sizeof(uint32_t), r.stream));  // ordered before the pass (the stream is non-blocking)
    p.trace_stack_blocks = uint32_t(zygpu::numSmsOfCurrentDevice()) * 8;
    p.capacity      = capacity;
    p.shadow_stride = shadow_stride;
    p.lanes         = lanes;
    return 0;
}

}  // namespace

void zygpuReleaseRender(zygpu_device* dev) {
    RenderState& r = dev->render;
    freeAll(r.scene_buffers);
    freeAll(r.path_buffers);
    cudaFree(r.film);
    cudaFree(r.resolved);
    cudaFree(r.tally);
    cudaFree(r.film_alpha);
    r.film_alpha = nullptr;
    for (float4*& layer : r.aov.layers) {
        cudaFree(layer);
        layer = nullptr;
    }
    r.tally    = nullptr;
    r.counting = false;
    if (r.stream) {
        cudaStreamSynchronize(r.stream);  // host functions of finished passes still point at this state
        cudaStreamDestroy(r.stream);
    }
    r.scene_buffer_bytes.clear();
    r.scene_buffer_cursor = 0;
    r.scene               = zygpu::SceneDevice{};
    r.paths               = zygpu::PathState{};
    r.has_scene = r.has_view = r.has_meshes = false;
    r.film = r.resolved = nullptr;
    r.film_pixels       = 0;
    r.stream            = nullptr;
}

extern "C" {

int zygpu_upload_scene(zygpu_device* dev, const ZygpuScene* scene) {
    if (!dev || !scene) return fail("zygpu_upload_scene: null argument");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    RenderState& r = dev->render;
    if (!r.stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
        CUDA_OK(zygpu::uploadSobolDirections());
    }
    CUDA_OK(cudaStreamSynchronize(r.stream));
    r.scene_buffer_cursor = 0;  // the arrays below take the buffers of the previous upload in the same order
    r.has_scene           = false;

    if (!scene->ggx_luts) return fail("zygpu_upload_scene: scene carries no GGX tables");

    zygpu::SceneDevice& d = r.scene;
    d                     = zygpu::SceneDevice{};

    const float4* f4 = nullptr;
    if (0 != uploadArray(r, scene->props, scene->num_props, &d.props)) return -1;
    if (0 != uploadArray(r, reinterpret_cast<const float4*>(scene->trafos), size_t(scene->num_props) * 4, &f4)) return -1;
    d.trafos = f4;
    if (0 != uploadArray(r, reinterpret_cast<const float4*>(scene->aabbs), size_t(scene->num_props) * 2, &f4)) return -1;
    d.aabbs = f4;
    if (0 != uploadArray(r, scene->material_ids, scene->num_parts, &d.material_ids)) return -1;
    if (0 != uploadArray(r, scene->light_ids, scene->num_parts, &d.light_ids)) return -1;
    if (0 != uploadArray(r, scene->materials, scene->num_materials, &d.materials)) return -1;
    if (0 != uploadArray(r, scene->lights, scene->num_lights, &d.lights)) return -1;
    if (0 != uploadArray(r, reinterpret_cast<const float4*>(scene->light_aabbs), size_t(scene->num_lights) * 2, &f4)) return -1;
    d.light_aabbs = f4;
    if (0 != uploadArray(r, reinterpret_cast<const float4*>(scene->light_cones), scene->num_lights, &f4)) return -1;
    d.light_cones = f4;

    const ZygpuLightTree& lt = scene->light_tree;
    if (0 != uploadArray(r, lt.nodes, lt.num_nodes, &d.lt_nodes)) return -1;
    if (0 != uploadArray(r, lt.node_middles, lt.num_nodes, &d.lt_middles)) return -1;
    if (0 != uploadArray(r, lt.light_orders, lt.num_lights, &d.lt_orders)) return -1;
    if (0 != uploadArray(r, lt.light_mapping, lt.num_lights, &d.lt_mapping)) return -1;
    d.lt_bounds_min      = make_float4(lt.bounds.min[0], lt.bounds.min[1], lt.bounds.min[2], lt.bounds.min[3]);
    d.lt_bounds_max      = make_float4(lt.bounds.max[0], lt.bounds.max[1], lt.bounds.max[2], lt.bounds.max[3]);
    d.lt_infinite_weight = lt.infinite_weight;
    d.lt_infinite_guard  = lt.infinite_guard;
    d.lt_infinite_end    = lt.infinite_end;
    d.lt_max_split_depth = lt.max_split_depth;
    d.lt_num_infinite    = lt.num_infinite_lights;
    d.lt_num_nodes       = lt.num_nodes;
    if (lt.num_infinite_lights > 0 && !lt.infinite_cdf) return fail("zygpu_upload_scene: light tree carries no infinite_cdf");
    if (0 != uploadArray(r, lt.infinite_cdf, lt.num_infinite_lights > 0 ? size_t(lt.num_infinite_lights) + 1 : 0, &d.lt_infinite_cdf)) return -1;

    if (0 != uploadArray(r, reinterpret_cast<const float4*>(scene->solid_bvh.nodes), size_t(scene->solid_bvh.num_nodes) * 2, &f4)) return -1;
    d.solid_nodes = f4;
    if (0 != uploadArray(r, scene->solid_bvh.indices, scene->solid_bvh.num_indices, &d.solid_indices)) return -1;
    d.num_solid_nodes = scene->solid_bvh.num_nodes;
    if (0 != uploadArray(r, reinterpret_cast<const float4*>(scene->unoccluding_bvh.nodes), size_t(scene->unoccluding_bvh.num_nodes) * 2, &f4)) return -1;
    d.unocc_nodes = f4;
    if (0 != uploadArray(r, scene->unoccluding_bvh.indices, scene->unoccluding_bvh.num_indices, &d.unocc_indices)) return -1;
    d.num_unocc_nodes = scene->unoccluding_bvh.num_nodes;

    {  // the grid of the ray sort: 32 x 8 x 32 cells over the box of the finite props (y is up in zyg's scenes: the thin axis)
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t p = 0; p < scene->num_props; ++p) {
            const ZygpuAabb& b = scene->aabbs[p];
            bool finite = true;
            for (int a = 0; a < 3; ++a) finite = finite && std::fabs(b.min[a]) < 1e30f && std::fabs(b.max[a]) < 1e30f && b.min[a] <= b.max[a];
            if (!finite) continue;
            for (int a = 0; a < 3; ++a) {
                lo[a] = std::min(lo[a], b.min[a]);
                hi[a] = std::max(hi[a], b.max[a]);
            }
        }
        const float cells[3] = {32.f, 8.f, 32.f};
        float       per[3];
        for (int a = 0; a < 3; ++a) {
            if (!(lo[a] <= hi[a])) lo[a] = hi[a] = 0.f;
            per[a] = hi[a] > lo[a] ? cells[a] / (hi[a] - lo[a]) : 0.f;
        }
        d.world_lo    = make_float4(lo[0], lo[1], lo[2], 0.f);
        d.world_cells = make_float4(per[0], per[1], per[2], 0.f);
    }

    {  // "flattened on upload": the solid prop tree in the 8-wide quantised layout the fused traversal kernel walks
        // world-space bounding spheres of the mesh props: the mesh's object-space sphere through the prop's transformation
        std::vector<float> spheres(size_t(scene->num_props) * 4, 0.f);
        for (uint32_t p = 0; p < scene->num_props; ++p) {
            float* s = &spheres[size_t(p) * 4];
            s[3]     = FLT_MAX;
            const ZygpuProp& prop = scene->props[p];
            if (ZYG_SHAPE_TRIANGLE_MESH != prop.shape || prop.mesh >= scene->num_meshes) continue;
            const zyg::WideBvh& wb = scene->meshes[prop.mesh]->wide;
            const ZygpuTrafo&   t  = scene->trafos[p];
            double              c[3] = {t.position[0], t.position[1], t.position[2]};
            double              max_scale = 0.0;
            for (int k = 0; k < 3; ++k) {
                const double sk = t.r[k][3];
                max_scale       = std::max(max_scale, std::fabs(sk));
                for (int i = 0; i < 3; ++i) c[i] += double(wb.bound_center[k]) * sk * double(t.r[k][i]);
            }
            const double len = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
            for (int i = 0; i < 3; ++i) s[i] = float(c[i]);
            // slack for the fp32 arithmetic of the device test and of the object-space traversal it stands in for
            s[3] = float(double(wb.bound_radius) * max_scale * (1.0 + 1e-4) + 1e-5 * (len + 1.0));
        }
        zyg::WidePropBvh wide;
        zyg::buildWidePropBvh(scene->solid_bvh.nodes, scene->solid_bvh.num_nodes, scene->solid_bvh.indices, scene->aabbs, spheres.data(), wide);
        if (0 != uploadArray(r, reinterpret_cast<const float4*>(wide.nodes.data()), wide.nodes.size() * (sizeof(zyg::WideNode) / 16), &f4)) return -1;
        d.tlas_nodes = f4;
        if (0 != uploadArray(r, reinterpret_cast<const float4*>(wide.records.data()), wide.records.size() * (sizeof(zyg::PropRecord) / 16), &f4)) return -1;
        d.tlas_recs = f4;
        uint32_t mesh_depth = 0;
        for (uint32_t m = 0; m < scene->num_meshes; ++m) mesh_depth = std::max(mesh_depth, scene->meshes[m]->wide.max_depth);
        d.trace_stack_bound = 2 * (wide.max_depth + mesh_depth) + 6;
        if (getenv("ZYGPU_DEBUG_TLAS")) {
            fprintf(stderr, "[zygpu] prop tree: %zu wide nodes, depth %u; deepest mesh tree %u; stack bound %u\n", wide.nodes.size(), wide.max_depth,
                    mesh_depth, d.trace_stack_bound);
        }
    }

    if (0 != uploadArray(r, scene->infinite_props, scene->num_infinite_props, &d.infinite_props)) return -1;
    d.num_infinite_props = scene->num_infinite_props;

    // meshes: uploaded once per zyg_mesh, referenced through a per-scene table
    std::vector<zygpu::MeshDevice>  mesh_views(scene->num_meshes);
    std::vector<zygpu::MeshShading> mesh_shading(scene->num_meshes);
    for (uint32_t m = 0; m < scene->num_meshes; ++m) {
        const int id = zygpu_upload_mesh(dev, scene->meshes[m]);
        if (id < 0) return -1;
        mesh_views[m]   = dev->meshes[size_t(id)].view;
        mesh_shading[m] = dev->meshes[size_t(id)].shading;
    }
    if (0 != uploadArray(r, mesh_views.data(), mesh_views.size(), &d.meshes)) return -1;
    if (0 != uploadArray(r, mesh_shading.data(), mesh_shading.size(), &d.mesh_shading)) return -1;

    if (0 != uploadArray(r, scene->ggx_luts, size_t(ZYGPU_GGX_LUT_FLOATS), &d.luts)) return -1;

    // mesh-light samplers (shape_sampler.MeshImpl + PrimitiveTree)
    std::vector<zygpu::MeshSamplerDevice> samplers(scene->num_mesh_samplers);
    for (uint32_t i = 0; i < scene->num_mesh_samplers; ++i) {
        const ZygpuMeshSampler&   ms = scene->mesh_samplers[i];
        zygpu::MeshSamplerDevice& sd = samplers[i];
        if (ms.mesh >= scene->num_meshes) return fail("zygpu_upload_scene: mesh sampler %u references mesh %u", i, ms.mesh);
        sd.bounds_min    = make_float4(ms.bounds.min[0], ms.bounds.min[1], ms.bounds.min[2], ms.bounds.min[3]);
        sd.bounds_max    = make_float4(ms.bounds.max[0], ms.bounds.max[1], ms.bounds.max[2], ms.bounds.max[3]);
        sd.num_triangles = ms.num_triangles;
        sd.num_nodes     = ms.num_nodes;
        sd.two_sided     = ms.two_sided;
        sd.mesh          = ms.mesh;
        const uint32_t tree_triangles = uint32_t(scene->meshes[ms.mesh]->tree.numTriangles());
        if (0 != uploadArray(r, ms.nodes, ms.num_nodes, &sd.nodes) || 0 != uploadArray(r, ms.node_middles, ms.num_nodes, &sd.node_middles) ||
            0 != uploadArray(r, ms.light_orders, ms.num_triangles, &sd.light_orders) ||
            0 != uploadArray(r, ms.light_mapping, ms.num_triangles, &sd.light_mapping) ||
            0 != uploadArray(r, ms.triangle_mapping, ms.num_triangles, &sd.triangle_mapping) ||
            0 != uploadArray(r, ms.triangle_pdfs, ms.num_triangles, &sd.triangle_pdfs) ||
            0 != uploadArray(r, ms.primitive_mapping, tree_triangles, &sd.primitive_mapping)) {
            return -1;
        }
        // centre, radius and normal of every light triangle, derived once on the device (device/render.cu meshLightPropsKernel)
        if (0 != uploadArray<float4>(r, nullptr, size_t(ms.num_triangles) * 2, &sd.triangle_props)) return -1;
        CUDA_OK(zygpu::launchMeshLightProps(mesh_views[ms.mesh], sd, const_cast<float4*>(sd.triangle_props), r.stream));
    }
    if (0 != uploadArray(r, samplers.data(), samplers.size(), &d.mesh_samplers)) return -1;
    d.num_mesh_samplers = scene->num_mesh_samplers;
    if (0 != uploadArray(r, scene->mesh_part_areas, scene->mesh_part_areas ? scene->num_parts : 0, &d.mesh_part_areas)) return -1;

    r.has_textures          = false;
    r.has_image_area_lights = false;
    for (uint32_t l = 0; l < scene->num_lights; ++l) {
        const ZygpuLight& light = scene->lights[l];
        // Disk.sampleTo lives in the shade kernel instances that carry the rarer features (device/render.cu shadeAKernel<.., Textured>)
        if (light.prop < scene->num_props && ZYG_SHAPE_DISK == scene->props[light.prop].shape) r.has_textures = true;
        if (ZYG_LIGHT_PROP_IMAGE != light.light_class) continue;
        const uint32_t shape = scene->props[light.prop].shape;
        if (ZYG_SHAPE_CANOPY != shape && ZYG_SHAPE_RECTANGLE != shape) {
            return fail("zygpu_upload_scene: light %u: image-mapped lights are supported on Canopy and Rectangle shapes", l);
        }
        if (light.sampler >= scene->num_image_samplers) return fail("zygpu_upload_scene: light %u references image sampler %u", l, light.sampler);
        if (ZYG_SHAPE_RECTANGLE == shape) r.has_image_area_lights = true;
    }
    // emission images with their Distribution2D rows (shape_sampler.ImageImpl)
    std::vector<zygpu::ImageSamplerDevice> image_samplers(scene->num_image_samplers);
    for (uint32_t i = 0; i < scene->num_image_samplers; ++i) {
        const ZygpuImageSampler&   is = scene->image_samplers[i];
        zygpu::ImageSamplerDevice& id = image_samplers[i];
        if (0 == is.width || 0 == is.height || !is.pixels || (nullptr == is.marginal_cdf) != (nullptr == is.conditional_cdf)) {
            return fail("zygpu_upload_scene: image sampler %u is incomplete", i);
        }
        id.width        = is.width;
        id.height       = is.height;
        id.address_u    = is.address_u;
        id.address_v    = is.address_v;
        id.filter       = is.filter;
        id.total_weight = is.total_weight;
        id.scale_u      = is.scale[0];
        id.scale_v      = is.scale[1];
        // the same image may back several samplers (one per uv-weight class): upload its pixels once
        id.pixels = nullptr;
        for (uint32_t k = 0; k < i; ++k) {
            if (scene->image_samplers[k].pixels == is.pixels) id.pixels = image_samplers[k].pixels;
        }
        if (!id.pixels && 0 != uploadArray(r, is.pixels, size_t(is.width) * is.height * 3, &id.pixels)) return -1;
        id.marginal_cdf = id.conditional_cdf = nullptr;  // an image that is only looked up (a colour map) has no distribution
        if (is.marginal_cdf && (0 != uploadArray(r, is.marginal_cdf, size_t(is.height) + 1, &id.marginal_cdf) ||
                                0 != uploadArray(r, is.conditional_cdf, size_t(is.height) * (is.width + 1), &id.conditional_cdf))) {
            return -1;
        }
    }
    if (0 != uploadArray(r, image_samplers.data(), image_samplers.size(), &d.image_samplers)) return -1;
    for (uint32_t m = 0; m < scene->num_materials; ++m) {
        if (ZYGPU_NULL != scene->materials[m].emission_map && scene->materials[m].emission_map >= scene->num_image_samplers) {
            return fail("zygpu_upload_scene: material %u references image sampler %u", m, scene->materials[m].emission_map);
        }
        if (ZYGPU_NULL != scene->materials[m].color_map && scene->materials[m].color_map >= scene->num_image_samplers) {
            return fail("zygpu_upload_scene: material %u references image sampler %u", m, scene->materials[m].color_map);
        }
        const ZygpuMaterial& mat = scene->materials[m];
        if (ZYGPU_NULL != mat.color_map || ZYGPU_NULL != mat.roughness_map || ZYGPU_NULL != mat.metallic_map || ZYGPU_NULL != mat.normal_map) {
            r.has_textures = true;
        }
        // a coated Substitute runs the same shade kernel instances (device/shading.cuh materialSample<Glass, Coated>)
        if (ZYG_MATERIAL_SUBSTITUTE == mat.type && mat.coating_thickness > 0.f) r.has_textures = true;
    }

    // shadow records one path vertex can need: every light the tree may return times its sample count
    // (Tree.potentialMaxLights, light_tree.zig:331-344), capped like the reference's buffers (64 picks x 64 samples)
    uint64_t potential = 0, worst = 0;
    uint32_t most = 1, worst_most = 1;
    for (uint32_t l = 0; l < scene->num_lights; ++l) {
        // Light.potentialMaxSamples, light.zig:77-85: a mesh light can return a sample per leaf its primitive tree splits into
        // (2^6 = Shape.MaxSamples). The first reservation assumes 8; a pass in which a vertex needed more is run again with a
        // larger one (zygpu_render), so no sample is ever dropped.
        const bool     mesh = ZYGPU_NULL != scene->lights[l].sampler && ZYG_LIGHT_PROP_IMAGE != scene->lights[l].light_class;
        const uint32_t n    = mesh ? 8u : std::max(1u, scene->lights[l].num_samples);
        const uint32_t w    = mesh ? 64u : n;
        potential += n;
        worst += w;
        most       = std::max(most, n);
        worst_most = std::max(worst_most, w);
    }
    // the tree returns at most Tree.MaxLights = 64 picks (light_tree.zig:249), each with up to `most` samples
    r.max_light_samples   = uint32_t(std::min<uint64_t>(std::max<uint64_t>(potential, 1), 64ull * std::min(most, 64u)));
    r.worst_light_samples = uint32_t(std::min<uint64_t>(std::max<uint64_t>(worst, 1), 64ull * std::min(worst_most, 64u)));
    {  // records are reserved per path slot: keep the first reservation moderate (ZYGPU_MAX_LIGHT_SAMPLES overrides)
        const char*    v   = getenv("ZYGPU_MAX_LIGHT_SAMPLES");
        const uint32_t cap = v ? uint32_t(std::max(1, atoi(v))) : 128u;
        r.max_light_samples = std::min(r.max_light_samples, cap);
    }

    // Glass splits a path into its reflected and refracted branch (glass_sample.zig:256-269, 350-395): such scenes run with
    // Pool.NumVertices vertex records per camera sample and one shade round per record
    r.can_split    = false;
    for (uint32_t m = 0; m < scene->num_materials; ++m) {
        const ZygpuMaterial& mat = scene->materials[m];
        if (ZYG_MATERIAL_GLASS == mat.type) {
            if (mat.thickness > 0.f || 0.f != mat.abbe) return fail("zygpu_upload_scene: thin (thickness > 0) and dispersive (abbe != 0) Glass are not supported");
            r.can_split = true;
        }
        if (mat.priority < -127 || mat.priority > 127) return fail("zygpu_upload_scene: material priority outside i8");
    }
    // with many lights the light selection / sampling of a vertex varies from one tree descent to dozens of picks: such
    // scenes run it in the persistent light kernels instead of inside shade_a (ZYGPU_DEFERRED_LIGHTS=0/1 overrides)
    r.deferred_lights = scene->num_lights >= 8;
    if (const char* v = getenv("ZYGPU_DEFERRED_LIGHTS")) r.deferred_lights = 0 != atoi(v);
    if (scene->num_mesh_samplers > 0) r.deferred_lights = true;  // triangle-mesh lights are only sampled by the light kernels

    r.has_meshes = scene->num_meshes > 0;
    r.has_scene  = true;
    return 0;
}

int zygpu_set_view(zygpu_device* dev, const ZygpuView* view) {
    if (!dev || !view) return fail("zygpu_set_view: null argument");
    if (view->resolution[0] <= 0 || view->resolution[1] <= 0) return fail("zygpu_set_view: empty resolution");
    if (view->filter_radius_int < 0 || view->filter_radius_int > 2) return fail("zygpu_set_view: filter radius must be 0, 1 or 2");
    if (view->max_depth_surface > 255) return fail("zygpu_set_view: max surface depth above 255");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    RenderState& r = dev->render;
    if (!r.stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking));
        CUDA_OK(zygpu::uploadSobolDirections());
    }
    CUDA_OK(cudaStreamSynchronize(r.stream));

    const uint32_t pixels = uint32_t(view->resolution[0]) * uint32_t(view->resolution[1]);
    if (pixels != r.film_pixels) {  // Sensor.resize, sensor.zig:128-150
        cudaFree(r.film);
        cudaFree(r.resolved);
        r.film = r.resolved = nullptr;
        CUDA_OK(cudaMalloc(&r.film, size_t(pixels) * sizeof(float4)));
        CUDA_OK(cudaMalloc(&r.resolved, size_t(pixels) * sizeof(float4)));
        CUDA_OK(cudaMemset(r.film, 0, size_t(pixels) * sizeof(float4)));
        for (float4*& layer : r.aov.layers) {
            cudaFree(layer);
            layer = nullptr;
        }
        cudaFree(r.film_alpha);
        r.film_alpha  = nullptr;
        r.film_pixels = pixels;
    }
    // Sensor.Buffer.Class: Opaque or Transparent (buffer.zig:9-23)
    if (0 != view->alpha_transparency && !r.film_alpha) {
        CUDA_OK(cudaMalloc(&r.film_alpha, size_t(pixels) * sizeof(float)));
        CUDA_OK(cudaMemsetAsync(r.film_alpha, 0, size_t(pixels) * sizeof(float), r.stream));
    } else if (0 == view->alpha_transparency && r.film_alpha) {
        cudaFree(r.film_alpha);
        r.film_alpha = nullptr;
    }
    // aov.Buffer.resize, aov_buffer.zig:28-37: a layer per active class
    bool new_layers = false;
    for (uint32_t c = 0; c < ZYG_AOV_NUM_CLASSES; ++c) {
        const bool active = 0 != (view->aov_slots & (1u << c));
        if (active && !r.aov.layers[c]) {
            CUDA_OK(cudaMalloc(&r.aov.layers[c], size_t(pixels) * sizeof(float4)));
            new_layers = true;
        } else if (!active && r.aov.layers[c]) {
            cudaFree(r.aov.layers[c]);
            r.aov.layers[c] = nullptr;
        }
    }
    if (new_layers) CUDA_OK(zygpu::launchAovClear(r.aov, pixels, r.stream));
    r.view     = *view;
    r.has_view = true;
    return 0;
}

int zygpu_clear_film(zygpu_device* dev) {
    if (!dev || !dev->render.film) return fail("zygpu_clear_film: no view set");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    RenderState& r = dev->render;
    CUDA_OK(cudaMemsetAsync(r.film, 0, size_t(r.film_pixels) * sizeof(float4), r.stream));
    if (0 != r.view.aov_slots) CUDA_OK(zygpu::launchAovClear(r.aov, r.film_pixels, r.stream));
    if (r.film_alpha) CUDA_OK(cudaMemsetAsync(r.film_alpha, 0, size_t(r.film_pixels) * sizeof(float), r.stream));
    if (r.paths.counters) CUDA_OK(cudaMemsetAsync(r.paths.counters, 0, 16 * sizeof(uint32_t), r.stream));
    r.stats = ZygpuRenderStats{};
    r.stats_carry[0] = r.stats_carry[1] = 0;
    return 0;
}

int zygpu_render(zygpu_device* dev, uint32_t iteration, uint32_t num_samples) {
    if (!dev) return fail("zygpu_render: null device");
    RenderState& r = dev->render;
    if (!r.has_scene || !r.has_view) return fail("zygpu_render: upload a scene and set a view first");
    CUDA_OK(cudaSetDevice(dev->ordinal));

    const ZygpuView& view = r.view;
    const uint32_t   fr   = uint32_t(view.filter_radius_int);
    const uint32_t   pw   = uint32_t(view.resolution[0]) + 2 * fr;
    const uint32_t   ph   = uint32_t(view.resolution[1]) + 2 * fr;
    const uint64_t   padded = uint64_t(pw) * ph;
    if (padded > 0xFFFFFFFFull) return fail("zygpu_render: resolution too large");

    const uint32_t lanes  = r.can_split ? 4 : 1;
    const uint32_t rounds = lanes;
    // extend / shadow are one kernel each (prop-tree walk) plus the persistent mesh kernel when the scene has meshes
    const uint32_t trace_extra = zygpu::sceneTraceLaunches(r.has_meshes, r.scene.num_solid_nodes) - 1;

    // Samples per pass for the current shadow-record reservation: a pass holds at most 64 Mi shadow records (3 GiB), so scenes
    // whose vertices can sample many lights trace fewer paths per pass. 0 = a single frame of paths does not fit.
    auto samplesPerPass = [&]() -> uint32_t {
        const uint64_t target   = std::min<uint64_t>(targetPathsPerPass(), std::max<uint64_t>(padded, (64ull << 20) / r.max_light_samples));
        const uint32_t per_pass = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(num_samples, target / padded)));
        const uint64_t capacity = padded * per_pass;
        if (capacity * lanes > 0xFFFFFFFFull || capacity * r.max_light_samples > 0xFFFFFFFFull) return 0;  // 32-bit vertex / record ids
        return per_pass;
    };
    uint32_t per_pass = samplesPerPass();
    if (0 == per_pass) return fail("zygpu_render: %llu paths x %u shadow records per pass exceed the 32-bit record ids", (unsigned long long)padded, r.max_light_samples);
    if (0 != ensurePaths(r, uint32_t(padded * per_pass), r.max_light_samples, lanes, r.deferred_lights, r.has_textures, r.has_image_area_lights)) return -1;

    for (uint32_t done = 0; done < num_samples;) {
        const uint32_t k = std::min(per_pass, num_samples - done);

        zygpu::PassParams pass;
        pass.iteration       = iteration + done;
        pass.samples_in_pass = k;
        pass.padded_w        = pw;
        pass.padded_h        = ph;
        pass.num_paths       = uint32_t(padded) * k;
        pass.debug_slot      = getenv("ZYGPU_DEBUG_SLOT") ? uint32_t(strtoul(getenv("ZYGPU_DEBUG_SLOT"), nullptr, 10)) : 0xFFFFFFFFu;

        r.paths.tally = r.counting ? r.tally : nullptr;
        CUDA_OK(zygpu::launchGenerate(view, r.paths, pass, r.stream));
        r.stats.kernel_launches += 1;

        // one bounce = extend, shade_a, shadow, shade_b (+ queue swap); depth max_depth_surface is the last vertex
        // that can be reached (pathtracer_mis.zig:76-86), so max_depth + 1 extend / shade_a rounds
        for (uint32_t bounce = 0; bounce <= view.max_depth_surface; ++bounce) {
            const bool last = bounce == view.max_depth_surface;
            // all vertices of the generation are extended at once (no sampler draws in between) ...
            static const bool verify_trace = nullptr != getenv("ZYGPU_VERIFY_TRACE");
            if (verify_trace) {  // diagnostics: the fused walk against one thread per ray in the reference's order
                const size_t bytes = size_t(r.paths.capacity) * lanes * sizeof(float4);
                float4 *before = nullptr, *ta = nullptr, *ha = nullptr;
                CUDA_OK(cudaMalloc(&before, bytes));
                CUDA_OK(cudaMalloc(&ta, bytes));
                CUDA_OK(cudaMalloc(&ha, bytes));
                CUDA_OK(cudaMemcpyAsync(before, r.paths.ray_d, bytes, cudaMemcpyDeviceToDevice, r.stream));
                CUDA_OK(zygpu::launchExtend(r.scene, r.paths, pass.num_paths * lanes, r.has_meshes, bounce, r.stream));
                CUDA_OK(cudaMemcpyAsync(ta, r.paths.ray_d, bytes, cudaMemcpyDeviceToDevice, r.stream));
                CUDA_OK(cudaMemcpyAsync(ha, r.paths.hit, bytes, cudaMemcpyDeviceToDevice, r.stream));
                CUDA_OK(cudaMemcpyAsync(r.paths.ray_d, before, bytes, cudaMemcpyDeviceToDevice, r.stream));
                CUDA_OK(zygpu::launchExtendReference(r.scene, r.paths, pass.num_paths * lanes, r.stream));
                CUDA_OK(zygpu::launchCompareHits(r.paths, before, ta, ha, pass.num_paths * lanes, bounce, r.stream));
                CUDA_OK(cudaStreamSynchronize(r.stream));
                cudaFree(before);
                cudaFree(ta);
                cudaFree(ha);
            } else
            CUDA_OK(zygpu::launchExtend(r.scene, r.paths, pass.num_paths * lanes, r.has_meshes, bounce, r.stream));
            r.stats.kernel_launches += 1 + trace_extra;
            if (lanes > 1) {
                CUDA_OK(zygpu::launchBeginGeneration(r.paths, r.stream));
                r.stats.kernel_launches += 1;
            }
            // ... and shaded in rounds: round k handles the k-th vertex of every camera sample, because the vertices of a
            // sample share its sampler and draw from it in turn (VertexPool.consume order)
            for (uint32_t round = 0; round < rounds; ++round) {
                if (round > 0) {
                    CUDA_OK(zygpu::launchBeginRound(r.paths, r.stream));
                    r.stats.kernel_launches += 1;
                }
                CUDA_OK(zygpu::launchShadeA(r.scene, view, r.paths, pass, pass.num_paths, round, r.stream));
                r.stats.kernel_launches += 1;
                if (last) continue;
                if (r.deferred_lights) {
                    CUDA_OK(zygpu::launchLightStages(r.scene, view, r.paths, pass, pass.num_paths, r.stream));
                    r.stats.kernel_launches += 2;
                }
                CUDA_OK(zygpu::launchShadow(r.scene, r.paths, pass.num_paths, r.has_meshes, bounce, r.stream));
                CUDA_OK(zygpu::launchShadeB(r.scene, view, r.paths, pass, pass.num_paths, round, r.stream));
                r.stats.kernel_launches += 2 + trace_extra + (1 == lanes ? 1 : 0);  // shadow, shade_b (+ queue swap)
            }
            if (lanes > 1) {
                CUDA_OK(zygpu::launchEndGeneration(r.paths, r.stream));
                r.stats.kernel_launches += 1;
            }
        }

        // A vertex that produced more light samples than a slot reserves sets counters[3] and drops the extra records. Where the
        // reservation is below what the light tree can return (mesh lights, the 128-record cap) the flag is read before the pass
        // reaches the film: an overflowed pass is discarded and run again with twice the reservation, so the film never sees a
        // biased sample. Scenes whose reservation is exact skip the read and stay fully asynchronous.
        if (r.max_light_samples < r.worst_light_samples) {
            uint32_t c[8];
            CUDA_OK(cudaMemcpyAsync(c, r.paths.counters, sizeof(c), cudaMemcpyDeviceToHost, r.stream));
            CUDA_OK(cudaStreamSynchronize(r.stream));
            if (0 != c[3]) {
                r.stats_carry[0] += c[5];  // the ray counters restart with the new buffers
                r.stats_carry[1] += c[6];
                r.max_light_samples = std::min(r.worst_light_samples, r.max_light_samples * 2);
                per_pass            = samplesPerPass();
                if (0 == per_pass) return fail("zygpu_render: %llu paths x %u shadow records per pass exceed the 32-bit record ids", (unsigned long long)padded, r.max_light_samples);
                if (0 != ensurePaths(r, uint32_t(padded * per_pass), r.max_light_samples, lanes, r.deferred_lights, r.has_textures, r.has_image_area_lights)) return -1;
                r.stats.overflow_retries += 1;
                continue;  // same `done`: the pass is rendered again
            }
        }

        CUDA_OK(zygpu::launchFilm(view, r.paths, pass, r.film, r.film_alpha, r.stream));
        if (0 != view.aov_slots) {
            CUDA_OK(zygpu::launchAovFilm(view, r.paths, pass, r.aov, r.stream));
            r.stats.kernel_launches += 1;
        }
        r.stats.kernel_launches += 1;

        r.stats.camera_samples += uint64_t(view.crop[2] - view.crop[0] + 2 * fr) * uint64_t(view.crop[3] - view.crop[1] + 2 * fr) * k;
        r.stats.passes += 1;
        r.passes_enqueued += 1;
        CUDA_OK(cudaLaunchHostFunc(r.stream, [](void* counter) { static_cast<std::atomic<uint32_t>*>(counter)->fetch_add(1); }, &r.passes_completed));
        done += k;
    }
    return 0;
}

// Instrumented traversal: counts the 80-byte node, 64-byte triangle-record and 32-byte prop-record fetches of
// the render path's closest-hit and shadow traversal. The counting kernels are slower; timing runs leave it off.
int zygpu_set_counting(zygpu_device* dev, int on) {
    if (!dev) return fail("zygpu_set_counting: null device");
    RenderState& r = dev->render;
    CUDA_OK(cudaSetDevice(dev->ordinal));
    if (on && !r.tally) {
        CUDA_OK(cudaMalloc(&r.tally, 12 * sizeof(unsigned long long)));
    }
    if (r.tally) CUDA_OK(cudaMemset(r.tally, 0, 12 * sizeof(unsigned long long)));
    r.counting = 0 != on;
    return 0;
}

int zygpu_traversal_counts(zygpu_device* dev, ZygpuTraversalCounts* closest, ZygpuTraversalCounts* shadow) {
    if (!dev || !closest || !shadow) return fail("zygpu_traversal_counts: null argument");
    RenderState& r = dev->render;
    if (!r.tally) return fail("zygpu_traversal_counts: counting was never enabled");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    if (r.stream) CUDA_OK(cudaStreamSynchronize(r.stream));
    unsigned long long c[12];
    CUDA_OK(cudaMemcpy(c, r.tally, sizeof(c), cudaMemcpyDeviceToHost));
    ZygpuTraversalCounts* out[2] = {closest, shadow};
    for (int k = 0; k < 2; ++k) {
        out[k]->nodes = c[6 * k], out[k]->triangles = c[6 * k + 1], out[k]->props = c[6 * k + 2];
        out[k]->node_steps = c[6 * k + 3], out[k]->triangle_steps = c[6 * k + 4], out[k]->prop_steps = c[6 * k + 5];
    }
    return 0;
}

uint32_t zygpu_passes_enqueued(zygpu_device* dev) { return dev ? dev->render.passes_enqueued : 0; }
uint32_t zygpu_passes_completed(zygpu_device* dev) { return dev ? dev->render.passes_completed.load() : 0; }

int zygpu_device_ordinal(zygpu_device* dev) { return dev ? dev->ordinal : -1; }

// One ncclReduce(sum, fp32) of the weighted-sum film to `root`, enqueued on the render stream behind the passes. NCCL is not a
// link-time dependency: the symbols are taken from the NCCL the process already loaded (the one that made `nccl_comm`).
int zygpu_reduce_film(zygpu_device* dev, void* nccl_comm, int root) {
    if (!dev || !nccl_comm) return fail("zygpu_reduce_film: null argument");
    RenderState& r = dev->render;
    if (!r.film) return fail("zygpu_reduce_film: no view set");
    using ReduceFn = int (*)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);
    using ErrorFn  = const char* (*)(int);
    static ReduceFn reduce = nullptr;
    static ErrorFn  error  = nullptr;
    if (!reduce) {
        void* sym = dlsym(RTLD_DEFAULT, "ncclReduce");
        if (!sym) {
            if (void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL)) sym = dlsym(h, "ncclReduce");
        }
        if (!sym) return fail("zygpu_reduce_film: no NCCL in this process (ncclReduce not found)");
        reduce = reinterpret_cast<ReduceFn>(sym);
        error  = reinterpret_cast<ErrorFn>(dlsym(RTLD_DEFAULT, "ncclGetErrorString"));
    }
    CUDA_OK(cudaSetDevice(dev->ordinal));
    constexpr int kNcclFloat32 = 7, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t, nccl.h
    const int rc = reduce(r.film, r.film, size_t(r.film_pixels) * 4, kNcclFloat32, kNcclSum, root, nccl_comm, r.stream);
    if (0 != rc) return fail("zygpu_reduce_film: ncclReduce failed: %s", error ? error(rc) : "?");
    if (r.film_alpha) {  // the Transparent buffer's alpha lane is a weighted sum like the colour
        const int ra = reduce(r.film_alpha, r.film_alpha, size_t(r.film_pixels), kNcclFloat32, kNcclSum, root, nccl_comm, r.stream);
        if (0 != ra) return fail("zygpu_reduce_film: ncclReduce (alpha) failed: %s", error ? error(ra) : "?");
    }
    return 0;
}

int zygpu_synchronize(zygpu_device* dev) {
    if (!dev) return fail("zygpu_synchronize: null device");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    if (dev->render.stream) CUDA_OK(cudaStreamSynchronize(dev->render.stream));
    if (dev->render.paths.counters) {  // cannot happen any more (zygpu_render re-runs such a pass); kept as a tripwire
        uint32_t overflow = 0;
        CUDA_OK(cudaMemcpy(&overflow, dev->render.paths.counters + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (0 != overflow) return fail("zygpu_render: a path vertex produced more light samples than the %u reserved", dev->render.max_light_samples);
    }
    return 0;
}

int zygpu_resolve(zygpu_device* dev, float* rgba, uint32_t num_pixels) {
    if (!dev || !rgba) return fail("zygpu_resolve: null argument");
    RenderState& r = dev->render;
    if (!r.film) return fail("zygpu_resolve: no view set");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    const uint32_t n = std::min(num_pixels, r.film_pixels);
    CUDA_OK(zygpu::launchResolve(r.view, r.film, r.film_alpha, r.resolved, n, r.stream));
    r.stats.kernel_launches += 1;
    CUDA_OK(cudaMemcpyAsync(rgba, r.resolved, size_t(n) * sizeof(float4), cudaMemcpyDeviceToHost, r.stream));
    CUDA_OK(cudaStreamSynchronize(r.stream));
    return 0;
}

int zygpu_resolve_aov(zygpu_device* dev, uint32_t aov_class, float* rgba, uint32_t num_pixels, int download_layer) {
    if (!dev || !rgba) return fail("zygpu_resolve_aov: null argument");
    RenderState& r = dev->render;
    if (!r.film) return fail("zygpu_resolve_aov: no view set");
    if (aov_class >= ZYG_AOV_NUM_CLASSES || !r.aov.layers[aov_class]) return -2;  // the class is not active
    CUDA_OK(cudaSetDevice(dev->ordinal));
    const uint32_t n = std::min(num_pixels, r.film_pixels);
    const float4*  src = r.aov.layers[aov_class];
    if (0 == download_layer) {
        CUDA_OK(zygpu::launchResolveAov(aov_class, r.aov.layers[aov_class], r.resolved, n, r.stream));
        r.stats.kernel_launches += 1;
        src = r.resolved;
    }
    CUDA_OK(cudaMemcpyAsync(rgba, src, size_t(n) * sizeof(float4), cudaMemcpyDeviceToHost, r.stream));
    CUDA_OK(cudaStreamSynchronize(r.stream));
    return 0;
}

int zygpu_denoise(zygpu_device* dev, float sigma, float* rgba, uint32_t num_pixels) {
    if (!dev || !rgba) return fail("zygpu_denoise: null argument");
    RenderState& r = dev->render;
    if (!r.film) return fail("zygpu_denoise: no view set");
    if (!(sigma > 0.f) || sigma > 16.f) return fail("zygpu_denoise: sigma must be in (0, 16]");
    if (!r.aov.layers[ZYG_AOV_SHADING_NORMAL] || !r.aov.layers[ZYG_AOV_ALBEDO]) return -2;
    CUDA_OK(cudaSetDevice(dev->ordinal));
    // Denoise.init, denoise.zig:34-72
    const int32_t      radius = int32_t(std::ceil(3.f * sigma));
    std::vector<float> weights;
    float              sum    = 0.f;
    const float        sigma2 = sigma * sigma;
    for (int32_t y = -radius; y <= radius; ++y) {
        for (int32_t x = -radius; x <= radius; ++x) {
            const float p = (float(x) * float(x) + float(y) * float(y)) / (2.f * sigma2);
            weights.push_back(std::exp(-p));
            sum += weights.back();
        }
    }
    for (float& g : weights) g /= sum;
    float* d_weights = nullptr;
    CUDA_OK(cudaMalloc(&d_weights, weights.size() * sizeof(float)));
    cudaError_t e = cudaMemcpyAsync(d_weights, weights.data(), weights.size() * sizeof(float), cudaMemcpyHostToDevice, r.stream);
    if (cudaSuccess == e) {
        e = zygpu::launchDenoise(r.view, r.film, r.aov.layers[ZYG_AOV_SHADING_NORMAL], r.aov.layers[ZYG_AOV_ALBEDO], d_weights, radius, r.resolved, r.stream);
    }
    r.stats.kernel_launches += 1;
    const uint32_t n = std::min(num_pixels, r.film_pixels);
    if (cudaSuccess == e) e = cudaMemcpyAsync(rgba, r.resolved, size_t(n) * sizeof(float4), cudaMemcpyDeviceToHost, r.stream);
    if (cudaSuccess == e) e = cudaStreamSynchronize(r.stream);
    cudaFree(d_weights);
    CUDA_OK(e);
    return 0;
}

int zygpu_download_film(zygpu_device* dev, float* film, uint32_t num_pixels) {
    if (!dev || !film) return fail("zygpu_download_film: null argument");
    RenderState& r = dev->render;
    if (!r.film) return fail("zygpu_download_film: no view set");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    const uint32_t n = std::min(num_pixels, r.film_pixels);
    CUDA_OK(cudaMemcpyAsync(film, r.film, size_t(n) * sizeof(float4), cudaMemcpyDeviceToHost, r.stream));
    CUDA_OK(cudaStreamSynchronize(r.stream));
    return 0;
}

int zygpu_upload_film(zygpu_device* dev, const float* film, uint32_t num_pixels) {
    if (!dev || !film) return fail("zygpu_upload_film: null argument");
    RenderState& r = dev->render;
    if (!r.film) return fail("zygpu_upload_film: no view set");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    const uint32_t n = std::min(num_pixels, r.film_pixels);
    CUDA_OK(cudaMemcpyAsync(r.film, film, size_t(n) * sizeof(float4), cudaMemcpyHostToDevice, r.stream));
    CUDA_OK(cudaStreamSynchronize(r.stream));
    return 0;
}

void* zygpu_film_device(zygpu_device* dev, uint64_t* num_floats) {
    if (!dev || !dev->render.film) return nullptr;
    if (num_floats) *num_floats = uint64_t(dev->render.film_pixels) * 4;
    return dev->render.film;
}

void* zygpu_render_stream(zygpu_device* dev) { return dev ? dev->render.stream : nullptr; }

int zygpu_render_stats(zygpu_device* dev, ZygpuRenderStats* stats) {
    if (!dev || !stats) return fail("zygpu_render_stats: null argument");
    RenderState& r = dev->render;
    CUDA_OK(cudaSetDevice(dev->ordinal));
    if (r.stream) CUDA_OK(cudaStreamSynchronize(r.stream));
    *stats = r.stats;
    if (r.paths.counters) {
        uint32_t c[8];
        CUDA_OK(cudaMemcpy(c, r.paths.counters, sizeof(c), cudaMemcpyDeviceToHost));
        stats->closest_rays = r.stats_carry[0] + c[5];
        stats->shadow_rays  = r.stats_carry[1] + c[6];
    }
    return 0;
}

}  // extern "C"
