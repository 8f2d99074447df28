// Device BVH build and refit behind the C ABI (include/zygpu.h). The result is an ordinary zyg_mesh: the arrays the
// device made are copied back into the handle, so Scene.compile, the mesh-light sampler build and zygpu_upload_mesh treat it like a
// host-built mesh.
#include "../../../include/zygpu.h"

#include "device_state.hpp"
#include "../device/build.cuh"
#include "../device/light_build.cuh"
#include "../device/trace_device.cuh"
#include "../host/light_tree_builder.hpp"
#include "../host/mesh_handle.hpp"

#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

namespace {

struct DeviceBuffers {
    std::vector<void*> owned;
    ~DeviceBuffers() {
        for (void* p : owned) cudaFree(p);
    }
    template <typename T>
    cudaError_t upload(T*& dst, const void* src, size_t bytes, cudaStream_t stream) {
        void*             raw = nullptr;
        const cudaError_t e   = cudaMalloc(&raw, std::max<size_t>(bytes, 16));
        if (cudaSuccess != e) return e;
        owned.push_back(raw);
        dst = static_cast<T*>(raw);
        return cudaMemcpyAsync(raw, src, bytes, cudaMemcpyHostToDevice, stream);
    }
};

// first wide node of every level + the total: both builders number the nodes breadth first, a level after the other
std::vector<uint32_t> wideLevelOffsets(const std::vector<zyg::WideNode>& nodes) {
    std::vector<uint32_t> offsets{0u};
    uint32_t              begin = 0, end = 1;
    while (begin < end) {
        offsets.push_back(end);
        uint32_t next = 0;
        for (uint32_t n = begin; n < end; ++n) next += uint32_t(__builtin_popcount(nodes[n].imask));
        begin = end;
        end += next;
    }
    return offsets;  // offsets.size() - 1 levels
}

uint32_t binaryDepth(const std::vector<zyg::BvhNode>& nodes) {
    struct Item {
        uint32_t node, depth;
    };
    std::vector<Item> stack{{0u, 1u}};
    uint32_t          deepest = 0;
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        deepest = std::max(deepest, it.depth);
        if (0 == nodes[it.node].numIndices()) {
            stack.push_back({nodes[it.node].children(), it.depth + 1});
            stack.push_back({nodes[it.node].children() + 1, it.depth + 1});
        }
    }
    return deepest;
}

// ---- light trees ---------------------------------------------------------------------------------------------------

zygpu_device* g_light_device = nullptr;  // the device the host builder's hook builds on (zygpu_set_light_tree_builder)
float         g_light_build_ms = 0.f;

bool deviceLightTree(const zyg::LightSet& set, const uint32_t* lights, uint32_t num, uint32_t first_order, zyg::LightTreeResult& out,
                     std::vector<uint32_t>& order) {
    zygpu_device* dev = g_light_device;
    if (!dev || num < 2) return false;
    if (cudaSuccess != cudaSetDevice(dev->ordinal)) return false;
    cudaStream_t stream = dev->streams[0];

    // what the builder reads of a light, compacted to the lights of this tree
    std::vector<float4>  lo(num), hi(num), cones(num);
    std::vector<float>   powers(num);
    std::vector<uint8_t> two_sided(num);
    for (uint32_t i = 0; i < num; ++i) {
        const uint32_t   l = lights[i];
        const zyg::AABB& b = set.aabbs[l];
        lo[i]              = make_float4(b.b[0][0], b.b[0][1], b.b[0][2], 0.f);
        hi[i]              = make_float4(b.b[1][0], b.b[1][1], b.b[1][2], 0.f);
        cones[i]           = make_float4(set.cones[l][0], set.cones[l][1], set.cones[l][2], set.cones[l][3]);
        powers[i]          = set.powers[l];
        two_sided[i]       = set.twoSided(l) ? 1 : 0;
    }
    DeviceBuffers          buffers;
    zygpu::LightBuildInput in{};
    float4 *               d_lo = nullptr, *d_hi = nullptr, *d_cones = nullptr;
    float*                 d_powers = nullptr;
    uint8_t*               d_two    = nullptr;
    if (cudaSuccess != buffers.upload(d_lo, lo.data(), lo.size() * sizeof(float4), stream) ||
        cudaSuccess != buffers.upload(d_hi, hi.data(), hi.size() * sizeof(float4), stream) ||
        cudaSuccess != buffers.upload(d_cones, cones.data(), cones.size() * sizeof(float4), stream) ||
        cudaSuccess != buffers.upload(d_powers, powers.data(), powers.size() * sizeof(float), stream) ||
        cudaSuccess != buffers.upload(d_two, two_sided.data(), two_sided.size(), stream)) {
        return false;
    }
    in.aabb_min      = d_lo;
    in.aabb_max      = d_hi;
    in.cones         = d_cones;
    in.powers        = d_powers;
    in.two_sided     = d_two;
    in.num_lights    = num;
    in.all_two_sided = set.all_two_sided;
    in.primitive     = set.primitive;
    in.first_order   = first_order;

    zygpu::LightBuildOutput built;
    if (cudaSuccess != zygpu::buildLightTreeOnDevice(in, built, stream)) {
        zygpu::freeLightBuildOutput(built);
        cudaGetLastError();
        return false;
    }
    std::vector<ZygpuLightNode> slots(built.num_nodes);
    std::vector<uint32_t>       slot_middles(built.num_nodes);
    order.resize(num);
    const bool copied =
        cudaSuccess == cudaMemcpyAsync(slots.data(), built.nodes, slots.size() * sizeof(ZygpuLightNode), cudaMemcpyDeviceToHost, stream) &&
        cudaSuccess == cudaMemcpyAsync(slot_middles.data(), built.node_middles, slot_middles.size() * 4, cudaMemcpyDeviceToHost, stream) &&
        cudaSuccess == cudaMemcpyAsync(order.data(), built.order, order.size() * 4, cudaMemcpyDeviceToHost, stream) &&
        cudaSuccess == cudaStreamSynchronize(stream);
    g_light_build_ms += built.device_ms;
    for (int a = 0; a < 3; ++a) {
        out.bounds.b[0][a] = built.bounds_min[a];
        out.bounds.b[1][a] = built.bounds_max[a];
    }
    out.bounds.b[0][3] = out.bounds.b[1][3] = 0.f;
    out.root_power     = built.root_power;
    zygpu::freeLightBuildOutput(built);
    if (!copied) return false;

    // the device leaves the slots of subtrees that became leaves unused: compact, keeping children adjacent (breadth first)
    out.nodes.clear();
    out.node_middles.clear();
    out.nodes.reserve(slots.size());
    std::vector<uint32_t> queue{0u};
    out.nodes.push_back(slots[0]);
    out.node_middles.push_back(slot_middles[0]);
    for (size_t q = 0; q < queue.size(); ++q) {
        const uint32_t       slot = queue[q];
        const ZygpuLightNode n    = slots[slot];
        if (0 == (n.meta & 1u)) continue;
        const uint32_t c0    = n.meta >> 2;
        const uint32_t new_c = uint32_t(out.nodes.size());
        out.nodes[q].meta    = (n.meta & 3u) | (new_c << 2);
        for (uint32_t k = 0; k < 2; ++k) {
            out.nodes.push_back(slots[c0 + k]);
            out.node_middles.push_back(slot_middles[c0 + k]);
            queue.push_back(c0 + k);
        }
    }
    return true;
}

}  // namespace

extern "C" {

int zygpu_set_light_tree_builder(zygpu_device* dev, uint32_t min_lights) {
    g_light_device = dev;
    zyg::setDeviceLightTreeBuilder(dev ? deviceLightTree : nullptr, min_lights);
    return 0;
}

float zygpu_light_tree_build_ms(int reset) {
    const float ms = g_light_build_ms;
    if (reset) g_light_build_ms = 0.f;
    return ms;
}

int zygpu_mesh_build(zygpu_device* dev, uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles, const uint32_t* indices,
                     uint32_t num_vertices, const float* positions, uint32_t positions_stride, const float* normals,
                     uint32_t normals_stride, const float* uvs, uint32_t uvs_stride, zyg_mesh** out, float* device_ms) {
    if (!dev || !out || !positions || 0 == num_vertices || positions_stride < 3) return fail("zygpu_mesh_build: invalid arguments");
    if (num_triangles < 4) return fail("zygpu_mesh_build: fewer than 4 triangles, use zyg_mesh_build");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    cudaStream_t stream = dev->streams[0];

    std::unique_ptr<zyg_mesh> mesh(new zyg_mesh);
    zyg::TriangleTree&        t = mesh->tree;
    zyg::packVertexStreams(zyg::VertexStreams{num_vertices, positions, positions_stride, normals, normals_stride, uvs, uvs_stride}, t);

    // the caller's triangle list, part by part (shape_provider.zig:863-898)
    std::vector<uint32_t> tri_indices(size_t(num_triangles) * 3, 0u);
    std::vector<uint16_t> tri_parts(num_triangles, 0);
    const uint32_t        empty_part[3] = {0, num_triangles * 3, 0};
    const uint32_t*       ps            = (num_parts > 0 && parts) ? parts : empty_part;
    const uint32_t        np            = num_parts > 0 ? num_parts : 1;
    for (uint32_t p = 0; p < np; ++p) {
        const uint32_t begin = ps[p * 3] / 3, end = std::min((ps[p * 3] + ps[p * 3 + 1]) / 3, num_triangles);
        for (uint32_t i = begin; i < end; ++i) {
            for (uint32_t k = 0; k < 3; ++k) tri_indices[size_t(i) * 3 + k] = indices ? indices[size_t(i) * 3 + k] : i * 3 + k;
            tri_parts[i] = uint16_t(p);
        }
    }
    for (uint32_t v : tri_indices) {
        if (v >= num_vertices) return fail("zygpu_mesh_build: vertex index out of range");
    }

    DeviceBuffers          buffers;
    zygpu::MeshBuildInput  in{};
    uint32_t*              d_indices   = nullptr;
    uint16_t*              d_parts     = nullptr;
    float*                 d_positions = nullptr;
    CUDA_OK(buffers.upload(d_indices, tri_indices.data(), tri_indices.size() * 4, stream));
    CUDA_OK(buffers.upload(d_parts, tri_parts.data(), tri_parts.size() * 2, stream));
    CUDA_OK(buffers.upload(d_positions, t.positions.data(), t.positions.size() * 4, stream));
    in.indices       = d_indices;
    in.parts         = d_parts;
    in.positions     = d_positions;
    in.num_triangles = num_triangles;
    in.num_vertices  = num_vertices;

    zygpu::MeshBuildOutput built;
    const cudaError_t      e = zygpu::buildMeshOnDevice(in, built, stream);
    if (cudaSuccess != e) {
        zygpu::freeMeshBuildOutput(built);
        return fail("zygpu_mesh_build: %s", cudaGetErrorString(e));
    }
    // the traversal stacks are sized for the depth the host builder produces; a degenerate Morton tree could exceed them
    if (built.binary_max_depth > zygpu::kBinaryStack || 2 * built.wide_max_depth + 8 > zygpu::kWideStack) {
        const uint32_t bd = built.binary_max_depth, wd = built.wide_max_depth;
        zygpu::freeMeshBuildOutput(built);
        return fail("zygpu_mesh_build: tree too deep for the traversal stacks (binary %u, wide %u levels): use zyg_mesh_build", bd, wd);
    }

    t.num_source_triangles = num_triangles;
    t.num_parts            = np;
    t.nodes.resize(built.num_binary_nodes);
    t.triangles.resize(size_t(num_triangles) * 3);
    t.original.resize(num_triangles);
    t.triangle_parts.resize(num_triangles);
    mesh->wide.nodes.resize(built.num_wide_nodes);
    mesh->wide.triangles.resize(num_triangles);
    const cudaError_t copies[6] = {
        cudaMemcpyAsync(t.nodes.data(), built.binary_nodes, t.nodes.size() * sizeof(zyg::BvhNode), cudaMemcpyDeviceToHost, stream),
        cudaMemcpyAsync(t.triangles.data(), built.triangles, t.triangles.size() * 4, cudaMemcpyDeviceToHost, stream),
        cudaMemcpyAsync(t.original.data(), built.original, t.original.size() * 4, cudaMemcpyDeviceToHost, stream),
        cudaMemcpyAsync(t.triangle_parts.data(), built.triangle_parts, t.triangle_parts.size() * 2, cudaMemcpyDeviceToHost, stream),
        cudaMemcpyAsync(mesh->wide.nodes.data(), built.wide_nodes, mesh->wide.nodes.size() * sizeof(zyg::WideNode), cudaMemcpyDeviceToHost, stream),
        cudaMemcpyAsync(mesh->wide.triangles.data(), built.wide_tris, mesh->wide.triangles.size() * sizeof(zyg::TriRecord), cudaMemcpyDeviceToHost,
                        stream)};
    const cudaError_t synced = cudaStreamSynchronize(stream);
    mesh->wide.max_depth     = built.wide_max_depth;
    for (int a = 0; a < 3; ++a) mesh->wide.bound_center[a] = built.bound_center[a];
    mesh->wide.bound_radius = built.bound_radius;
    if (device_ms) *device_ms = built.device_ms;
    zygpu::freeMeshBuildOutput(built);
    for (cudaError_t c : copies) CUDA_OK(c);
    CUDA_OK(synced);

    *out = mesh.release();
    return 0;
}

int zygpu_mesh_refit(zygpu_device* dev, zyg_mesh* mesh, const float* positions, uint32_t positions_stride, const float* normals,
                     uint32_t normals_stride, float* device_ms) {
    if (!dev || !mesh || !positions || positions_stride < 3) return fail("zygpu_mesh_refit: invalid arguments");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    cudaStream_t stream = dev->streams[0];

    // the moved vertices (and their normals) replace the packed streams; uvs and the topology stay
    zyg::TriangleTree&   t = mesh->tree;
    std::vector<float>   uvs;
    uvs.swap(t.uvs);
    std::vector<uint16_t> old_normals;
    if (!normals) old_normals = t.normals;
    zyg::packVertexStreams(zyg::VertexStreams{t.num_vertices, positions, positions_stride, normals, normals_stride, nullptr, 0}, t);
    t.uvs.swap(uvs);
    if (!normals) t.normals.swap(old_normals);

    const int id = zygpu_upload_mesh(dev, mesh);
    if (id < 0) return -1;
    DeviceMesh& dm = dev->meshes[size_t(id)];
    CUDA_OK(cudaMemcpyAsync(dm.buffers[4], t.positions.data(), t.positions.size() * 4, cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaMemcpyAsync(dm.buffers[5], t.normals.data(), t.normals.size() * 2, cudaMemcpyHostToDevice, stream));

    const std::vector<uint32_t> levels = wideLevelOffsets(mesh->wide.nodes);
    const uint32_t              depth  = binaryDepth(t.nodes);

    cudaEvent_t ev0, ev1;
    CUDA_OK(cudaEventCreate(&ev0));
    CUDA_OK(cudaEventCreate(&ev1));
    CUDA_OK(cudaEventRecord(ev0, stream));
    float             bound[4] = {0.f, 0.f, 0.f, 0.f};
    const cudaError_t e = zygpu::refitMeshOnDevice(static_cast<float4*>(dm.buffers[0]), static_cast<float4*>(dm.buffers[1]),
                                                   static_cast<float4*>(dm.buffers[2]), static_cast<const uint32_t*>(dm.buffers[3]),
                                                   static_cast<const float*>(dm.buffers[4]), levels.data(), uint32_t(levels.size() - 1),
                                                   uint32_t(t.nodes.size()), depth, bound, stream);
    if (cudaSuccess != e) return fail("zygpu_mesh_refit: %s", cudaGetErrorString(e));
    CUDA_OK(cudaEventRecord(ev1, stream));
    CUDA_OK(cudaMemcpyAsync(mesh->wide.nodes.data(), dm.buffers[0], mesh->wide.nodes.size() * sizeof(zyg::WideNode), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaMemcpyAsync(mesh->wide.triangles.data(), dm.buffers[1], mesh->wide.triangles.size() * sizeof(zyg::TriRecord), cudaMemcpyDeviceToHost,
                            stream));
    CUDA_OK(cudaMemcpyAsync(t.nodes.data(), dm.buffers[2], t.nodes.size() * sizeof(zyg::BvhNode), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    if (device_ms) CUDA_OK(cudaEventElapsedTime(device_ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    for (int a = 0; a < 3; ++a) mesh->wide.bound_center[a] = bound[a];
    mesh->wide.bound_radius = bound[3];
    return 0;
}

}  // extern "C"
