// C-ABI of the device backend (include/zygpu.h). CUDA runtime API underneath; no torch types.
#include "../../../include/zygpu.h"

#include "device_state.hpp"
#include "../host/mesh_handle.hpp"
#include "../host/wide_bvh.hpp"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <memory>
#include <string>
#include <thread>
#include <vector>



namespace {

// shape_provider.zig:863-898 (buildDescAsync): triangles are filled part by part
int fillIndexTriangles(const char* who, uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles, const uint32_t* indices,
                       uint32_t num_vertices, std::vector<zyg::IndexTriangle>& triangles, uint32_t& np) {
    triangles.assign(num_triangles, zyg::IndexTriangle{{0, 0, 0}, 0});
    const uint32_t  empty_part[3] = {0, num_triangles * 3, 0};
    const uint32_t* ps            = (num_parts > 0 && parts) ? parts : empty_part;
    np                            = num_parts > 0 ? num_parts : 1;
    for (uint32_t p = 0; p < np; ++p) {
        const uint32_t start_index = ps[p * 3 + 0];
        const uint32_t num_indices = ps[p * 3 + 1];
        const uint32_t begin       = start_index / 3;
        const uint32_t end         = std::min((start_index + num_indices) / 3, num_triangles);
        for (uint32_t i = begin; i < end; ++i) {
            const uint32_t t = i * 3;
            if (indices) {
                triangles[i].i[0] = indices[t + 0];
                triangles[i].i[1] = indices[t + 1];
                triangles[i].i[2] = indices[t + 2];
            } else {
                triangles[i].i[0] = t + 0;
                triangles[i].i[1] = t + 1;
                triangles[i].i[2] = t + 2;
            }
            triangles[i].part = p;
        }
    }
    for (const zyg::IndexTriangle& t : triangles) {
        if (t.i[0] >= num_vertices || t.i[1] >= num_vertices || t.i[2] >= num_vertices) {
            return fail("%s: vertex index out of range", who);
        }
    }
    return 0;
}

}  // namespace

extern "C" {

const char* zygpu_last_error(void) { return zygpuError().c_str(); }

int zyg_mesh_build(uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles, const uint32_t* indices,
                   uint32_t num_vertices, const float* positions, uint32_t positions_stride, const float* normals,
                   uint32_t normals_stride, const float* uvs, uint32_t uvs_stride, uint32_t num_threads,
                   zyg_mesh** out) {
    if (!out || !positions || 0 == num_triangles || 0 == num_vertices || positions_stride < 3) {
        return fail("zyg_mesh_build: invalid arguments");
    }

    std::vector<zyg::IndexTriangle> triangles;
    uint32_t                        np = 1;
    if (0 != fillIndexTriangles("zyg_mesh_build", num_parts, parts, num_triangles, indices, num_vertices, triangles, np)) return -1;

    if (0 == num_threads) num_threads = std::max(1u, std::thread::hardware_concurrency());

    std::unique_ptr<zyg_mesh> mesh(new zyg_mesh);
    zyg::VertexStreams        vs{num_vertices, positions, positions_stride, normals, normals_stride, uvs, uvs_stride};
    zyg::buildTriangleTree(triangles, vs, num_threads, mesh->tree);
    mesh->tree.num_parts = np;
    zyg::buildWideBvh(mesh->tree, mesh->wide);

    *out = mesh.release();
    return 0;
}

void zyg_mesh_free(zyg_mesh* mesh) { delete mesh; }

const void* zyg_mesh_data(const zyg_mesh* mesh, int which, uint64_t* num_bytes) {
    if (!mesh) return nullptr;
    const void* p = nullptr;
    uint64_t    n = 0;
    const auto& t = mesh->tree;
    switch (which) {
        case ZYG_MESH_BINARY_NODES: p = t.nodes.data(), n = t.nodes.size() * sizeof(zyg::BvhNode); break;
        case ZYG_MESH_TRIANGLES: p = t.triangles.data(), n = t.triangles.size() * 4; break;
        case ZYG_MESH_ORIGINAL: p = t.original.data(), n = t.original.size() * 4; break;
        case ZYG_MESH_POSITIONS: p = t.positions.data(), n = t.positions.size() * 4; break;
        case ZYG_MESH_NORMALS: p = t.normals.data(), n = t.normals.size() * 2; break;
        case ZYG_MESH_UVS: p = t.uvs.data(), n = t.uvs.size() * 4; break;
        case ZYG_MESH_PARTS: p = t.triangle_parts.data(), n = t.triangle_parts.size() * 2; break;
        case ZYG_MESH_WIDE_NODES: p = mesh->wide.nodes.data(), n = mesh->wide.nodes.size() * sizeof(zyg::WideNode); break;
        case ZYG_MESH_WIDE_TRIS:
            p = mesh->wide.triangles.data(), n = mesh->wide.triangles.size() * sizeof(zyg::TriRecord);
            break;
        default: return nullptr;
    }
    if (num_bytes) *num_bytes = n;
    return p;
}

int zyg_mesh_info(const zyg_mesh* mesh, ZygMeshInfo* info) {
    if (!mesh || !info) return fail("zyg_mesh_info: null argument");
    const auto& t               = mesh->tree;
    info->num_source_triangles  = t.num_source_triangles;
    info->num_tree_triangles    = t.numTriangles();
    info->num_vertices          = t.num_vertices;
    info->num_binary_nodes      = uint32_t(t.nodes.size());
    info->num_wide_nodes        = uint32_t(mesh->wide.nodes.size());
    info->wide_max_depth        = mesh->wide.max_depth;
    info->num_degenerate_leaves = t.num_degenerate_leaves;
    info->num_leaf_order_fixups = t.num_leaf_order_fixups;
    for (int i = 0; i < 3; ++i) {
        info->aabb_min[i] = t.nodes[0].min[i];
        info->aabb_max[i] = t.nodes[0].max[i];
    }
    return 0;
}

int zygpu_create(int device_ordinal, zygpu_device** out) {
    if (!out) return fail("zygpu_create: null out");
    int count = 0;
    CUDA_OK(cudaGetDeviceCount(&count));
    if (device_ordinal < 0 || device_ordinal >= count) return fail("zygpu_create: no CUDA device %d", device_ordinal);
    CUDA_OK(cudaSetDevice(device_ordinal));

    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major != 10) {
        return fail("zygpu_create: device %d is sm_%d%d; this library carries sm_100a code only", device_ordinal,
                    prop.major, prop.minor);
    }

    std::unique_ptr<zygpu_device> dev(new zygpu_device);
    dev->ordinal = device_ordinal;
    for (int s = 0; s < zygpu_device::kStreams; ++s) {
        CUDA_OK(cudaStreamCreateWithFlags(&dev->streams[s], cudaStreamNonBlocking));
        CUDA_OK(cudaMalloc(&dev->d_rays[s], zygpu_device::kChunkRays * sizeof(ZygpuRay)));
        CUDA_OK(cudaMalloc(&dev->d_out[s], zygpu_device::kChunkRays * sizeof(ZygpuHit)));
    }
    CUDA_OK(cudaMalloc(&dev->d_counters, sizeof(zygpu::TraceCounters)));
    CUDA_OK(cudaMalloc(&dev->d_work, zygpu_device::kWorkCounters * sizeof(uint32_t)));
    // traversal stacks of the ray-pool kernel: one region for the device-buffer entry point (as many blocks as can be resident), one
    // smaller region per staging stream of the host-buffer entry point (its chunk kernels run concurrently)
    dev->stack_bytes_main   = zygpu::traceStackBytesPerBlock() * size_t(prop.multiProcessorCount) * 8;
    dev->stack_bytes_stream = zygpu::traceStackBytesPerBlock() * size_t(prop.multiProcessorCount) * 4;
    CUDA_OK(cudaMalloc(&dev->d_stacks, dev->stack_bytes_main + zygpu_device::kStreams * dev->stack_bytes_stream));
    *out = dev.release();
    return 0;
}

void zygpu_destroy(zygpu_device* dev) {
    if (!dev) return;
    cudaSetDevice(dev->ordinal);
    cudaDeviceSynchronize();
    for (DeviceMesh& m : dev->meshes) {
        for (void* b : m.buffers) cudaFree(b);
    }
    for (int s = 0; s < zygpu_device::kStreams; ++s) {
        cudaFree(dev->d_rays[s]);
        cudaFree(dev->d_out[s]);
        if (dev->streams[s]) cudaStreamDestroy(dev->streams[s]);
    }
    cudaFree(dev->d_counters);
    cudaFree(dev->d_work);
    cudaFree(dev->d_stacks);
    zygpuReleaseRender(dev);
    delete dev;
}

int zygpu_upload_mesh(zygpu_device* dev, const zyg_mesh* mesh) {
    if (!dev || !mesh) return fail("zygpu_upload_mesh: null argument");
    CUDA_OK(cudaSetDevice(dev->ordinal));

    DeviceMesh  dm;
    const auto& t = mesh->tree;
    const auto& w = mesh->wide;

    // an already uploaded mesh keeps its id
    for (size_t i = 0; i < dev->meshes.size(); ++i) {
        if (dev->meshes[i].source == mesh->serial) return int(i);
    }

    const void*  src[8]   = {w.nodes.data(),     w.triangles.data(), t.nodes.data(), t.triangles.data(),
                             t.positions.data(), t.normals.data(),   t.uvs.data(),   t.triangle_parts.data()};
    const size_t bytes[8] = {w.nodes.size() * sizeof(zyg::WideNode), w.triangles.size() * sizeof(zyg::TriRecord),
                             t.nodes.size() * sizeof(zyg::BvhNode),  t.triangles.size() * 4,
                             t.positions.size() * 4,                 t.normals.size() * 2,
                             t.uvs.size() * 4,                       t.triangle_parts.size() * 2};
    for (int i = 0; i < 8; ++i) {
        CUDA_OK(cudaMalloc(&dm.buffers[i], std::max<size_t>(bytes[i], 16)));
        CUDA_OK(cudaMemcpy(dm.buffers[i], src[i], bytes[i], cudaMemcpyHostToDevice));
    }
    dm.source              = mesh->serial;
    dm.view.wide_nodes     = static_cast<const float4*>(dm.buffers[0]);
    dm.view.wide_tris      = static_cast<const float4*>(dm.buffers[1]);
    dm.view.binary_nodes   = static_cast<const float4*>(dm.buffers[2]);
    dm.view.triangles      = static_cast<const uint32_t*>(dm.buffers[3]);
    dm.view.positions      = static_cast<const float*>(dm.buffers[4]);
    dm.view.num_wide_nodes = uint32_t(w.nodes.size());
    dm.view.num_tris       = uint32_t(w.triangles.size());
    dm.shading.triangles   = dm.view.triangles;
    dm.shading.positions   = dm.view.positions;
    dm.shading.normals     = static_cast<const uint16_t*>(dm.buffers[5]);
    dm.shading.uvs         = static_cast<const float*>(dm.buffers[6]);
    dm.shading.parts       = static_cast<const uint16_t*>(dm.buffers[7]);

    dev->meshes.push_back(dm);
    return int(dev->meshes.size() - 1);
}

int zygpu_trace_batch_device(zygpu_device* dev, int mesh, int mode, const void* d_rays, uint64_t n, void* d_out,
                             void* stream, ZygpuTraceCounters* counters) {
    if (!dev || mesh < 0 || size_t(mesh) >= dev->meshes.size()) return fail("zygpu_trace_batch_device: bad mesh id");
    if (n > 0xFFFFFFFFull) return fail("zygpu_trace_batch_device: at most 2^32-1 rays per call");
    CUDA_OK(cudaSetDevice(dev->ordinal));
    cudaStream_t s = static_cast<cudaStream_t>(stream);

    if (counters) CUDA_OK(cudaMemsetAsync(dev->d_counters, 0, sizeof(zygpu::TraceCounters), s));

    CUDA_OK(zygpu::launchTrace(dev->meshes[mesh].view, mode, static_cast<const zygpu::RayIn*>(d_rays), d_out,
                               uint32_t(n), counters ? dev->d_counters : nullptr, dev->workCounter(), dev->d_stacks, dev->stack_bytes_main, s));

    if (counters) {
        zygpu::TraceCounters h;
        CUDA_OK(cudaMemcpyAsync(&h, dev->d_counters, sizeof(h), cudaMemcpyDeviceToHost, s));
        CUDA_OK(cudaStreamSynchronize(s));
        counters->nodes     = h.nodes;
        counters->triangles = h.triangles;
        counters->rays      = h.rays;
        counters->max_stack = h.max_stack;
    }
    return 0;
}

int zygpu_trace_batch(zygpu_device* dev, int mesh, int mode, const ZygpuRay* rays, uint64_t n, void* out) {
    if (!dev || mesh < 0 || size_t(mesh) >= dev->meshes.size()) return fail("zygpu_trace_batch: bad mesh id");
    if (n > 0 && (!rays || !out)) return fail("zygpu_trace_batch: null buffer");
    CUDA_OK(cudaSetDevice(dev->ordinal));

    const bool   any       = ZYGPU_ANY == mode || ZYGPU_ANY_BINARY == mode;
    const size_t out_bytes = any ? sizeof(uint32_t) : sizeof(ZygpuHit);

    // chunk i: H2D -> kernel -> D2H on stream i % kStreams; consecutive chunks overlap copy and compute
    uint64_t done = 0;
    for (int c = 0; done < n; ++c) {
        const int      s     = c % zygpu_device::kStreams;
        const uint64_t count = std::min<uint64_t>(zygpu_device::kChunkRays, n - done);
        cudaStream_t   st    = dev->streams[s];

        CUDA_OK(cudaMemcpyAsync(dev->d_rays[s], rays + done, count * sizeof(ZygpuRay), cudaMemcpyHostToDevice, st));
        CUDA_OK(zygpu::launchTrace(dev->meshes[mesh].view, mode, static_cast<const zygpu::RayIn*>(dev->d_rays[s]),
                                   dev->d_out[s], uint32_t(count), nullptr, dev->workCounter(),
                                   static_cast<char*>(dev->d_stacks) + dev->stack_bytes_main + size_t(s) * dev->stack_bytes_stream,
                                   dev->stack_bytes_stream, st));
        CUDA_OK(cudaMemcpyAsync(static_cast<char*>(out) + done * out_bytes, dev->d_out[s], count * out_bytes,
                                cudaMemcpyDeviceToHost, st));
        done += count;
    }
    for (int s = 0; s < zygpu_device::kStreams; ++s) CUDA_OK(cudaStreamSynchronize(dev->streams[s]));
    return 0;
}

}  // extern "C"
