// Device-side build of a triangle mesh's traversal data: what builder_base.zig:65-283 +
// triangle_tree_builder.zig:33-207 + the wide-node collapse of host/wide_bvh.cpp produce on the host, made on the GPU so that
// Scene.compile stops dominating the time to the first pixel. The tree is an LBVH (Morton order + Karras' binary radix tree),
// not the reference's SAH / spatial-split tree: the set of triangles is the same, the visiting order and the leaf boxes differ.
// The outputs use the same layouts as the host path (32-byte bvh.Node array, tree-order triangle list, 96-byte wide nodes,
// 64-byte triangle records), so every traversal and shading kernel runs on them unchanged.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace zygpu {

struct MeshBuildInput {
    const uint32_t* indices;    // device, 3 per triangle
    const uint16_t* parts;      // device, 1 per triangle
    const float*    positions;  // device, 3 per vertex, tightly packed
    uint32_t        num_triangles;
    uint32_t        num_vertices;
};

// Device buffers owned by the result (cudaFree each one, or hand them to a DeviceMesh).
struct MeshBuildOutput {
    float4*   wide_nodes     = nullptr;  // 6 float4 per node
    float4*   wide_tris      = nullptr;  // 4 float4 per record, one per triangle
    float4*   binary_nodes   = nullptr;  // 2 float4 per node, 2 * num_triangles - 1 slots (unused slots are zero)
    uint32_t* triangles      = nullptr;  // 3 per tree-order triangle
    uint32_t* original       = nullptr;  // tree-order triangle -> index in the caller's list
    uint16_t* triangle_parts = nullptr;
    uint32_t  num_wide_nodes   = 0;
    uint32_t  num_binary_nodes = 0;
    uint32_t  wide_max_depth   = 0;
    uint32_t  binary_max_depth = 0;
    float     bound_center[3]  = {0.f, 0.f, 0.f};
    float     bound_radius     = 0.f;
    float     aabb_min[3]      = {0.f, 0.f, 0.f};
    float     aabb_max[3]      = {0.f, 0.f, 0.f};
    float     device_ms        = 0.f;  // CUDA-event time of the whole build on `stream`
};

// Synchronises `stream` (a few small device -> host reads per tree level). `num_triangles` >= 4.
cudaError_t buildMeshOnDevice(const MeshBuildInput& in, MeshBuildOutput& out, cudaStream_t stream);
void        freeMeshBuildOutput(MeshBuildOutput& out);

// Refit after the vertices moved (same topology): recomputes the triangle records, the leaf gates and every quantised child box of
// the wide tree bottom-up, and the boxes of the binary tree. `level_offsets` (host): first wide node of every level, plus the
// total as the last entry, in breadth-first order (both builders number wide nodes level by level).
cudaError_t refitMeshOnDevice(float4* wide_nodes, float4* wide_tris, float4* binary_nodes, const uint32_t* triangles, const float* positions,
                              const uint32_t* level_offsets, uint32_t num_levels, uint32_t num_binary_nodes, uint32_t binary_max_depth,
                              float bound[4], cudaStream_t stream);

}  // namespace zygpu
