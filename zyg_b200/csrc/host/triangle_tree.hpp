// Host-side triangle mesh: indexed-triangle BVH ("BLAS") in the reference's layout.
//
// Mirrors src/core/scene/shape/triangle/{triangle_tree_builder,triangle_data,vertex_buffer}.zig:
// the tree is built over per-triangle references, leaves are rewritten in tree order, vertex
// positions are tightly packed (3 floats + 1 pad float at the end), normals are oct-encoded
// snorm16 pairs. `primitive` ids returned by the intersection entry points index `triangles`.
#pragma once

#include "bvh_builder.hpp"

#include <vector>

namespace zyg {

struct IndexTriangle {  // triangle_tree_builder.zig:18-21
    uint32_t i[3];
    uint32_t part;
};

struct TriangleTree {
    std::vector<BvhNode>  nodes;           // serialised order: children adjacent, child0 subtree first
    std::vector<uint32_t> triangles;       // 3 vertex ids per BVH-order triangle (triangle.zig:6-10)
    std::vector<uint16_t> triangle_parts;  // material part per BVH-order triangle
    std::vector<uint32_t> original;        // BVH-order triangle -> index in the caller's triangle list
    std::vector<float>    positions;       // 3 * num_vertices + 1 (triangle_data.zig:49,54)
    std::vector<uint16_t> normals;         // 2 per vertex, enc.compressNormal (encoding.zig:100-103)
    std::vector<float>    uvs;             // 2 per vertex

    uint32_t num_vertices          = 0;
    uint32_t num_source_triangles  = 0;
    uint32_t num_parts             = 1;
    uint32_t num_degenerate_leaves = 0;
    uint32_t num_leaf_order_fixups = 0;  // leaves whose builder offset differed from the serialised one

    uint32_t numTriangles() const { return uint32_t(triangles.size() / 3); }
    AABB     aabb() const { return nodes[0].aabb(); }
};

// C-API style vertex streams, strides in floats (vertex_buffer.zig:44,219; capi.zig:379-423).
struct VertexStreams {
    uint32_t     num_vertices;
    const float* positions;
    uint32_t     positions_stride;
    const float* normals;  // may be null -> (0,0,1)
    uint32_t     normals_stride;
    const float* uvs;  // may be null -> (0,0)
    uint32_t     uvs_stride;
};

// Fills positions / normals / uvs and num_vertices of `tree` from the caller's streams.
void packVertexStreams(const VertexStreams& vertices, TriangleTree& tree);

// shape_provider.zig:915-924 (16 slices, sweep 64, 4 primitives) + triangle_tree_builder.zig:33-65.
void buildTriangleTree(const std::vector<IndexTriangle>& triangles, const VertexStreams& vertices,
                       uint32_t num_threads, TriangleTree& tree);

}  // namespace zyg
