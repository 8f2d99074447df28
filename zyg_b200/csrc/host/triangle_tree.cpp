#include "triangle_tree.hpp"

#include <algorithm>

namespace zyg {

namespace {

// encoding.zig:81-86 (octEncode) + :65-78 (floatToSnorm16, vector form)
inline void compressNormal(Vec4f n, uint16_t out[2]) {
    const float inorm = 1.f / (std::fabs(n[0]) + std::fabs(n[1]) + std::fabs(n[2]));
    const float t     = fmax_(n[2], 0.f);
    for (int i = 0; i < 2; ++i) {
        const float v = n[i];
        const float o = (v + (v > 0.f ? t : -t)) * inorm;
        const float s = (o + 1.f) * (o > 0.f ? 32767.5f : 32768.f);
        out[i]        = uint16_t(s);
    }
}

struct Serializer {  // triangle_tree_builder.zig:166-207
    const BuildResult&                build;
    const std::vector<IndexTriangle>& source;
    TriangleTree&                     tree;
    uint32_t                          current_node     = 0;
    uint32_t                          current_triangle = 0;

    void run(uint32_t source_node, uint32_t dest_node) {
        // Explicit stack instead of recursion: SAH trees over a million references get deep.
        struct Item {
            uint32_t src, dst;
        };
        std::vector<Item> stack;
        stack.push_back({source_node, dest_node});
        while (!stack.empty()) {
            const Item it = stack.back();
            stack.pop_back();

            const BvhNode node = build.build_nodes[it.src];
            BvhNode       n    = node;

            if (0 == node.numIndices()) {
                const uint32_t child0 = current_node;
                n.setSplitNode(child0);
                tree.nodes[it.dst] = n;
                current_node += 2;

                const uint32_t source_child0 = node.children();
                // child0 subtree is serialised completely before child1's (depth-first); node ids are
                // handed out in visiting order, so push child1 first.
                stack.push_back({source_child0 + 1, child0 + 1});
                stack.push_back({source_child0, child0});
            } else {
                const uint32_t begin = node.children();
                const uint32_t num   = node.numIndices();
                uint32_t       i     = current_triangle;

                // The reference keeps `begin` (the builder's reference offset) in the node while writing
                // triangles at the running counter; both agree whenever leaves were emitted in
                // depth-first order, which is the case for every tree whose main kernel emits no leaf
                // above the task depth. We store the running counter and count disagreements.
                if (begin != i) tree.num_leaf_order_fixups += 1;
                n.setLeafNode(i, num);
                tree.nodes[it.dst] = n;

                for (uint32_t p = begin; p < begin + num; ++p, ++i) {
                    const uint32_t       src = build.reference_ids[p];
                    const IndexTriangle& t   = source[src];
                    tree.triangles[i * 3 + 0] = t.i[0];
                    tree.triangles[i * 3 + 1] = t.i[1];
                    tree.triangles[i * 3 + 2] = t.i[2];
                    tree.triangle_parts[i]    = uint16_t(t.part);
                    tree.original[i]          = src;
                }
                current_triangle = i;
            }
        }
    }
};

}  // namespace

// triangle_data.zig:40-55 + vertex_buffer.zig:221-245 (CAPI.copy): packed positions (+1 pad float), oct-encoded normals, uvs
void packVertexStreams(const VertexStreams& vertices, TriangleTree& tree) {
    const uint32_t nv = vertices.num_vertices;
    tree.num_vertices = nv;
    tree.positions.assign(size_t(nv) * 3 + 1, 0.f);
    tree.normals.assign(size_t(nv) * 2, 0);
    tree.uvs.assign(size_t(nv) * 2, 0.f);
    for (uint32_t i = 0; i < nv; ++i) {
        const size_t s          = size_t(i) * vertices.positions_stride;
        tree.positions[i * 3 + 0] = vertices.positions[s + 0];
        tree.positions[i * 3 + 1] = vertices.positions[s + 1];
        tree.positions[i * 3 + 2] = vertices.positions[s + 2];

        Vec4f n = {{0.f, 0.f, 1.f, 0.f}};
        if (vertices.normals) {
            const size_t ns = size_t(i) * vertices.normals_stride;
            n               = {{vertices.normals[ns + 0], vertices.normals[ns + 1], vertices.normals[ns + 2], 0.f}};
        }
        compressNormal(n, &tree.normals[size_t(i) * 2]);

        if (vertices.uvs) {
            const size_t us      = size_t(i) * vertices.uvs_stride;
            tree.uvs[i * 2 + 0] = vertices.uvs[us + 0];
            tree.uvs[i * 2 + 1] = vertices.uvs[us + 1];
        }
    }
}

void buildTriangleTree(const std::vector<IndexTriangle>& triangles, const VertexStreams& vertices,
                       uint32_t num_threads, TriangleTree& tree) {
    const uint32_t num_triangles = uint32_t(triangles.size());

    auto position = [&](uint32_t i) -> Vec4f {  // vertex_buffer.zig:42-46
        const size_t id = size_t(i) * vertices.positions_stride;
        return {{vertices.positions[id + 0], vertices.positions[id + 1], vertices.positions[id + 2], 0.f}};
    };

    // triangle_tree_builder.zig:112-135 (ReferencesContext.run)
    std::vector<Reference> references(num_triangles);
    AABB                   bounds = AABB::empty();
    for (uint32_t r = 0; r < num_triangles; ++r) {
        const IndexTriangle& t   = triangles[r];
        const Vec4f          a   = position(t.i[0]);
        const Vec4f          b   = position(t.i[1]);
        const Vec4f          c   = position(t.i[2]);
        const Vec4f          min = min4(a, min4(b, c));  // triangle.zig:18-24
        const Vec4f          max = max4(a, max4(b, c));
        references[r].set(min, max, r);
        bounds.b[0] = min4(bounds.b[0], min);
        bounds.b[1] = max4(bounds.b[1], max);
    }

    BuildResult build;
    buildBinaryBvh(std::move(references), bounds, 16, 64, 4, num_threads, build);

    const uint32_t num_tree_triangles = uint32_t(build.reference_ids.size());
    const uint32_t nv                 = vertices.num_vertices;

    tree.num_vertices          = nv;
    tree.num_source_triangles  = num_triangles;
    tree.num_degenerate_leaves = build.num_degenerate_leaves;
    tree.nodes.assign(build.build_nodes.size(), BvhNode{});
    tree.triangles.assign(size_t(num_tree_triangles) * 3, 0);
    tree.triangle_parts.assign(num_tree_triangles, 0);
    tree.original.assign(num_tree_triangles, 0);

    packVertexStreams(vertices, tree);

    Serializer s{build, triangles, tree};
    s.current_node = 1;  // super.newNode() before serialize, triangle_tree_builder.zig:63
    s.run(0, 0);
}

}  // namespace zyg
