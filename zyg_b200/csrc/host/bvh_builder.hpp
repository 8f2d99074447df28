// SAH + spatial-split binary BVH builder of the host-side scene compile.
//
// Restates src/core/scene/bvh/builder_base.zig and split_candidate.zig so the binary tree (and
// with it the BVH-order `primitive` ids the C API hands back) is the one the Zig host builds:
// same candidate planes, same cost, same leaf rules, same sub-tree task decomposition and the
// same node / reference numbering after the tasks are appended. The device layout is derived
// from this tree (wide_bvh.hpp); the tree itself is also uploaded for the order-exact kernels.
#pragma once

#include "zmath.hpp"

#include <vector>

namespace zyg {

// split_candidate.zig:9-76
struct Reference {
    float    min[3];
    uint32_t index;
    float    max[3];
    uint32_t pad;

    void set(Vec4f mi, Vec4f ma, uint32_t prim) {
        for (int i = 0; i < 3; ++i) {
            min[i] = mi[i];
            max[i] = ma[i];
        }
        index = prim;
        pad   = 0;
    }
    AABB     aabb() const { return {{{{min[0], min[1], min[2], 0.f}}, {{max[0], max[1], max[2], 0.f}}}}; }
    uint32_t primitive() const { return index; }
};

struct BuildSettings {
    uint32_t num_slices;
    uint32_t sweep_threshold;
    uint32_t max_primitives;
    uint32_t spatial_split_threshold = 0;
    uint32_t parallel_build_depth    = 0;
};

struct BuildResult {
    std::vector<BvhNode>  build_nodes;    // builder order (children adjacent), before serialisation
    std::vector<uint32_t> reference_ids;  // leaf payload: primitive ids, duplicates after spatial splits
    uint32_t              num_degenerate_leaves = 0;  // nodes the reference would have left unsplittable
};

// builder_base.zig:323-352 (Base.split) + :354-390 (workOnTasks). `num_threads` only changes
// wall time, never the result.
void buildBinaryBvh(std::vector<Reference>&& references, const AABB& bounds, uint32_t num_slices,
                    uint32_t sweep_threshold, uint32_t max_primitives, uint32_t num_threads, BuildResult& out);

}  // namespace zyg
