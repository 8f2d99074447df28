// Device-side light-tree build: what light_tree_builder.zig:281-428 (Builder.build / buildPrimitive) makes on the
// host, for scenes whose lights change per frame. The tree is an LBVH over the light centres (Morton order, Karras' radix tree) with
// the node statistics of the reference (bounds, cone, power, variance, two-sidedness) aggregated bottom-up, serialised into the
// reference's 32-byte light_tree.Node: Tree.randomLight / pdf and PrimitiveTree.randomLight / pdf (light_tree.zig:346-719) read it
// unchanged. It is a different tree than the reference's cost-driven one: the same estimator, other pdfs.
#pragma once

#include "../../../include/zygpu_scene.h"

#include <cstdint>
#include <cuda_runtime.h>

namespace zygpu {

struct LightBuildInput {  // device arrays, one entry per light of the tree
    const float4* aabb_min;  // xyz
    const float4* aabb_max;
    const float4* cones;     // axis xyz, cos of the half angle in w (emissive triangles: the normal)
    const float*  powers;
    const uint8_t* two_sided;  // null: `all_two_sided`
    uint32_t       num_lights;
    bool           all_two_sided;
    bool           primitive;    // per-part tree over triangles: cone = dominant axis + largest deviation, leaves of up to 4
    uint32_t       first_order;  // tree position of the first light (the scene tree lists the infinite lights before)
};

struct LightBuildOutput {  // device arrays owned by the result
    ZygpuLightNode* nodes        = nullptr;  // 2 * num_lights - 1 slots; children of a node are adjacent
    uint32_t*       node_middles = nullptr;
    uint32_t*       order        = nullptr;  // tree position (without first_order) -> index of the light in the input
    uint32_t        num_nodes    = 0;
    float           bounds_min[3] = {0.f, 0.f, 0.f};
    float           bounds_max[3] = {0.f, 0.f, 0.f};
    float           root_power    = 0.f;
    float           device_ms     = 0.f;
};

// Needs at least 2 lights. Synchronises `stream`.
cudaError_t buildLightTreeOnDevice(const LightBuildInput& in, LightBuildOutput& out, cudaStream_t stream);
void        freeLightBuildOutput(LightBuildOutput& out);

}  // namespace zygpu
