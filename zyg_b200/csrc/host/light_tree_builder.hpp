// Light tree construction, restated from src/core/scene/light/light_tree_builder.zig (A. Conty, C. Kulla: Importance
// Sampling of Many Lights with Adaptive Tree Splitting). One builder serves the scene-level tree over lights
// (Builder.build, :281-376) and the per-part tree over emissive triangles (Builder.buildPrimitive, :378-428); the two differ
// in their candidate evaluation (evaluateScene / evaluateSampler, :131-263) and leaf size.
#pragma once

#include "../../../include/zygpu_scene.h"
#include "zmath.hpp"

#include <cstdint>
#include <vector>

namespace zyg {

// What the builder reads of a light (Scene.lightAabb / lightCone / lightPower / lightTwoSided, scene.zig:650-664) or of an
// emissive triangle (MeshImpl.lightAabb / lightCone / lightPower, shape_sampler.zig:240-262).
struct LightSet {
    const AABB*  aabbs;      // bounds[0][3] = power, bounds[1][3] = cached radius for scene lights
    const Vec4f* cones;      // axis xyz, cos of the half angle in w (triangles: the normal)
    const float* powers;
    const uint8_t* two_sided;  // per light; null => `all_two_sided`
    bool           all_two_sided;
    bool           primitive;  // per-part tree: evaluateSampler, leaves of up to 4 triangles

    bool twoSided(uint32_t l) const { return two_sided ? 0 != two_sided[l] : all_two_sided; }
};

struct LightTreeResult {
    std::vector<ZygpuLightNode> nodes;
    std::vector<uint32_t>       node_middles;
    std::vector<uint32_t>       light_orders;   // per light
    std::vector<uint32_t>       light_mapping;  // tree order -> light
    AABB                        bounds;         // of the root, radius cached
    uint32_t                    max_split_depth;
    float                       root_power;
};

constexpr uint32_t kLightTreeMaxSplitDepth = 10;  // Tree.MaxSplitDepth, light_tree.zig:248
constexpr uint32_t kLightTreeMaxLights     = 64;  // Tree.MaxLights, :249

// Scene-level tree over `mapping[num_infinite..]` (finite lights); `mapping` lists the infinite lights first and is
// reordered in place. Light orders of the infinite lights are assigned by the caller (first `first_order` orders).
void buildLightTree(const LightSet& set, std::vector<uint32_t>& mapping, uint32_t num_infinite, uint32_t first_order,
                    std::vector<uint32_t>& light_orders, LightTreeResult& out);

// Per-part tree over `num_triangles` emissive triangles with the part's bounds, cone and total power.
void buildPrimitiveLightTree(const LightSet& set, uint32_t num_triangles, const AABB& bounds, Vec4f cone, float total_power,
                             LightTreeResult& out);

// Optional device builder (device/light_build.cu), installed by the C-ABI layer: builds the tree over
// `lights[0..num)` (indices into `set`) and fills nodes / node_middles / bounds / root_power of `out` plus `order`: tree position ->
// index into `lights`. Leaf and middle indices already include `first_order`. Returns false when it declines (then the host builds).
using DeviceLightTreeFn = bool (*)(const LightSet& set, const uint32_t* lights, uint32_t num, uint32_t first_order, LightTreeResult& out,
                                   std::vector<uint32_t>& order);
void setDeviceLightTreeBuilder(DeviceLightTreeFn fn, uint32_t min_lights);

Vec4f coneMerge(Vec4f a, Vec4f b);  // math.cone.merge, src/base/math/cone.zig:8-44

}  // namespace zyg
