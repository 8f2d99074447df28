// Sampling data of one emissive mesh part: shape_sampler.MeshImpl (src/core/scene/shape/shape_sampler.zig:149-262), built
// by Part.configure (shape/triangle/triangle_mesh.zig:57-149) and Mesh.prepareSampling / calculateAreas (:705-746), with
// its PrimitiveTree (light/light_tree.zig:520-719) from Builder.buildPrimitive (light_tree_builder.zig:378-428).
#pragma once

#include "light_tree_builder.hpp"
#include "triangle_tree.hpp"

#include <vector>

namespace zyg {

struct MeshSamplerData {
    AABB  aabb;   // of the emitting triangles, object space
    Vec4f cone;   // dominant axis, cos of the widest deviation
    float power;  // Distribution1D.integral = emitting area, object space
    bool  two_sided;

    std::vector<uint32_t> triangle_mapping;  // part triangle -> tree triangle (Part.triangle_mapping)
    std::vector<float>    triangle_pdfs;     // Distribution1D.pdfI per part triangle
    LightTreeResult       tree;
};

// Mesh.primitive_mapping (tree triangle -> index within its part) and Part.area per part (triangle_mesh.zig:705-746).
void meshPartTables(const TriangleTree& tree, std::vector<uint32_t>& primitive_mapping, std::vector<float>& part_areas);

// Part.configure for a material with uniform emission (every triangle of the part emits).
void buildMeshSampler(const TriangleTree& tree, uint32_t part, bool two_sided, MeshSamplerData& out);

}  // namespace zyg
