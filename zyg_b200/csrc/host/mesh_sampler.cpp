#include "mesh_sampler.hpp"

#include <cmath>

namespace zyg {

namespace {

// Data.position, triangle_data.zig:62-64: a 16-byte load from the tightly packed positions; lane 3 is the next float
Vec4f position(const TriangleTree& tree, uint32_t index) {
    const float* p = tree.positions.data() + size_t(index) * 3;
    return {{p[0], p[1], p[2], p[3]}};
}

struct Tri {
    Vec4f a, b, c;
};
Tri triangleP(const TriangleTree& tree, uint32_t t) {
    return {position(tree, tree.triangles[3 * size_t(t)]), position(tree, tree.triangles[3 * size_t(t) + 1]),
            position(tree, tree.triangles[3 * size_t(t) + 2])};
}

float triangleArea(const Tri& t) { return 0.5f * length3(cross3(t.b - t.a, t.c - t.a)); }  // triangle.zig:151-153
Vec4f triangleNormal(const Tri& t) { return normalize3(cross3(t.b - t.a, t.c - t.a)); }      // triangle_data.zig:140-149
AABB  triangleAabb(const Tri& t) { return {{min4(t.a, min4(t.b, t.c)), max4(t.a, max4(t.b, t.c))}}; }  // :170-176

}  // namespace

void meshPartTables(const TriangleTree& tree, std::vector<uint32_t>& primitive_mapping, std::vector<float>& part_areas) {
    const uint32_t num_triangles = tree.numTriangles();
    primitive_mapping.resize(num_triangles);
    std::vector<uint32_t> counts(tree.num_parts, 0);
    part_areas.assign(tree.num_parts, 0.f);
    for (uint32_t t = 0; t < num_triangles; ++t) {
        const uint32_t part = tree.triangle_parts[t];
        primitive_mapping[t] = counts[part]++;
        part_areas[part] += triangleArea(triangleP(tree, t));  // calculateAreas, :705-720
    }
}

void buildMeshSampler(const TriangleTree& tree, uint32_t part, bool two_sided, MeshSamplerData& out) {
    out           = MeshSamplerData{};
    out.two_sided = two_sided;

    const uint32_t len = tree.numTriangles();
    for (uint32_t t = 0; t < len; ++t) {
        if (tree.triangle_parts[t] == part) out.triangle_mapping.push_back(t);
    }
    const uint32_t num = uint32_t(out.triangle_mapping.size());

    // EvalContext.run as one task (:167-224): uniform emission => the power of a triangle is its area
    std::vector<float> powers(num);
    std::vector<AABB>  aabbs(num);
    std::vector<Vec4f> normals(num);
    AABB               bb            = AABB::empty();
    Vec4f              dominant_axis = splat(0.f);
    float              total_power   = 0.f;
    for (uint32_t i = 0; i < num; ++i) {
        const Tri   tri = triangleP(tree, out.triangle_mapping[i]);
        const float pow = triangleArea(tri);
        powers[i]       = pow;
        aabbs[i]        = triangleAabb(tri);
        const Vec4f n   = triangleNormal(tri);
        normals[i]      = {{n[0], n[1], n[2], 1.f}};  // MeshImpl.lightCone, shape_sampler.zig:187-191
        if (pow > 0.f) {
            dominant_axis = dominant_axis + splat(pow) * n;
            bb.mergeAssign(aabbs[i]);
            total_power += pow;
        }
    }

    if (dominant_axis[0] == dominant_axis[1] && dominant_axis[1] == dominant_axis[2]) {
        out.cone = {{0.f, 0.f, 1.f, -1.f}};
    } else {
        const Vec4f da    = normalize3(dominant_axis / splat(total_power));
        float       angle = 0.f;
        for (uint32_t i = 0; i < num; ++i) {
            const float c = dot3(da, normals[i]);
            angle         = fmax_(angle, std::acos(c));
        }
        out.cone = {{da[0], da[1], da[2], std::cos(angle)}};
    }
    out.aabb = bb;

    // Distribution1D.precomputePdfCdf, src/base/math/distribution_1d.zig:91-124
    float integral = 0.f;
    for (float d : powers) integral += d;
    out.power = integral;
    std::vector<float> cdf(size_t(num) + 1, 0.f);
    if (0.f == integral) {
        out.triangle_pdfs.assign(num, 0.f);
    } else {
        const float ii = 1.f / integral;
        float       p  = 0.f;
        for (uint32_t i = 0; i + 1 < num; ++i) {
            const float c = std::fmaf(powers[i], ii, p);
            cdf[i + 1]    = c;
            p             = c;
        }
        cdf[num] = 1.f;
        out.triangle_pdfs.resize(num);
        for (uint32_t i = 0; i < num; ++i) out.triangle_pdfs[i] = cdf[i + 1] - cdf[i];  // pdfI, :83-85
    }

    // Builder.buildPrimitive over MeshImpl.lightAabb / lightCone / lightPower (= pdfI), shape_sampler.zig:182-196
    const LightSet set{aabbs.data(), normals.data(), out.triangle_pdfs.data(), nullptr, two_sided, true};
    buildPrimitiveLightTree(set, num, out.aabb, out.cone, out.power, out.tree);
}

}  // namespace zyg
