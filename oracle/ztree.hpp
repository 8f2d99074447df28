// ORACLE — test infrastructure only (see zmath.hpp). bvh.Node, NodeStack, the Moeller-Trumbore test and
// the triangle-BVH queries of the reference over plain arrays in the reference's own layout:
//   nodes      32-byte bvh.Node                    src/core/scene/bvh/node.zig:9-20
//   triangles  u32[3] per tree-order triangle      src/core/scene/shape/triangle/triangle.zig:6-10
//   positions  f32[3] per vertex, tightly packed   src/core/scene/shape/triangle/triangle_data.zig:62-64
#pragma once

#include "zmath.hpp"
#include "zyg_oracle.h"

#include <algorithm>

namespace zo {

struct Node {  // node.zig:9-20
    float    min[3];
    uint32_t min_data;
    float    max[3];
    uint32_t max_data;

    uint32_t children() const { return min_data; }
    uint32_t numIndices() const { return max_data; }
    uint32_t indicesStart() const { return min_data; }

    // node.zig:73-87. The 4th lane of min/max holds the data words in the reference; it is replaced
    // by min_t/max_t before the horizontal reduction, so its value never matters.
    float intersect(const Ray& ray) const {
        const Vec4f mi    = {{min[0], min[1], min[2], 0.f}};
        const Vec4f ma    = {{max[0], max[1], max[2], 0.f}};
        const Vec4f lower = (mi - ray.origin) * ray.inv_direction;
        const Vec4f upper = (ma - ray.origin) * ray.inv_direction;
        const Vec4f t0    = min4(lower, upper);
        const Vec4f t1    = max4(lower, upper);
        const Vec4f tmins = {{t0[0], t0[1], t0[2], ray.min_t}};
        const Vec4f tmaxs = {{t1[0], t1[1], t1[2], ray.max_t}};
        const float tboxmin = hmax4(tmins);
        const float tboxmax = hmin4(tmaxs);
        return tboxmin <= tboxmax ? tboxmin : FLT_MAX;
    }
};
static_assert(sizeof(Node) == 32, "size_test.zig:44");

struct NodeStack {  // node_stack.zig:1-30
    static constexpr uint32_t End = 0xFFFFFFFFu;
    uint32_t                  end = 0;
    uint32_t                  stack[127];
    void                      push(uint32_t v) { stack[end++] = v; }
    uint32_t                  pop() { return 0 == end ? End : stack[--end]; }
};

struct Hit {
    float t, u, v;
};

// triangle.zig:26-52
inline bool intersectTriangle(const Ray& ray, Vec4f a, Vec4f b, Vec4f c, Hit& hit) {
    const Vec4f e1 = b - a;
    const Vec4f e2 = c - a;

    const Vec4f tvec = ray.origin - a;
    const Vec4f pvec = cross3(ray.direction, e2);
    const Vec4f qvec = cross3(tvec, e1);

    const float e1_d_pv = dot3(e1, pvec);
    const float tv_d_pv = dot3(tvec, pvec);
    const float di_d_qv = dot3(ray.direction, qvec);
    const float e2_d_qv = dot3(e2, qvec);

    const float inv_det = 1.f / e1_d_pv;

    const float u     = tv_d_pv * inv_det;
    const float v     = di_d_qv * inv_det;
    const float hit_t = e2_d_qv * inv_det;

    const float uv = u + v;

    if (u >= 0.f && 1.f >= u && v >= 0.f && 1.f >= uv && hit_t >= ray.min_t && ray.max_t >= hit_t) {
        hit = {hit_t, u, v};
        return true;
    }
    return false;
}

struct Mesh {
    const Node*     nodes;
    const uint32_t* triangles;
    const float*    positions;

    // triangle_data.zig:62-64 : a 4-float load from the packed array; lane 3 is the next vertex's x
    // (or the pad float) and never reaches a result (dot3 / cross3 lanes 0-2 only).
    Vec4f position(uint32_t index) const {
        const float* p = positions + size_t(index) * 3;
        return {{p[0], p[1], p[2], 0.f}};
    }

    bool intersectIndexed(const Ray& ray, uint32_t index, Hit& hit) const {  // triangle_data.zig:66-74
        const uint32_t* tri = triangles + size_t(index) * 3;
        return intersectTriangle(ray, position(tri[0]), position(tri[1]), position(tri[2]), hit);
    }

    // triangle_tree.zig:46-109 with an identity transformation (object-space rays).
    bool intersect(Ray local_ray, ZoHit& isec, uint64_t* visited_nodes, uint64_t* tested_tris) const {
        NodeStack stack;
        uint32_t  n = 0;

        Hit      hpoint{0.f, 0.f, 0.f};
        uint32_t primitive = 0xFFFFFFFFu;

        while (NodeStack::End != n) {
            const Node& node = nodes[n];
            if (visited_nodes) ++*visited_nodes;

            const uint32_t num = node.numIndices();
            if (0 != num) {
                uint32_t       i = node.indicesStart();
                const uint32_t e = i + num;
                for (; i < e; ++i) {
                    Hit hit;
                    if (tested_tris) ++*tested_tris;
                    if (intersectIndexed(local_ray, i, hit)) {
                        local_ray.max_t = hit.t;
                        hpoint          = hit;
                        primitive       = i;
                    }
                }
                n = stack.pop();
                continue;
            }

            uint32_t a = node.children();
            uint32_t b = a + 1;

            float dista = nodes[a].intersect(local_ray);
            float distb = nodes[b].intersect(local_ray);

            if (dista > distb) {
                std::swap(a, b);
                std::swap(dista, distb);
            }

            if (FLT_MAX == dista) {
                n = stack.pop();
            } else {
                n = a;
                if (FLT_MAX != distb) stack.push(b);
            }
        }

        if (0xFFFFFFFFu == primitive) {
            isec = {local_ray.max_t, 0.f, 0.f, 0xFFFFFFFFu};
            return false;
        }
        isec = {hpoint.t, hpoint.u, hpoint.v, primitive};
        return true;
    }

    // The traversal of Tree.emission, triangle_tree.zig:405-477: every triangle the (fixed) ray interval hits, in tree order.
    template <typename Visit>
    void allHits(const Ray& ray, Visit&& visit) const {
        NodeStack stack;
        uint32_t  n = 0;
        while (NodeStack::End != n) {
            const Node&    node = nodes[n];
            const uint32_t num  = node.numIndices();
            if (0 != num) {
                uint32_t       i = node.indicesStart();
                const uint32_t e = i + num;
                for (; i < e; ++i) {
                    Hit hit;
                    if (intersectIndexed(ray, i, hit)) visit(i, hit);
                }
                n = stack.pop();
                continue;
            }
            uint32_t a = node.children();
            uint32_t b = a + 1;
            float dista = nodes[a].intersect(ray);
            float distb = nodes[b].intersect(ray);
            if (dista > distb) {
                std::swap(a, b);
                std::swap(dista, distb);
            }
            if (FLT_MAX == dista) {
                n = stack.pop();
            } else {
                n = a;
                if (FLT_MAX != distb) stack.push(b);
            }
        }
    }

    // triangle_tree.zig:197-242
    bool intersectP(const Ray& ray) const {
        NodeStack stack;
        uint32_t  n = 0;

        while (NodeStack::End != n) {
            const Node& node = nodes[n];

            const uint32_t num = node.numIndices();
            if (0 != num) {
                uint32_t       i = node.indicesStart();
                const uint32_t e = i + num;
                for (; i < e; ++i) {
                    Hit hit;
                    if (intersectIndexed(ray, i, hit)) return true;  // triangle.zig:54-80 == intersect != null
                }
                n = stack.pop();
                continue;
            }

            uint32_t a = node.children();
            uint32_t b = a + 1;

            float dista = nodes[a].intersect(ray);
            float distb = nodes[b].intersect(ray);

            if (dista > distb) {
                std::swap(a, b);
                std::swap(dista, distb);
            }

            if (FLT_MAX == dista) {
                n = stack.pop();
            } else {
                n = a;
                if (FLT_MAX != distb) stack.push(b);
            }
        }
        return false;
    }
};

}  // namespace zo
