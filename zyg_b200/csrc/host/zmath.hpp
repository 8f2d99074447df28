// Host-side strict-fp32 math used by the scene-compile code (BVH builders, transforms).
//
// Mirrors the arithmetic of the reference's base/math package so that trees and world-space
// bounds come out identical to what the Zig host would produce:
//   - lane-wise Vec4f ops                           (src/base/math/vector4.zig:36-64)
//   - min/max as `x < y ? x : y` / `y < x ? x : y`  (src/base/math/util.zig:17-29)
//   - FMA only where the reference writes @mulAdd   (vector4.zig:73-92, matrix3x3.zig:123-136)
// Compile with -ffp-contract=off so the compiler never fuses on its own.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace zyg {

struct Vec4f {
    float v[4];

    float  operator[](int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
};

inline Vec4f splat(float s) { return {{s, s, s, s}}; }

inline Vec4f operator+(Vec4f a, Vec4f b) { return {{a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}}; }
inline Vec4f operator-(Vec4f a, Vec4f b) { return {{a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]}}; }
inline Vec4f operator*(Vec4f a, Vec4f b) { return {{a[0] * b[0], a[1] * b[1], a[2] * b[2], a[3] * b[3]}}; }
inline Vec4f operator/(Vec4f a, Vec4f b) { return {{a[0] / b[0], a[1] / b[1], a[2] / b[2], a[3] / b[3]}}; }
inline Vec4f operator-(Vec4f a) { return {{-a[0], -a[1], -a[2], -a[3]}}; }

// @mulAdd(Vec4f, a, b, c)
inline Vec4f mulAdd(Vec4f a, Vec4f b, Vec4f c) {
    return {{std::fmaf(a[0], b[0], c[0]), std::fmaf(a[1], b[1], c[1]), std::fmaf(a[2], b[2], c[2]),
             std::fmaf(a[3], b[3], c[3])}};
}

// util.zig:17-29 (x86 branch)
inline float fmin_(float x, float y) { return x < y ? x : y; }
inline float fmax_(float x, float y) { return y < x ? x : y; }

inline Vec4f min4(Vec4f a, Vec4f b) {
    return {{fmin_(a[0], b[0]), fmin_(a[1], b[1]), fmin_(a[2], b[2]), fmin_(a[3], b[3])}};
}
inline Vec4f max4(Vec4f a, Vec4f b) {
    return {{fmax_(a[0], b[0]), fmax_(a[1], b[1]), fmax_(a[2], b[2]), fmax_(a[3], b[3])}};
}

// vector4.zig:36-39 : (x + y) + z
inline float dot3(Vec4f a, Vec4f b) {
    const float x = a[0] * b[0], y = a[1] * b[1], z = a[2] * b[2];
    return (x + y) + z;
}
inline float length3(Vec4f a) { return std::sqrt(dot3(a, a)); }
inline Vec4f normalize3(Vec4f a) { return a / splat(length3(a)); }

// vector4.zig:73-92 : one FMA per lane on the shuffled operands
inline Vec4f cross3(Vec4f a, Vec4f b) {
    return {{std::fmaf(b[2], a[1], -(a[2] * b[1])), std::fmaf(b[0], a[2], -(a[0] * b[2])),
             std::fmaf(b[1], a[0], -(a[1] * b[0])), std::fmaf(b[3], a[3], -(a[3] * b[3]))}};
}

inline uint32_t indexMaxComponent3(Vec4f v) {  // vector4.zig:185-191
    if (v[0] > v[1]) return v[0] > v[2] ? 0 : 2;
    return v[1] > v[2] ? 1 : 2;
}

// src/base/math/aabb.zig
struct AABB {
    Vec4f b[2];

    static AABB empty() { return {{splat(FLT_MAX), splat(-FLT_MAX)}}; }

    Vec4f position() const { return splat(0.5f) * (b[0] + b[1]); }  // aabb.zig:24
    Vec4f extent() const { return b[1] - b[0]; }                    // aabb.zig:32
    float surfaceArea() const {                                     // aabb.zig:36-39
        const Vec4f d = b[1] - b[0];
        return 2.f * (d[0] * d[1] + d[0] * d[2] + d[1] * d[2]);
    }
    AABB intersection(const AABB& o) const { return {{max4(b[0], o.b[0]), min4(b[1], o.b[1])}}; }  // :199
    void mergeAssign(const AABB& o) {                                                              // :206
        b[0] = min4(b[0], o.b[0]);
        b[1] = max4(b[1], o.b[1]);
    }
    void clipMin(float d, uint8_t axis) { b[0][axis] = fmax_(d, b[0][axis]); }  // :211-222
    void clipMax(float d, uint8_t axis) { b[1][axis] = fmin_(d, b[1][axis]); }  // :224-228
    bool covers(const AABB& o) const {                                           // :230-237
        return b[0][0] <= o.b[0][0] && b[0][1] <= o.b[0][1] && b[0][2] <= o.b[0][2] && b[1][0] >= o.b[1][0] &&
               b[1][1] >= o.b[1][1] && b[1][2] >= o.b[1][2];
    }
    void translate(Vec4f t) {
        b[0] = b[0] + t;
        b[1] = b[1] + t;
    }
    void cacheRadius() {  // aabb.zig:141-145
        b[0][3] = 0.f;
        b[1][3] = 0.5f * length3(extent());
    }
};

// 32-byte binary BVH node, src/core/scene/bvh/node.zig:9-71
struct BvhNode {
    float    min[3];
    uint32_t min_data;  // children index (inner) or first primitive (leaf)
    float    max[3];
    uint32_t max_data;  // 0 => inner, else number of primitives

    uint32_t children() const { return min_data; }
    uint32_t numIndices() const { return max_data; }
    uint32_t indicesStart() const { return min_data; }
    void     setAABB(const AABB& box) {
        for (int i = 0; i < 3; ++i) {
            min[i] = box.b[0][i];
            max[i] = box.b[1][i];
        }
    }
    AABB aabb() const { return {{{{min[0], min[1], min[2], 0.f}}, {{max[0], max[1], max[2], 0.f}}}}; }
    void setSplitNode(uint32_t child) {
        min_data = child;
        max_data = 0;
    }
    void setLeafNode(uint32_t start, uint32_t num) {
        min_data = start;
        max_data = num;
    }
};
static_assert(sizeof(BvhNode) == 32, "reference size_test.zig:44");

}  // namespace zyg
