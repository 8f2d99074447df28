// ORACLE — test infrastructure only. Nothing under zyg_b200/ may include, link or call this.
//
// CPU restatement of the reference's base/math package (strict fp32, FMA only where the Zig source
// writes @mulAdd). Build with -ffp-contract=off. PARITY UNPINNED for the geometry code: the
// reference ships no tests or golden vectors for it; the published PCG32 and
// Sobol/Owen constants are the only external pins (tests/test_oracle_pins.py).
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace zo {

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

// src/base/math/util.zig:17-29, x86 branch
inline float min(float x, float y) { return x < y ? x : y; }
inline float max(float x, float y) { return y < x ? x : y; }
inline float clamp(float x, float mi, float ma) { return min(max(x, mi), ma); }  // util.zig:31-33

// util.zig:3-15
inline float lerp(float a, float b, float t) {
    const float u = 1.f - t;
    return std::fmaf(u, a, t * b);
}
inline Vec4f lerp(Vec4f a, Vec4f b, Vec4f t) {
    const Vec4f u = splat(1.f) - t;
    return mulAdd(u, a, t * b);
}

// src/base/math/vector4.zig
inline float dot3(Vec4f a, Vec4f b) {  // :36-39
    const Vec4f ab = a * b;
    return ab[0] + ab[1] + ab[2];
}
inline float squaredLength3(Vec4f v) { return dot3(v, v); }
inline float length3(Vec4f v) { return std::sqrt(dot3(v, v)); }
inline float squaredDistance3(Vec4f a, Vec4f b) { return squaredLength3(a - b); }
inline float distance3(Vec4f a, Vec4f b) { return length3(a - b); }
inline Vec4f normalize3(Vec4f v) { return v / splat(length3(v)); }         // :58-60
inline Vec4f reciprocal3(Vec4f v) { return splat(1.f) / v; }               // :62-64
inline Vec4f shuffle1203(Vec4f a) { return {{a[1], a[2], a[0], a[3]}}; }
inline Vec4f cross3(Vec4f a, Vec4f b) {                                    // :73-92
    const Vec4f tmp0 = shuffle1203(b);
    const Vec4f tmp1 = shuffle1203(a);
    const Vec4f tmp2 = mulAdd(tmp0, a, -(tmp1 * b));
    return shuffle1203(tmp2);
}
inline Vec4f reflect3(Vec4f n, Vec4f v) { return splat(2.f * dot3(v, n)) * n - v; }  // :94-96
inline void  orthonormalBasis3(Vec4f n, Vec4f& t, Vec4f& b) {                        // :98-112
    const float sign = std::copysign(1.f, n[2]);
    const float c    = -1.f / (sign + n[2]);
    const float d    = n[0] * n[1] * c;
    t                = {{1.f + sign * n[0] * n[0] * c, sign * d, -sign * n[0], 0.f}};
    b                = {{d, sign + n[1] * n[1] * c, -n[1], 0.f}};
}
inline Vec4f tangent3(Vec4f n) {  // :114-120
    const float sign = std::copysign(1.f, n[2]);
    const float c    = -1.f / (sign + n[2]);
    const float d    = n[0] * n[1] * c;
    return {{1.f + sign * n[0] * n[0] * c, sign * d, -sign * n[0], 0.f}};
}
inline Vec4f gramSchmidt(Vec4f v, Vec4f w) { return mulAdd(splat(-dot3(v, w)), w, v); }  // :122-124
inline Vec4f min4(Vec4f a, Vec4f b) {                                                    // :126-136
    return {{min(a[0], b[0]), min(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])}};
}
inline Vec4f max4(Vec4f a, Vec4f b) {  // :138-148
    return {{max(a[0], b[0]), max(a[1], b[1]), max(a[2], b[2]), max(a[3], b[3])}};
}
inline Vec4f clamp4(Vec4f v, Vec4f mi, Vec4f ma) { return min4(max4(v, mi), ma); }
inline float hmin3(Vec4f v) { return min(v[0], min(v[1], v[2])); }                 // :154-156
inline float hmax3(Vec4f v) { return max(v[0], max(v[1], v[2])); }                 // :158-160
inline float hmin4(Vec4f v) { return min(v[0], min(v[1], min(v[2], v[3]))); }      // :167-172
inline float hmax4(Vec4f v) { return max(v[0], max(v[1], max(v[2], v[3]))); }      // :174-179
inline uint32_t indexMinComponent3(Vec4f v) {                                      // :181-187
    if (v[0] < v[1]) return v[0] < v[2] ? 0 : 2;
    return v[1] < v[2] ? 1 : 2;
}
inline uint32_t indexMaxComponent3(Vec4f v) {  // :189-195
    if (v[0] > v[1]) return v[0] > v[2] ? 0 : 2;
    return v[1] > v[2] ? 1 : 2;
}
inline float average3(Vec4f v) { return (v[0] + v[1] + v[2]) / 3.f; }
inline bool  allLessEqualZero3(Vec4f v) { return v[0] <= 0.f && v[1] <= 0.f && v[2] <= 0.f; }  // :226-228
inline bool  anyGreaterZero3(Vec4f v) { return v[0] > 0.f || v[1] > 0.f || v[2] > 0.f; }       // :230-232
inline bool  anyNaN3(Vec4f v) { return std::isnan(v[0]) || std::isnan(v[1]) || std::isnan(v[2]); }

// src/base/math/ray.zig
struct Ray {
    Vec4f origin, direction, inv_direction;
    float min_t, max_t;

    static Ray init(Vec4f origin, Vec4f direction, float min_t, float max_t) {  // :11-20
        Ray r;
        r.origin        = origin;
        r.direction     = direction;
        r.inv_direction = reciprocal3({{direction[0], direction[1], direction[2], 1.f}});
        r.min_t         = min_t;
        r.max_t         = max_t;
        return r;
    }
    Vec4f point(float t) const { return mulAdd(splat(t), direction, origin); }  // :27-29
};

// src/base/math/aabb.zig (query side)
struct AABB {
    Vec4f bounds[2];

    bool intersect(const Ray& ray) const {  // :46-60
        const Vec4f lower = (bounds[0] - ray.origin) * ray.inv_direction;
        const Vec4f upper = (bounds[1] - ray.origin) * ray.inv_direction;
        const Vec4f t0    = min4(lower, upper);
        const Vec4f t1    = max4(lower, upper);
        const Vec4f tmins = {{t0[0], t0[1], t0[2], ray.min_t}};
        const Vec4f tmaxs = {{t1[0], t1[1], t1[2], ray.max_t}};
        return hmax4(tmins) <= hmin4(tmaxs);
    }
};

}  // namespace zo
