// ORACLE — test infrastructure only (see zmath.hpp). Scene-side value types of the reference restated
// over the flattened arrays of include/zygpu_scene.h:
//   ComposedTransformation   src/core/scene/composed_transformation.zig
//   Frame                    src/base/math/frame.zig
//   safe dots                src/base/math/safe.zig
//   ray offset               src/core/scene/ray_offset.zig
//   sampling                 src/base/math/sampling.zig
//   interpolated LUTs        src/base/math/interpolated_function.zig:229-300
#pragma once

#include "../include/zygpu_scene.h"
#include "zmath.hpp"

namespace zo {

constexpr float kPi    = 3.14159265358979323846f;
constexpr float kPiInv = 0.318309886183790671538f;

constexpr float RayMaxT = 2.14748313e+09f;  // ray_offset.zig:5

inline Vec4f load4(const float* p) { return {{p[0], p[1], p[2], p[3]}}; }

struct Trafo {  // composed_transformation.zig
    Vec4f r[3];
    Vec4f position;

    static Trafo load(const ZygpuTrafo& t) { return {{load4(t.r[0]), load4(t.r[1]), load4(t.r[2])}, load4(t.position)}; }

    float scaleX() const { return r[0][3]; }
    float scaleY() const { return r[1][3]; }
    float scaleZ() const { return r[2][3]; }
    Vec4f scale() const { return {{r[0][3], r[1][3], r[2][3], 1.f}}; }  // :43-45

    static Vec4f transformVector(const Vec4f rows[3], Vec4f v) {  // matrix3x3.zig:118-131
        Vec4f result = splat(v[0]) * rows[0];
        result       = mulAdd(splat(v[1]), rows[1], result);
        return mulAdd(splat(v[2]), rows[2], result);
    }
    static Vec4f transformVectorTransposed(const Vec4f rows[3], Vec4f v) {  // matrix3x3.zig:133-144
        const Vec4f x = v * rows[0];
        const Vec4f y = v * rows[1];
        const Vec4f z = v * rows[2];
        return {{x[0] + x[1] + x[2], y[0] + y[1] + y[2], z[0] + z[1] + z[2], 0.f}};
    }

    Vec4f objectToWorldVector(Vec4f v) const {  // :66-86
        const Vec4f s    = scale();
        const Vec4f a[3] = {r[0] * splat(s[0]), r[1] * splat(s[1]), r[2] * splat(s[2])};
        return transformVector(a, v);
    }
    Vec4f objectToWorldPoint(Vec4f p) const { return objectToWorldVector(p) + position; }  // :88-90
    Vec4f objectToWorldNormal(Vec4f n) const { return transformVector(r, n); }             // :96-98
    Vec4f frameToWorldPoint(Vec4f p) const { return objectToWorldNormal(p) + position; }   // :92-94
    Vec4f worldToObjectVector(Vec4f v) const {                                             // :100-105
        const Vec4f o = transformVectorTransposed(r, v);
        return o / scale();
    }
    Vec4f worldToObjectPoint(Vec4f p) const { return worldToObjectVector(p - position); }         // :107-109
    Vec4f worldToObjectNormal(Vec4f n) const { return transformVectorTransposed(r, n); }          // :115-117
    Vec4f worldToFramePoint(Vec4f p) const { return worldToObjectNormal(p - position); }          // :111-113
    Ray   worldToObjectRay(const Ray& ray) const {                                                // :119-126
        return Ray::init(worldToObjectPoint(ray.origin), worldToObjectVector(ray.direction), ray.min_t, ray.max_t);
    }
};

namespace safe {  // safe.zig
constexpr float DotMin = 0.00001f;
inline float    absDotC(Vec4f a, Vec4f b, bool c) {
    const float d = dot3(a, b);
    return c ? std::fabs(d) : d;
}
inline float clamp(float x) { return zo::clamp(x, DotMin, 1.f); }
inline float clampAbs(float x) { return zo::clamp(std::fabs(x), DotMin, 1.f); }
inline float clampDot(Vec4f a, Vec4f b) { return zo::clamp(dot3(a, b), DotMin, 1.f); }
inline float clampAbsDot(Vec4f a, Vec4f b) { return zo::clamp(std::fabs(dot3(a, b)), DotMin, 1.f); }
}  // namespace safe

inline float saturate(float x) { return clamp(x, 0.f, 1.f); }  // math.zig:122-124
inline float pow2(float x) { return x * x; }
inline float pow5(float x) {  // math.zig:149-153
    const float x2 = x * x;
    const float x4 = x2 * x2;
    return x4 * x;
}

struct Frame {  // frame.zig
    Vec4f x, y, z;

    static Frame init(Vec4f n) {
        Frame f;
        orthonormalBasis3(n, f.x, f.y);
        f.z = n;
        return f;
    }
    Vec4f frameToWorld(Vec4f v) const {  // :25-39
        Vec4f result = splat(v[0]) * x;
        result       = mulAdd(splat(v[1]), y, result);
        return mulAdd(splat(v[2]), z, result);
    }
    Vec4f worldToFrame(Vec4f v) const {  // :41-52
        const Vec4f t = v * x;
        const Vec4f b = v * y;
        const Vec4f n = v * z;
        return {{t[0] + t[1] + t[2], b[0] + b[1] + b[2], n[0] + n[1] + n[2], 0.f}};
    }
    Frame swapped(bool same_side) const { return same_side ? *this : Frame{x, y, -z}; }  // :15-21
    float nDot(Vec4f v) const { return dot3(z, v); }
    float clampNdot(Vec4f v) const { return safe::clampDot(z, v); }
    float clampAbsNdot(Vec4f v) const { return safe::clampAbsDot(z, v); }
};

// ray_offset.zig:14-27
inline Vec4f offsetRay(Vec4f p, Vec4f n) {
    const float origin      = 1.f / 32.f;
    const float float_scale = 1.f / 65536.f;
    const float int_scale   = 256.f;

    Vec4f r;
    for (int i = 0; i < 4; ++i) {
        const int32_t of_i = int32_t(int_scale * n[i]);
        int32_t       p_ii;
        const float   pi = p[i];
        std::memcpy(&p_ii, &pi, 4);
        const int32_t in = int32_t(uint32_t(p_ii) - uint32_t(of_i));
        const int32_t ip = int32_t(uint32_t(p_ii) + uint32_t(of_i));
        float         p_in, p_ip;
        std::memcpy(&p_in, &in, 4);
        std::memcpy(&p_ip, &ip, 4);
        const float p_i = pi < 0.f ? p_in : p_ip;
        const float mad = std::fmaf(float_scale, n[i], pi);
        r[i]            = std::fabs(pi) < origin ? mad : p_i;
    }
    return {{r[0], r[1], r[2], 0.f}};
}

// sampling.zig:8-32
inline void diskConcentric(const float uv[2], float out[2]) {
    const float s0 = (uv[0] * 2.f) - 1.f;
    const float s1 = (uv[1] * 2.f) - 1.f;
    if (0.f == s0 && 0.f == s1) {
        out[0] = out[1] = 0.f;
        return;
    }
    float r, theta;
    if (std::fabs(s0) > std::fabs(s1)) {
        r     = s0;
        theta = (kPi / 4.f) * (s1 / s0);
    } else {
        r     = s1;
        theta = (kPi / 2.f) - (kPi / 4.f) * (s0 / s1);
    }
    const float sin_theta = std::sin(theta);
    const float cos_theta = std::cos(theta);
    out[0]                = cos_theta * r;
    out[1]                = sin_theta * r;
}

// sampling.zig:50-55
inline Vec4f hemisphereCosine(const float uv[2]) {
    float xy[2];
    diskConcentric(uv, xy);
    const float z = std::sqrt(max(0.f, 1.f - xy[0] * xy[0] - xy[1] * xy[1]));
    return {{xy[0], xy[1], z, 0.f}};
}

inline float bilinear(const float c[4], float s, float t) {  // math.zig:172-179
    const float _s = 1.f - s;
    const float _t = 1.f - t;
    return _t * (_s * c[0] + s * c[1]) + t * (_s * c[2] + s * c[3]);
}

// InterpolatedFunction1DN.eval with fromArray (range_end 1, inverse_interval N-1), interpolated_function.zig:131-143
inline float lut1(const float* s, int N, float x) {
    const float    cx     = min(x, 1.f);
    const float    o      = cx * float(N - 1);
    const uint32_t offset = uint32_t(o);
    const float    t      = o - float(offset);
    return lerp(s[offset], s[std::min(offset + 1, uint32_t(N - 1))], t);
}

// InterpolatedFunction2DN.eval, :162-184
inline float lut2(const float* s, int X, int Y, float x, float y) {
    const float   mx   = min(x, 1.f);
    const float   my   = min(y, 1.f);
    const float   o0   = mx * float(X - 1);
    const float   o1   = my * float(Y - 1);
    const int32_t off0 = int32_t(o0);
    const int32_t off1 = int32_t(o1);
    const float   t0   = o0 - float(off0);
    const float   t1   = o1 - float(off1);
    const int32_t col1 = std::min(off0 + 1, X - 1);
    const int32_t row0 = off1 * X;
    const int32_t row1 = std::min(off1 + 1, Y - 1) * X;
    const float   c[4] = {s[off0 + row0], s[col1 + row0], s[off0 + row1], s[col1 + row1]};
    return bilinear(c, t0, t1);
}

// InterpolatedFunction3DN.eval, :203-249
inline float lut3(const float* s, int X, int Y, int Z, float x, float y, float z) {
    const float   o0   = min(x, 1.f) * float(X - 1);
    const float   o1   = min(y, 1.f) * float(Y - 1);
    const float   o2   = min(z, 1.f) * float(Z - 1);
    const int32_t off0 = int32_t(o0);
    const int32_t off1 = int32_t(o1);
    const int32_t off2 = int32_t(o2);
    const float   t0   = o0 - float(off0);
    const float   t1   = o1 - float(off1);
    const float   t2   = o2 - float(off2);

    const int32_t col1   = std::min(off0 + 1, X - 1);
    const int32_t row0   = off1 * X;
    const int32_t row1   = std::min(off1 + 1, Y - 1) * X;
    const int32_t area   = X * Y;
    const int32_t slice0 = off2 * area;
    const int32_t slice1 = std::min(off2 + 1, Z - 1) * area;

    const float ca[4] = {s[off0 + row0 + slice0], s[col1 + row0 + slice0], s[off0 + row1 + slice0], s[col1 + row1 + slice0]};
    const float cb[4] = {s[off0 + row0 + slice1], s[col1 + row0 + slice1], s[off0 + row1 + slice1], s[col1 + row1 + slice1]};

    const float c0 = bilinear(ca, t0, t1);
    const float c1 = bilinear(cb, t0, t1);
    return lerp(c0, c1, t2);
}

// The five tables of ggx_integral.zig in the order of ZygpuScene.ggx_luts.
struct GgxLuts {
    const float* E_m;      // 32 x 32     (n_dot, alpha)
    const float* E_m_avg;  // 32          (alpha)
    const float* E;        // 16^3        (n_dot, alpha, f0)
    const float* E_avg;    // 16 x 16     (alpha, f0)
    const float* E_s;      // 16^3

    explicit GgxLuts(const float* base)
        : E_m(base), E_m_avg(base + 1024), E(base + 1056), E_avg(base + 1056 + 4096), E_s(base + 1056 + 4096 + 256) {}

    float eM(float n_dot, float alpha) const { return lut2(E_m, 32, 32, n_dot, alpha); }
    float eMAvg(float alpha) const { return lut1(E_m_avg, 32, alpha); }
    float e(float n_dot, float alpha, float f0) const { return lut3(E, 16, 16, 16, n_dot, alpha, f0); }
    float eAvg(float alpha, float f0) const { return lut2(E_avg, 16, 16, alpha, f0); }
    float eS(float n_dot, float alpha, float f0) const { return lut3(E_s, 16, 16, 16, n_dot, alpha, f0); }
};

}  // namespace zo
