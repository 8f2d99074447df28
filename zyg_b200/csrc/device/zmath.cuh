// Strict-fp32 device math of the render stages: the lane-wise Vec4f arithmetic of the reference's
// base/math package on 3-component values (the 4th lane never reaches a result on this path).
//   vector4.zig, util.zig, frame.zig, safe.zig, sampling.zig, interpolated_function.zig,
//   composed_transformation.zig, ray_offset.zig
// Built with -fmad=false: a*b+c stays two roundings unless written as fma3 / __fmaf_rn, which marks the
// places where the Zig source says @mulAdd.
#pragma once

#include "trace_device.cuh"

namespace zygpu {

constexpr float kPi      = 3.14159265358979323846f;
constexpr float kPiInv   = 0.318309886183790671538f;
constexpr float kRayMaxT = 2.14748313e+09f;  // ray_offset.zig:5
constexpr float kDotMin  = 0.00001f;         // safe.zig:7

__device__ __forceinline__ V3 v3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ V3 splat3(float s) { return {s, s, s}; }
__device__ __forceinline__ V3 add3(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 mul3(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 div3(V3 a, V3 b) { return {__fdiv_rn(a.x, b.x), __fdiv_rn(a.y, b.y), __fdiv_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 scale3(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 divs3(V3 a, float s) { return {__fdiv_rn(a.x, s), __fdiv_rn(a.y, s), __fdiv_rn(a.z, s)}; }
__device__ __forceinline__ V3 neg3(V3 a) { return {-a.x, -a.y, -a.z}; }
// @mulAdd(Vec4f, a, b, c)
__device__ __forceinline__ V3 fma3(V3 a, V3 b, V3 c) { return {__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y), __fmaf_rn(a.z, b.z, c.z)}; }
__device__ __forceinline__ V3 fmas3(float a, V3 b, V3 c) { return {__fmaf_rn(a, b.x, c.x), __fmaf_rn(a, b.y, c.y), __fmaf_rn(a, b.z, c.z)}; }

__device__ __forceinline__ float length3(V3 a) { return __fsqrt_rn(dot3(a, a)); }
__device__ __forceinline__ float squaredLength3(V3 a) { return dot3(a, a); }
__device__ __forceinline__ V3    normalize3(V3 a) { return divs3(a, length3(a)); }  // vector4.zig:58-60
__device__ __forceinline__ float zclamp(float x, float mi, float ma) { return zmin(zmax(x, mi), ma); }
__device__ __forceinline__ float hmax3(V3 v) { return zmax(v.x, zmax(v.y, v.z)); }
__device__ __forceinline__ float zlerp(float a, float b, float t) {  // util.zig:3-8
    const float u = 1.f - t;
    return __fmaf_rn(u, a, t * b);
}
__device__ __forceinline__ V3 lerp3(V3 a, V3 b, V3 t) {
    return {zlerp(a.x, b.x, t.x), zlerp(a.y, b.y, t.y), zlerp(a.z, b.z, t.z)};
}
__device__ __forceinline__ float saturate(float x) { return zclamp(x, 0.f, 1.f); }
__device__ __forceinline__ float pow5(float x) {
    const float x2 = x * x;
    const float x4 = x2 * x2;
    return x4 * x;
}

// safe.zig
__device__ __forceinline__ float clampDot(V3 a, V3 b) { return zclamp(dot3(a, b), kDotMin, 1.f); }
__device__ __forceinline__ float clampAbsDot(V3 a, V3 b) { return zclamp(fabsf(dot3(a, b)), kDotMin, 1.f); }
__device__ __forceinline__ float safeClamp(float x) { return zclamp(x, kDotMin, 1.f); }

// vector4.zig:98-110
__device__ __forceinline__ void orthonormalBasis3(V3 n, V3& t, V3& b) {
    const float sign = copysignf(1.f, n.z);
    const float c    = __fdiv_rn(-1.f, sign + n.z);
    const float d    = n.x * n.y * c;
    t                = {1.f + sign * n.x * n.x * c, sign * d, -sign * n.x};
    b                = {d, sign + n.y * n.y * c, -n.y};
}

struct FrameD {  // frame.zig
    V3 x, y, z;

    __device__ __forceinline__ V3 frameToWorld(V3 v) const {
        V3 r = scale3(v.x, x);
        r    = fmas3(v.y, y, r);
        return fmas3(v.z, z, r);
    }
    __device__ __forceinline__ V3 worldToFrame(V3 v) const {
        const V3 t = mul3(v, x), b = mul3(v, y), n = mul3(v, z);
        return {t.x + t.y + t.z, b.x + b.y + b.z, n.x + n.y + n.z};
    }
    __device__ __forceinline__ float clampNdot(V3 v) const { return clampDot(z, v); }
    __device__ __forceinline__ float clampAbsNdot(V3 v) const { return clampAbsDot(z, v); }
};

// composed_transformation.zig over the 64-byte record of ZygpuTrafo
struct TrafoD {
    V3 r0, r1, r2;  // rotation rows
    V3 scale;
    V3 position;

    __device__ __forceinline__ V3 transformVector(V3 v) const {  // matrix3x3.zig:118-131
        V3 r = scale3(v.x, r0);
        r    = fmas3(v.y, r1, r);
        return fmas3(v.z, r2, r);
    }
    __device__ __forceinline__ V3 transformVectorTransposed(V3 v) const {  // :133-144
        const V3 x = mul3(v, r0), y = mul3(v, r1), z = mul3(v, r2);
        return {x.x + x.y + x.z, y.x + y.y + y.z, z.x + z.y + z.z};
    }
    __device__ __forceinline__ V3 objectToWorldVector(V3 v) const {  // :66-86
        const V3 a = scale3(scale.x, r0), b = scale3(scale.y, r1), c = scale3(scale.z, r2);
        V3       r = scale3(v.x, a);
        r          = fmas3(v.y, b, r);
        return fmas3(v.z, c, r);
    }
    __device__ __forceinline__ V3 objectToWorldPoint(V3 p) const { return add3(objectToWorldVector(p), position); }
    __device__ __forceinline__ V3 objectToWorldNormal(V3 n) const { return transformVector(n); }
    __device__ __forceinline__ V3 frameToWorldPoint(V3 p) const { return add3(transformVector(p), position); }
    __device__ __forceinline__ V3 worldToObjectVector(V3 v) const { return div3(transformVectorTransposed(v), scale); }
    __device__ __forceinline__ V3 worldToObjectPoint(V3 p) const { return worldToObjectVector(sub3(p, position)); }
    __device__ __forceinline__ V3 worldToObjectNormal(V3 n) const { return transformVectorTransposed(n); }
    __device__ __forceinline__ V3 worldToFramePoint(V3 p) const { return transformVectorTransposed(sub3(p, position)); }
};

__device__ __forceinline__ TrafoD loadTrafo(const float4* trafos, uint32_t i) {
    const float4 a = __ldg(trafos + 4 * size_t(i) + 0);
    const float4 b = __ldg(trafos + 4 * size_t(i) + 1);
    const float4 c = __ldg(trafos + 4 * size_t(i) + 2);
    const float4 p = __ldg(trafos + 4 * size_t(i) + 3);
    return {{a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}, {a.w, b.w, c.w}, {p.x, p.y, p.z}};
}

// math.Ray.init, ray.zig:11-20
__device__ __forceinline__ RayT makeRay(V3 o, V3 d, float tmin, float tmax) {
    RayT r;
    r.o     = o;
    r.d     = d;
    r.inv_d = {__fdiv_rn(1.f, d.x), __fdiv_rn(1.f, d.y), __fdiv_rn(1.f, d.z)};
    r.tmin  = tmin;
    r.tmax  = tmax;
    return r;
}
__device__ __forceinline__ V3 rayPoint(const RayT& r, float t) { return fmas3(t, r.d, r.o); }  // ray.zig:27-29

// ray_offset.zig:14-27
__device__ __forceinline__ float offsetRayLane(float p, float n) {
    const float origin      = 1.f / 32.f;
    const float float_scale = 1.f / 65536.f;
    const float int_scale   = 256.f;

    const int   of_i = __float2int_rz(int_scale * n);
    const int   p_ii = __float_as_int(p);
    const float p_in = __int_as_float(int(uint32_t(p_ii) - uint32_t(of_i)));
    const float p_ip = __int_as_float(int(uint32_t(p_ii) + uint32_t(of_i)));
    const float p_i  = p < 0.f ? p_in : p_ip;
    const float mad  = __fmaf_rn(float_scale, n, p);
    return fabsf(p) < origin ? mad : p_i;
}
__device__ __forceinline__ V3 offsetRay(V3 p, V3 n) {
    return {offsetRayLane(p.x, n.x), offsetRayLane(p.y, n.y), offsetRayLane(p.z, n.z)};
}

// sampling.zig:8-32
__device__ __forceinline__ void diskConcentric(float u0, float u1, float& ox, float& oy) {
    const float s0 = (u0 * 2.f) - 1.f;
    const float s1 = (u1 * 2.f) - 1.f;
    if (0.f == s0 && 0.f == s1) {
        ox = oy = 0.f;
        return;
    }
    float r, theta;
    if (fabsf(s0) > fabsf(s1)) {
        r     = s0;
        theta = (kPi / 4.f) * __fdiv_rn(s1, s0);
    } else {
        r     = s1;
        theta = (kPi / 2.f) - (kPi / 4.f) * __fdiv_rn(s0, s1);
    }
    float sin_theta, cos_theta;
    sincosf(theta, &sin_theta, &cos_theta);
    ox = cos_theta * r;
    oy = sin_theta * r;
}

// sampling.zig:50-55
__device__ __forceinline__ V3 hemisphereCosine(float u0, float u1) {
    float x, y;
    diskConcentric(u0, u1, x, y);
    const float z = __fsqrt_rn(zmax(0.f, 1.f - x * x - y * y));
    return {x, y, z};
}

__device__ __forceinline__ float bilinear(float c0, float c1, float c2, float c3, float s, float t) {  // math.zig:172-179
    const float _s = 1.f - s;
    const float _t = 1.f - t;
    return _t * (_s * c0 + s * c1) + t * (_s * c2 + s * c3);
}

// InterpolatedFunction{1,2,3}DN.eval with fromArray, interpolated_function.zig:131-143, 162-184, 203-249
__device__ __forceinline__ float lut1(const float* s, int N, float x) {
    const float    o      = zmin(x, 1.f) * float(N - 1);
    const uint32_t offset = uint32_t(o);
    const float    t      = o - float(offset);
    return zlerp(__ldg(s + offset), __ldg(s + min(offset + 1, uint32_t(N - 1))), t);
}
__device__ __forceinline__ float lut2(const float* s, int X, int Y, float x, float y) {
    const float o0   = zmin(x, 1.f) * float(X - 1);
    const float o1   = zmin(y, 1.f) * float(Y - 1);
    const int   off0 = int(o0), off1 = int(o1);
    const float t0 = o0 - float(off0), t1 = o1 - float(off1);
    const int   col1 = min(off0 + 1, X - 1);
    const int   row0 = off1 * X;
    const int   row1 = min(off1 + 1, Y - 1) * X;
    return bilinear(__ldg(s + off0 + row0), __ldg(s + col1 + row0), __ldg(s + off0 + row1), __ldg(s + col1 + row1), t0, t1);
}
__device__ __forceinline__ float lut3(const float* s, int X, int Y, int Z, float x, float y, float z) {
    const float o0 = zmin(x, 1.f) * float(X - 1);
    const float o1 = zmin(y, 1.f) * float(Y - 1);
    const float o2 = zmin(z, 1.f) * float(Z - 1);
    const int   off0 = int(o0), off1 = int(o1), off2 = int(o2);
    const float t0 = o0 - float(off0), t1 = o1 - float(off1), t2 = o2 - float(off2);
    const int   col1   = min(off0 + 1, X - 1);
    const int   row0   = off1 * X;
    const int   row1   = min(off1 + 1, Y - 1) * X;
    const int   area   = X * Y;
    const int   slice0 = off2 * area;
    const int   slice1 = min(off2 + 1, Z - 1) * area;
    const float c0 = bilinear(__ldg(s + off0 + row0 + slice0), __ldg(s + col1 + row0 + slice0), __ldg(s + off0 + row1 + slice0),
                              __ldg(s + col1 + row1 + slice0), t0, t1);
    const float c1 = bilinear(__ldg(s + off0 + row0 + slice1), __ldg(s + col1 + row0 + slice1), __ldg(s + off0 + row1 + slice1),
                              __ldg(s + col1 + row1 + slice1), t0, t1);
    return zlerp(c0, c1, t2);
}

}  // namespace zygpu
