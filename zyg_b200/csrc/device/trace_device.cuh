// Device-side building blocks shared by the traversal kernels (trace.cu) and the wavefront render
// stages (render.cu): the reference's strict-fp32 ray / node / triangle arithmetic and the 8-wide
// quantised node test. Translation units including this header are compiled with -fmad=false; FMAs
// appear only where the reference writes @mulAdd.
#pragma once

#include "trace.cuh"

#include <cfloat>

namespace zygpu {

// ---------------------------------------------------------------------------------------------
// Reference arithmetic (strict fp32; this file is built with -fmad=false)
// ---------------------------------------------------------------------------------------------

struct V3 {
    float x, y, z;
};

__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

// src/base/math/vector4.zig:36-39 : (x + y) + z
__device__ __forceinline__ float dot3(V3 a, V3 b) {
    const float x = a.x * b.x, y = a.y * b.y, z = a.z * b.z;
    return (x + y) + z;
}

// src/base/math/vector4.zig:73-92 : one FMA per lane
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return {__fmaf_rn(b.z, a.y, -(a.z * b.y)), __fmaf_rn(b.x, a.z, -(a.x * b.z)), __fmaf_rn(b.y, a.x, -(a.y * b.x))};
}

// src/base/math/util.zig:17-29 (x86 branch)
__device__ __forceinline__ float zmin(float x, float y) { return x < y ? x : y; }
__device__ __forceinline__ float zmax(float x, float y) { return y < x ? x : y; }

struct RayT {
    V3    o, d, inv_d;
    float tmin, tmax;
};

// src/core/scene/shape/triangle/triangle.zig:26-52 with e1/e2 hoisted (same fp32 subtractions).
__device__ __forceinline__ bool intersectTriangle(const RayT& ray, V3 a, V3 e1, V3 e2, float& ht, float& hu, float& hv) {
    const V3 tvec = sub3(ray.o, a);
    const V3 pvec = cross3(ray.d, e2);
    const V3 qvec = cross3(tvec, e1);

    const float e1_d_pv = dot3(e1, pvec);
    const float tv_d_pv = dot3(tvec, pvec);
    const float di_d_qv = dot3(ray.d, qvec);
    const float e2_d_qv = dot3(e2, qvec);

    const float inv_det = __fdiv_rn(1.f, e1_d_pv);

    const float u     = tv_d_pv * inv_det;
    const float v     = di_d_qv * inv_det;
    const float hit_t = e2_d_qv * inv_det;

    const float uv = u + v;

    if (u >= 0.f && 1.f >= u && v >= 0.f && 1.f >= uv && hit_t >= ray.tmin && ray.tmax >= hit_t) {
        ht = hit_t;
        hu = u;
        hv = v;
        return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------------
// Order-exact binary traversal
// ---------------------------------------------------------------------------------------------

// src/core/scene/bvh/node.zig:73-87
__device__ __forceinline__ float intersectNode(const float4 nmin, const float4 nmax, const RayT& ray) {
    const float lx = (nmin.x - ray.o.x) * ray.inv_d.x, ly = (nmin.y - ray.o.y) * ray.inv_d.y,
                lz = (nmin.z - ray.o.z) * ray.inv_d.z;
    const float ux = (nmax.x - ray.o.x) * ray.inv_d.x, uy = (nmax.y - ray.o.y) * ray.inv_d.y,
                uz = (nmax.z - ray.o.z) * ray.inv_d.z;

    const float t0x = zmin(lx, ux), t0y = zmin(ly, uy), t0z = zmin(lz, uz);
    const float t1x = zmax(lx, ux), t1y = zmax(ly, uy), t1z = zmax(lz, uz);

    // hmax4 / hmin4, vector4.zig:163-175
    const float tboxmin = zmax(t0x, zmax(t0y, zmax(t0z, ray.tmin)));
    const float tboxmax = zmin(t1x, zmin(t1y, zmin(t1z, ray.tmax)));

    return tboxmin <= tboxmax ? tboxmin : FLT_MAX;
}

constexpr uint32_t kBinaryStack = 127;  // src/core/scene/bvh/node_stack.zig:2
constexpr uint32_t kEnd         = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// Wide traversal
// ---------------------------------------------------------------------------------------------

constexpr uint32_t kWideStack = 48;

struct WideRay {
    RayT     ray;
    V3       cid;  // clamped reciprocal direction for the quantised (culling-only) slab test
    uint32_t octinv;
    bool     px, py, pz;
};

__device__ __forceinline__ void setupWideRay(WideRay& w) {
    // octant: slot bit 2 <-> x, bit 1 <-> y, bit 0 <-> z; children further along the ray get lower priority
    w.px     = !signbit(w.ray.d.x);
    w.py     = !signbit(w.ray.d.y);
    w.pz     = !signbit(w.ray.d.z);
    w.octinv = (w.px ? 4u : 0u) | (w.py ? 2u : 0u) | (w.pz ? 1u : 0u);

    // A zero direction component would turn q * inf + (p - o) * inf into NaNs that switch the axis off,
    // and an axis-parallel ray would walk every node in its slab. With +-2^80 a ray parallel to a slab
    // gets (-huge, +huge) when inside and an empty interval when outside, as it should.
    constexpr float kBig = 1.2089258e24f;  // 2^80
    w.cid = {fminf(fmaxf(w.ray.inv_d.x, -kBig), kBig), fminf(fmaxf(w.ray.inv_d.y, -kBig), kBig),
             fminf(fmaxf(w.ray.inv_d.z, -kBig), kBig)};
}

// 256-bit global loads (LDG.E.256 on sm_100): the L1 handles one request per distinct 32-byte sector either way, so a record fetched
// with 32-byte loads costs half the requests of 16-byte loads. Divergent node / triangle fetches are bound by that request rate
// (profiles/r02: l1tex throughput 60 %+ with prefetches making it worse), not by bytes.
struct F8 {
    float4 lo, hi;
};
__device__ __forceinline__ F8 ldg256(const float4* p) {  // p must be 32-byte aligned
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
        : "l"(p));
    return r;
}

// The same load with an L2 eviction priority (LDG.E.ELL2 / .EFL2): Hint 1 = evict last (the node arrays, reused by every ray), 2 = evict
// first (triangle records of scenes that do not fit the 126 MB L2). ZYGPU_NODE_L2 / ZYGPU_TRI_L2 choose; measured in profiles/r02_sweeps.md.
#ifndef ZYGPU_NODE_L2
#define ZYGPU_NODE_L2 0
#endif
#ifndef ZYGPU_TRI_L2
#define ZYGPU_TRI_L2 0
#endif
template <int Hint>
__device__ __forceinline__ F8 ldg256Hint(const float4* p) {
    F8 r;
    if (1 == Hint) {
        asm("ld.global.nc.L1::evict_normal.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
            : "l"(p));
    } else if (2 == Hint) {
        asm("ld.global.nc.L1::evict_normal.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
            : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
            : "l"(p));
    } else {
        r = ldg256(p);
    }
    return r;
}

constexpr uint32_t kWideNodeWords = 6;  // float4 per node: 80 bytes of node + 16 bytes of padding = three 32-byte loads

struct WideNodeRegs {
    float4 n0, n1, n2, n3, n4;
};
__device__ __forceinline__ WideNodeRegs loadWideNode(const float4* nodes, uint32_t index) {
    const float4* np = nodes + kWideNodeWords * size_t(index);
    const F8      a = ldg256Hint<ZYGPU_NODE_L2>(np), b = ldg256Hint<ZYGPU_NODE_L2>(np + 2);
    const float4  c = __ldg(np + 4);
    return {a.lo, a.hi, b.lo, b.hi, c};
}

// byte j of `word` -> 32768 + byte as a float, one PRMT and no int->float conversion (the XU pipe
// was the busiest unit with I2F: profiles/r01_traceWide_a.md): 0x47000000 is 32768.0f and the byte
// lands in mantissa bits 8..15, i.e. at weight 1.
template <int J>
__device__ __forceinline__ float biasedByte(uint32_t word) {
    return __uint_as_float(__byte_perm(word, 0x47000000u, 0x7604u | (J << 4)));
}

// Tests the eight quantised child boxes of one node. Returns the hit mask: bits 24..31 inner children
// in traversal priority (slot ^ octinv), bits 0..23 one bit per triangle of the hit leaf slots.
__device__ __forceinline__ uint32_t testWideNode(const WideRay& w, float tmin_ray, float tmax_ray, const float4 n0,
                                                 const float4 n1, const float4 n2, const float4 n3, const float4 n4) {
    const uint32_t ew = __float_as_uint(n0.w);

    const float idx = __uint_as_float((ew & 0xffu) << 23) * w.cid.x;
    const float idy = __uint_as_float(((ew >> 8) & 0xffu) << 23) * w.cid.y;
    const float idz = __uint_as_float(((ew >> 16) & 0xffu) << 23) * w.cid.z;

    const float orx = (n0.x - w.ray.o.x) * w.cid.x;
    const float ory = (n0.y - w.ray.o.y) * w.cid.y;
    const float orz = (n0.z - w.ray.o.z) * w.cid.z;

    // t = (32768 + q) * idir + (orig - 32768 * idir). Conservative slack: this associates differently
    // from the reference's (min - o) * inv_d and the biased origin loses ~2^-9 of a cell; widen every
    // interval by 2^-21 |orig| + 2^-5 |idir| (1/32 of a quantisation cell) so no child the exact test
    // would accept is ever dropped.
    constexpr float kSlackO = 4.76837158e-7f;  // 2^-21
    constexpr float kSlackI = 0.03125f;        // 2^-5
    const float     pdx = fmaf(kSlackO, fabsf(orx), kSlackI * fabsf(idx));
    const float     pdy = fmaf(kSlackO, fabsf(ory), kSlackI * fabsf(idy));
    const float     pdz = fmaf(kSlackO, fabsf(orz), kSlackI * fabsf(idz));

    const float bx = fmaf(-32768.f, idx, orx), by = fmaf(-32768.f, idy, ory), bz = fmaf(-32768.f, idz, orz);
    const float olx = bx - pdx, ohx = bx + pdx;
    const float oly = by - pdy, ohy = by + pdy;
    const float olz = bz - pdz, ohz = bz + pdz;

    // quantised planes: near = lo when the ray travels in +axis, else hi
    const uint32_t qlox[2] = {__float_as_uint(n2.x), __float_as_uint(n2.y)};
    const uint32_t qloy[2] = {__float_as_uint(n2.z), __float_as_uint(n2.w)};
    const uint32_t qloz[2] = {__float_as_uint(n3.x), __float_as_uint(n3.y)};
    const uint32_t qhix[2] = {__float_as_uint(n3.z), __float_as_uint(n3.w)};
    const uint32_t qhiy[2] = {__float_as_uint(n4.x), __float_as_uint(n4.y)};
    const uint32_t qhiz[2] = {__float_as_uint(n4.z), __float_as_uint(n4.w)};
    const uint32_t meta[2] = {__float_as_uint(n1.z), __float_as_uint(n1.w)};

    uint32_t hitmask = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t nx = w.px ? qlox[h] : qhix[h], fx = w.px ? qhix[h] : qlox[h];
        const uint32_t ny = w.py ? qloy[h] : qhiy[h], fy = w.py ? qhiy[h] : qloy[h];
        const uint32_t nz = w.pz ? qloz[h] : qhiz[h], fz = w.pz ? qhiz[h] : qloz[h];

#define ZYGPU_CHILD(J)                                                                          \
    {                                                                                           \
        const float tminx = fmaf(biasedByte<J>(nx), idx, olx);                                  \
        const float tminy = fmaf(biasedByte<J>(ny), idy, oly);                                  \
        const float tminz = fmaf(biasedByte<J>(nz), idz, olz);                                  \
        const float tmaxx = fmaf(biasedByte<J>(fx), idx, ohx);                                  \
        const float tmaxy = fmaf(biasedByte<J>(fy), idy, ohy);                                  \
        const float tmaxz = fmaf(biasedByte<J>(fz), idz, ohz);                                  \
        const float tmin  = fmaxf(fmaxf(tminx, tminy), fmaxf(tminz, tmin_ray));                 \
        const float tmax  = fminf(fminf(tmaxx, tmaxy), fminf(tmaxz, tmax_ray));                 \
        if (tmin <= tmax) {                                                                     \
            const uint32_t m          = (meta[h] >> (8 * J)) & 0xffu;                           \
            const uint32_t child_bits = m >> 5;                                                 \
            uint32_t       bit_index  = m & 31u;                                                \
            if (bit_index >= 24u) bit_index ^= w.octinv;                                        \
            hitmask |= child_bits << bit_index;                                                 \
        }                                                                                       \
    }
        ZYGPU_CHILD(0)
        ZYGPU_CHILD(1)
        ZYGPU_CHILD(2)
        ZYGPU_CHILD(3)
#undef ZYGPU_CHILD
    }
    return hitmask;
}

// The reference's slab test (node.zig:73-87) against a box with the ray's max_t replaced by `gate_tmax`, then a conservative cull
// against the ray's current max_t. Why two limits: the reference's own test is not watertight against the shape / triangle
// test, so whether it passes can depend on max_t — and on the device the current max_t depends on the order in which the
// lock-step schedule found earlier hits. Gating with the max_t the ray STARTED with makes the decision independent of the
// schedule (equal-t ties included, see closerOrLater); the cull against the current max_t only drops boxes whose entry lies
// beyond it by more than the rounding of the test, where no hit could be accepted or tie.
// The limit of every cull against the ray's *current* max_t. How far max_t has come down when a box is met depends on the schedule, so
// such a cull must never drop a box that could still hold a hit at or below the final max_t. A slab test's entry and Moeller-Trumbore's t
// are rounded differently - the latter with errors relative to the distance between origin and triangle, not to t - so a hit can come out
// slightly in front of its own box: with a margin of 5e-7 two paths of the 66 M of a 4K frame flipped between two near-equal hits from
// run to run. 0.1 % and an absolute term in units of the origin's magnitude cost no measurable time.
__device__ __forceinline__ float originScale(const RayT& ray) { return fabsf(ray.o.x) + fabsf(ray.o.y) + fabsf(ray.o.z) + 1.f; }
__device__ __forceinline__ float cullLimit(const RayT& ray) { return fmaf(ray.tmax, 1.001f, 2e-5f * originScale(ray)); }

// `entry`: where the ray enters the box (its min_t when it starts inside)
__device__ __forceinline__ bool gateBox(const float4 bmin, const float4 bmax, const RayT& ray, float gate_tmax, float& entry) {
    RayT g = ray;
    g.tmax = gate_tmax;
    entry  = intersectNode(bmin, bmax, g);
    return FLT_MAX != entry && entry <= cullLimit(ray);
}
__device__ __forceinline__ bool gateBox(const float4 bmin, const float4 bmax, const RayT& ray, float gate_tmax) {
    float entry;
    return gateBox(bmin, bmax, ray, gate_tmax, entry);
}

// One gated triangle test against record `index`; true on an accepted hit (fills t, u, v, primitive). `gate_tmax`: the max_t
// the ray started with (closest hit), or its max_t (any hit, where it never changes).
__device__ __forceinline__ bool testWideTriangle(const MeshDevice& mesh, const RayT& ray, float gate_tmax, uint32_t index, float& t,
                                                 float& u, float& v, uint32_t& primitive) {
    const float4* tp = mesh.wide_tris + 4 * size_t(index);
    const F8      ta = ldg256Hint<ZYGPU_TRI_L2>(tp), tb = ldg256Hint<ZYGPU_TRI_L2>(tp + 2);
    const float4  t0 = ta.lo, t1 = ta.hi, t2 = tb.lo, t3 = tb.hi;

    // Gate with the reference's own (non-watertight) slab test on the reference leaf box: the
    // reference never tests a triangle whose leaf box it rejected.
    float entry;
    if (!gateBox(make_float4(t1.w, t2.w, t3.x, 0.f), make_float4(t3.y, t3.z, t3.w, 0.f), ray, gate_tmax, entry)) return false;
    primitive = __float_as_uint(t0.w);
    if (!intersectTriangle(ray, {t0.x, t0.y, t0.z}, {t1.x, t1.y, t1.z}, {t2.x, t2.y, t2.z}, t, u, v)) return false;
    // A hit in front of its own leaf box is an artefact of a degenerate triangle (a pole of a lat-long sphere met exactly: every term of
    // u, v and t rounds to 0 over a tiny determinant, "hit" at t = 0). The reference accepts it when its order happens to reach that leaf
    // before a nearer real hit has culled it; a parallel walk has no such order, so whether the leaf is reached would depend on the
    // schedule (two rays of the 66 M of a 4K frame of the instanced scene flipped from run to run). It is refused: no surface lies there.
    // The tolerance is half the margin of cullLimit, so a hit that passes could never have been culled in another order.
    return t >= fmaf(entry, 1.f - 5e-4f, -1e-5f * originScale(ray));
}

// One ray through one mesh, per-thread while-while loop over the wide layout (the body of the
// `traceWide` kernel): used by the render stages, where a ray reaches a mesh through the prop tree.
// Closest hit: shrinks w.ray.tmax and fills (ht, hu, hv, primitive); any hit: returns at the first hit.
template <bool AnyHit>
__device__ __forceinline__ bool traverseWide(const MeshDevice& mesh, WideRay& w, float& ht, float& hu, float& hv,
                                             uint32_t& primitive) {
    uint2    stack[kWideStack];
    uint32_t sp = 0;

    uint2 node_group = make_uint2(0u, 0x80000000u);
    uint2 tri_group  = make_uint2(0u, 0u);

    bool found = false;

    for (;;) {
        if (node_group.y > 0x00FFFFFFu) {
            const uint32_t hits  = node_group.y;
            const uint32_t gmask = hits & 0xffu;
            const uint32_t bit   = 31u - __clz(hits);
            node_group.y         = hits & ~(1u << bit);
            const uint32_t slot  = (bit - 24u) ^ w.octinv;
            const uint32_t rank  = __popc(gmask & ((1u << slot) - 1u));
            const uint32_t node_index = node_group.x + rank;
            if (node_group.y > 0x00FFFFFFu) stack[sp++] = node_group;

            const WideNodeRegs nd = loadWideNode(mesh.wide_nodes, node_index);
            const float4 n0 = nd.n0, n1 = nd.n1;

            const uint32_t hitmask = testWideNode(w, w.ray.tmin, cullLimit(w.ray), nd.n0, nd.n1, nd.n2, nd.n3, nd.n4);

            node_group.x = __float_as_uint(n1.x);
            node_group.y = (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
            tri_group.x  = __float_as_uint(n1.y);
            tri_group.y  = hitmask & 0x00FFFFFFu;
        } else {
            tri_group    = node_group;
            node_group.y = 0;
        }

        while (0 != tri_group.y) {
            const uint32_t bit = 31u - __clz(tri_group.y);
            tri_group.y &= ~(1u << bit);

            float    t, u, v;
            uint32_t prim;
            if (testWideTriangle(mesh, w.ray, w.ray.tmax, tri_group.x + bit, t, u, v, prim)) {
                if (AnyHit) return true;
                w.ray.tmax = t;
                ht         = t;
                hu         = u;
                hv         = v;
                primitive  = prim;
                found      = true;
            }
        }

        if (node_group.y <= 0x00FFFFFFu) {
            if (0 == sp) break;
            node_group = stack[--sp];
        }
    }
    return found;
}

}  // namespace zygpu
