#include "trace.cuh"

#include <algorithm>
#include <cfloat>
#include <cstdlib>

namespace zygpu {

namespace {

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

__device__ __forceinline__ RayT loadRay(const RayIn* rays, uint32_t i) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(rays) + 2 * size_t(i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(rays) + 2 * size_t(i) + 1);
    RayT         r;
    r.o     = {a.x, a.y, a.z};
    r.tmin  = a.w;
    r.d     = {b.x, b.y, b.z};
    r.tmax  = b.w;
    // src/base/math/ray.zig:11-20 : reciprocal3 is a true division (vector4.zig:62-64)
    r.inv_d = {__fdiv_rn(1.f, b.x), __fdiv_rn(1.f, b.y), __fdiv_rn(1.f, b.z)};
    return r;
}

template <bool Count>
struct Tally {
    uint32_t nodes = 0, tris = 0, stack = 0;
    __device__ __forceinline__ void node() {
        if (Count) ++nodes;
    }
    __device__ __forceinline__ void tri() {
        if (Count) ++tris;
    }
    __device__ __forceinline__ void depth(uint32_t d) {
        if (Count) stack = d > stack ? d : stack;
    }
    __device__ void flush(TraceCounters* c) {
        if (!Count) return;
        unsigned long long n = nodes, t = tris;
        uint32_t           s = stack;
        for (int o = 16; o > 0; o >>= 1) {
            n += __shfl_down_sync(0xffffffffu, n, o);
            t += __shfl_down_sync(0xffffffffu, t, o);
            const uint32_t so = __shfl_down_sync(0xffffffffu, s, o);
            s                 = so > s ? so : s;
        }
        if (0 == (threadIdx.x & 31)) {
            atomicAdd(&c->nodes, n);
            atomicAdd(&c->triangles, t);
            atomicMax(&c->max_stack, (unsigned long long)s);
        }
    }
};

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

template <bool AnyHit, bool Count>
__global__ void __launch_bounds__(128)
    traceBinary(MeshDevice mesh, const RayIn* __restrict__ rays, void* __restrict__ out, uint32_t n,
                TraceCounters* counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    Tally<Count>   tally;
    if (i < n) {
        RayT ray = loadRay(rays, i);

        uint32_t stack[kBinaryStack];
        uint32_t end = 0;
        uint32_t node = 0;

        float    ht = 0.f, hu = 0.f, hv = 0.f;
        uint32_t primitive = kEnd;
        bool     occluded  = false;

        while (kEnd != node) {
            const float4 nmin = __ldg(mesh.binary_nodes + 2 * size_t(node));
            const float4 nmax = __ldg(mesh.binary_nodes + 2 * size_t(node) + 1);
            tally.node();

            const uint32_t num = __float_as_uint(nmax.w);
            if (0 != num) {
                uint32_t       p = __float_as_uint(nmin.w);
                const uint32_t e = p + num;
                for (; p < e; ++p) {
                    // triangle_data.zig:62-78 : index triple, then three position loads
                    const uint32_t ia = __ldg(mesh.triangles + 3 * size_t(p) + 0);
                    const uint32_t ib = __ldg(mesh.triangles + 3 * size_t(p) + 1);
                    const uint32_t ic = __ldg(mesh.triangles + 3 * size_t(p) + 2);
                    const float*   pa = mesh.positions + 3 * size_t(ia);
                    const float*   pb = mesh.positions + 3 * size_t(ib);
                    const float*   pc = mesh.positions + 3 * size_t(ic);
                    const V3       a  = {__ldg(pa), __ldg(pa + 1), __ldg(pa + 2)};
                    const V3       b  = {__ldg(pb), __ldg(pb + 1), __ldg(pb + 2)};
                    const V3       c  = {__ldg(pc), __ldg(pc + 1), __ldg(pc + 2)};
                    tally.tri();
                    float t, u, v;
                    if (intersectTriangle(ray, a, sub3(b, a), sub3(c, a), t, u, v)) {
                        if (AnyHit) {
                            occluded = true;
                            break;
                        }
                        ray.tmax  = t;
                        ht        = t;
                        hu        = u;
                        hv        = v;
                        primitive = p;
                    }
                }
                if (AnyHit && occluded) break;
                node = 0 == end ? kEnd : stack[--end];
                continue;
            }

            uint32_t a = __float_as_uint(nmin.w);
            uint32_t b = a + 1;

            const float4 amin = __ldg(mesh.binary_nodes + 2 * size_t(a));
            const float4 amax = __ldg(mesh.binary_nodes + 2 * size_t(a) + 1);
            const float4 bmin = __ldg(mesh.binary_nodes + 2 * size_t(b));
            const float4 bmax = __ldg(mesh.binary_nodes + 2 * size_t(b) + 1);

            float dista = intersectNode(amin, amax, ray);
            float distb = intersectNode(bmin, bmax, ray);

            if (dista > distb) {
                const uint32_t tn = a;
                a                 = b;
                b                 = tn;
                const float td    = dista;
                dista             = distb;
                distb             = td;
            }

            if (FLT_MAX == dista) {
                node = 0 == end ? kEnd : stack[--end];
            } else {
                node = a;
                if (FLT_MAX != distb) {
                    stack[end++] = b;
                    tally.depth(end);
                }
            }
        }

        if (AnyHit) {
            reinterpret_cast<uint32_t*>(out)[i] = occluded ? 1u : 0u;
        } else {
            float4 h;
            h.x = kEnd == primitive ? ray.tmax : ht;
            h.y = hu;
            h.z = hv;
            h.w = __uint_as_float(primitive);
            reinterpret_cast<float4*>(out)[i] = h;
        }
    }
    tally.flush(counters);
}

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

// One gated triangle test against record `index`; true on an accepted hit (fills t, u, v, primitive).
__device__ __forceinline__ bool testWideTriangle(const MeshDevice& mesh, const RayT& ray, uint32_t index, float& t,
                                                 float& u, float& v, uint32_t& primitive) {
    const float4* tp = mesh.wide_tris + 4 * size_t(index);
    const float4  t0 = __ldg(tp + 0);
    const float4  t1 = __ldg(tp + 1);
    const float4  t2 = __ldg(tp + 2);
    const float4  t3 = __ldg(tp + 3);

    // Gate with the reference's own (non-watertight) slab test on the reference leaf box: the
    // reference never tests a triangle whose leaf box it rejected.
    if (FLT_MAX == intersectNode(make_float4(t1.w, t2.w, t3.x, 0.f), make_float4(t3.y, t3.z, t3.w, 0.f), ray)) {
        return false;
    }
    primitive = __float_as_uint(t0.w);
    return intersectTriangle(ray, {t0.x, t0.y, t0.z}, {t1.x, t1.y, t1.z}, {t2.x, t2.y, t2.z}, t, u, v);
}

// Simple variant: one thread per ray, while-while loop. Kept as the A/B baseline for the persistent
// kernel below (ZYGPU_WIDE_VARIANT=0).
template <bool AnyHit, bool Count>
__global__ void __launch_bounds__(128)
    traceWide(MeshDevice mesh, const RayIn* __restrict__ rays, void* __restrict__ out, uint32_t n,
              TraceCounters* counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    Tally<Count>   tally;
    if (i < n) {
        WideRay w;
        w.ray = loadRay(rays, i);
        setupWideRay(w);

        uint2    stack[kWideStack];
        uint32_t sp = 0;

        // The root (node 0) is entered as the only hit child of a virtual parent whose inner children
        // start at index 0: bit 31 set, imask 0 -> rank 0.
        uint2 node_group = make_uint2(0u, 0x80000000u);
        uint2 tri_group  = make_uint2(0u, 0u);

        float    ht = 0.f, hu = 0.f, hv = 0.f;
        uint32_t primitive = kEnd;
        bool     occluded  = false;

        for (;;) {
            if (node_group.y > 0x00FFFFFFu) {
                // node groups carry (child_base, hits << 24 | imask); take the highest-priority hit child
                const uint32_t hits  = node_group.y;
                const uint32_t gmask = hits & 0xffu;
                const uint32_t bit   = 31u - __clz(hits);
                node_group.y         = hits & ~(1u << bit);
                const uint32_t slot  = (bit - 24u) ^ w.octinv;
                const uint32_t rank  = __popc(gmask & ((1u << slot) - 1u));
                const uint32_t node_index = node_group.x + rank;
                if (node_group.y > 0x00FFFFFFu) {
                    stack[sp++] = node_group;
                    tally.depth(sp);
                }

                const float4* np = mesh.wide_nodes + 5 * size_t(node_index);
                const float4  n0 = __ldg(np + 0);
                const float4  n1 = __ldg(np + 1);
                const float4  n2 = __ldg(np + 2);
                const float4  n3 = __ldg(np + 3);
                const float4  n4 = __ldg(np + 4);
                tally.node();

                const uint32_t hitmask = testWideNode(w, w.ray.tmin, w.ray.tmax, n0, n1, n2, n3, n4);

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
                tally.tri();

                float    t, u, v;
                uint32_t prim;
                if (testWideTriangle(mesh, w.ray, tri_group.x + bit, t, u, v, prim)) {
                    if (AnyHit) {
                        occluded = true;
                        break;
                    }
                    w.ray.tmax = t;
                    ht         = t;
                    hu         = u;
                    hv         = v;
                    primitive  = prim;
                }
            }
            if (AnyHit && occluded) break;

            if (node_group.y <= 0x00FFFFFFu) {
                if (0 == sp) break;
                node_group = stack[--sp];
            }
        }

        if (AnyHit) {
            reinterpret_cast<uint32_t*>(out)[i] = occluded ? 1u : 0u;
        } else {
            float4 h;
            h.x = kEnd == primitive ? w.ray.tmax : ht;
            h.y = hu;
            h.z = hv;
            h.w = __uint_as_float(primitive);
            reinterpret_cast<float4*>(out)[i] = h;
        }
    }
    tally.flush(counters);
}

// Persistent variant (product path). One resident warp per scheduler slot; every warp runs a
// warp-synchronous loop in which each step is either a NODE step (lanes that have an inner child to
// enter test its eight boxes) or a TRIANGLE step (lanes that have pending triangle bits test one
// triangle each), whichever has more lanes ready — pending triangle groups are postponed (pushed)
// instead of serialising the warp. Lanes whose ray has finished are refilled from the warp's private
// block of `kPoolRays` rays (blocks are handed out by one global atomic) once `fetch_idle` lanes are
// idle, so incoherent batches keep their lanes busy.
constexpr uint32_t kPoolRays = 1024;

struct WideTuning {
    uint32_t fetch_idle;  // refill when at least this many lanes are idle
    uint32_t tri_num;     // triangle step when ready_tri * tri_den >= ready_node * tri_num
    uint32_t tri_den;
};

template <bool AnyHit, bool Count>
__global__ void __launch_bounds__(128)
    traceWidePersistent(MeshDevice mesh, const RayIn* __restrict__ rays, void* __restrict__ out, uint32_t n,
                        TraceCounters* counters, uint32_t* __restrict__ work_counter, WideTuning tune) {
    constexpr uint32_t kFull = 0xffffffffu;
    const uint32_t     lane  = threadIdx.x & 31u;
    Tally<Count>       tally;

    uint32_t pool_next = 0, pool_end = 0;  // warp-uniform
    bool     exhausted = false;            // warp-uniform: the global counter ran past n

    bool     has_ray = false;
    uint32_t ray_index = 0;
    WideRay  w;
    uint2    stack[kWideStack];
    uint32_t sp         = 0;
    uint2    node_group = make_uint2(0u, 0u);
    uint2    tri_group  = make_uint2(0u, 0u);
    float    ht = 0.f, hu = 0.f, hv = 0.f;
    uint32_t primitive = kEnd;

    for (;;) {
        // ---- refill idle lanes
        uint32_t idle = __ballot_sync(kFull, !has_ray);
        while (0 != idle && !exhausted) {
            if (pool_next >= pool_end) {
                uint32_t base = 0;
                if (0 == lane) base = atomicAdd(work_counter, kPoolRays);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n) {
                    exhausted = true;
                    break;
                }
                pool_next = base;
                pool_end  = min(base + kPoolRays, n);
            }
            const uint32_t avail = pool_end - pool_next;
            const uint32_t rank  = __popc(idle & ((1u << lane) - 1u));
            if (!has_ray && rank < avail) {
                ray_index = pool_next + rank;
                w.ray     = loadRay(rays, ray_index);
                setupWideRay(w);
                sp         = 0;
                node_group = make_uint2(0u, 0x80000000u);  // root as the only hit child of a virtual parent
                tri_group  = make_uint2(0u, 0u);
                primitive  = kEnd;
                hu         = 0.f;
                hv         = 0.f;
                has_ray    = true;
            }
            pool_next += min(avail, (uint32_t)__popc(idle));
            idle = __ballot_sync(kFull, !has_ray);
        }
        if (kFull == idle) break;  // nothing left anywhere in this warp

        // ---- traverse in lock step until enough lanes went idle
        for (;;) {
            const bool     ready_node = has_ray && node_group.y > 0x00FFFFFFu;
            const bool     ready_tri  = has_ray && 0 != tri_group.y;
            const uint32_t mn         = __ballot_sync(kFull, ready_node);
            const uint32_t mt         = __ballot_sync(kFull, ready_tri);
            const uint32_t cn = __popc(mn), ct = __popc(mt);

            if (0 != ct && (0 == cn || ct * tune.tri_den >= cn * tune.tri_num)) {
                // TRIANGLE step
                if (ready_tri) {
                    const uint32_t bit = 31u - __clz(tri_group.y);
                    tri_group.y &= ~(1u << bit);
                    tally.tri();
                    float    t, u, v;
                    uint32_t prim;
                    if (testWideTriangle(mesh, w.ray, tri_group.x + bit, t, u, v, prim)) {
                        if (AnyHit) {
                            // occluded: drop all remaining work of this ray
                            primitive    = 0;
                            sp           = 0;
                            node_group.y = 0;
                            tri_group.y  = 0;
                        } else {
                            w.ray.tmax = t;
                            ht         = t;
                            hu         = u;
                            hv         = v;
                            primitive  = prim;
                        }
                    }
                }
            } else if (0 != cn) {
                // NODE step
                if (ready_node) {
                    const uint32_t hits  = node_group.y;
                    const uint32_t gmask = hits & 0xffu;
                    const uint32_t bit   = 31u - __clz(hits);
                    node_group.y         = hits & ~(1u << bit);
                    const uint32_t slot  = (bit - 24u) ^ w.octinv;
                    const uint32_t rank  = __popc(gmask & ((1u << slot) - 1u));
                    const uint32_t node_index = node_group.x + rank;
                    if (node_group.y > 0x00FFFFFFu) stack[sp++] = node_group;
                    if (0 != tri_group.y) stack[sp++] = tri_group;  // postponed triangles
                    tally.depth(sp);

                    const float4* np = mesh.wide_nodes + 5 * size_t(node_index);
                    const float4  n0 = __ldg(np + 0);
                    const float4  n1 = __ldg(np + 1);
                    const float4  n2 = __ldg(np + 2);
                    const float4  n3 = __ldg(np + 3);
                    const float4  n4 = __ldg(np + 4);
                    tally.node();

                    const uint32_t hitmask = testWideNode(w, w.ray.tmin, w.ray.tmax, n0, n1, n2, n3, n4);

                    node_group.x = __float_as_uint(n1.x);
                    node_group.y = (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
                    tri_group.x  = __float_as_uint(n1.y);
                    tri_group.y  = hitmask & 0x00FFFFFFu;
                }
            }

            // ---- lanes that ran dry pop their stack or retire their ray
            if (has_ray && node_group.y <= 0x00FFFFFFu && 0 == tri_group.y) {
                if (0 == sp) {
                    if (AnyHit) {
                        reinterpret_cast<uint32_t*>(out)[ray_index] = kEnd == primitive ? 0u : 1u;
                    } else {
                        float4 h;
                        h.x = kEnd == primitive ? w.ray.tmax : ht;
                        h.y = hu;
                        h.z = hv;
                        h.w = __uint_as_float(primitive);
                        reinterpret_cast<float4*>(out)[ray_index] = h;
                    }
                    has_ray = false;
                } else {
                    const uint2 e = stack[--sp];
                    if (e.y > 0x00FFFFFFu) {
                        node_group = e;
                    } else {
                        tri_group = e;
                    }
                }
            }

            const uint32_t active = __ballot_sync(kFull, has_ray);
            if (0 == active) break;
            if (!exhausted && 32u - __popc(active) >= tune.fetch_idle) break;
        }
    }
    tally.flush(counters);
}

int envInt(const char* name, int fallback) {
    const char* v = getenv(name);
    return v ? atoi(v) : fallback;
}

struct WideLaunchConfig {
    int        variant;  // 0: one thread per ray, 1: persistent
    WideTuning tune;
    int        blocks_per_sm;
    int        num_sms;
};

const WideLaunchConfig& wideConfig() {
    static const WideLaunchConfig cfg = [] {
        WideLaunchConfig c;
        c.variant         = envInt("ZYGPU_WIDE_VARIANT", 1);
        c.tune.fetch_idle = uint32_t(envInt("ZYGPU_FETCH_IDLE", 6));
        c.tune.tri_num    = uint32_t(envInt("ZYGPU_TRI_NUM", 1));
        c.tune.tri_den    = uint32_t(envInt("ZYGPU_TRI_DEN", 2));
        c.blocks_per_sm   = envInt("ZYGPU_BLOCKS_PER_SM", 0);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev);
        return c;
    }();
    return cfg;
}

template <bool AnyHit, bool Count>
cudaError_t launchOne(const MeshDevice& mesh, bool wide, const RayIn* rays, void* out, uint32_t n,
                      TraceCounters* counters, uint32_t* work_counter, cudaStream_t stream) {
    if (0 == n) return cudaSuccess;
    const uint32_t block = 128;
    const uint32_t grid  = (n + block - 1) / block;
    if (!wide) {
        traceBinary<AnyHit, Count><<<grid, block, 0, stream>>>(mesh, rays, out, n, counters);
        return cudaGetLastError();
    }
    const WideLaunchConfig& cfg = wideConfig();
    if (0 == cfg.variant || !work_counter) {
        traceWide<AnyHit, Count><<<grid, block, 0, stream>>>(mesh, rays, out, n, counters);
        return cudaGetLastError();
    }
    // persistent: as many resident blocks as fit, a multiple of the SM count
    static int resident = 0;  // per template instance
    if (0 == resident) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, traceWidePersistent<AnyHit, Count>, int(block), 0);
        if (cfg.blocks_per_sm > 0) per_sm = std::min(per_sm, cfg.blocks_per_sm);
        resident = std::max(per_sm, 1) * cfg.num_sms;
    }
    const uint32_t needed = (n + 31) / 32;  // warps; never launch more warps than there are 32-ray groups
    const uint32_t pgrid  = std::min<uint32_t>(uint32_t(resident), (needed + 3) / 4);
    cudaError_t    err    = cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), stream);
    if (cudaSuccess != err) return err;
    traceWidePersistent<AnyHit, Count><<<pgrid, block, 0, stream>>>(mesh, rays, out, n, counters, work_counter, cfg.tune);
    return cudaGetLastError();
}

__global__ void addRays(TraceCounters* c, unsigned long long n) { c->rays += n; }

}  // namespace

cudaError_t launchTrace(const MeshDevice& mesh, int mode, const RayIn* rays, void* out, uint32_t n,
                        TraceCounters* counters, uint32_t* work_counter, cudaStream_t stream) {
    const bool wide = kClosestWide == mode || kAnyWide == mode;
    const bool any  = kAnyWide == mode || kAnyBinary == mode;
    if (mode < 0 || mode > 3) return cudaErrorInvalidValue;

    cudaError_t err;
    if (counters) {
        addRays<<<1, 1, 0, stream>>>(counters, n);
        err = any ? launchOne<true, true>(mesh, wide, rays, out, n, counters, work_counter, stream)
                  : launchOne<false, true>(mesh, wide, rays, out, n, counters, work_counter, stream);
    } else {
        err = any ? launchOne<true, false>(mesh, wide, rays, out, n, nullptr, work_counter, stream)
                  : launchOne<false, false>(mesh, wide, rays, out, n, nullptr, work_counter, stream);
    }
    return err;
}

}  // namespace zygpu
