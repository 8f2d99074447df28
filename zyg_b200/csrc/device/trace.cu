#include "trace_device.cuh"

#include <algorithm>
#include <cfloat>
#include <cstdlib>

namespace zygpu {

namespace {

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

                const WideNodeRegs nd = loadWideNode(mesh.wide_nodes, node_index);
                const float4 n0 = nd.n0, n1 = nd.n1, n2 = nd.n2, n3 = nd.n3, n4 = nd.n4;
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
                if (testWideTriangle(mesh, w.ray, w.ray.tmax, tri_group.x + bit, t, u, v, prim)) {
                    if (AnyHit) {
                        occluded = true;
                        break;
                    }
                    if (kEnd == primitive || t < w.ray.tmax || prim > primitive) {  // equal t: the larger primitive id wins
                        w.ray.tmax = t;
                        ht         = t;
                        hu         = u;
                        hv         = v;
                        primitive  = prim;
                    }
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
    float    tmax0 = 0.f;  // the max_t the ray started with: the limit of the leaf gates (gateBox)
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
                tmax0     = w.ray.tmax;
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
                    if (testWideTriangle(mesh, w.ray, tmax0, tri_group.x + bit, t, u, v, prim)) {
                        if (AnyHit) {
                            // occluded: drop all remaining work of this ray
                            primitive    = 0;
                            sp           = 0;
                            node_group.y = 0;
                            tri_group.y  = 0;
                        } else if (kEnd == primitive || t < w.ray.tmax || prim > primitive) {  // equal t: the larger id wins
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

                    const WideNodeRegs nd = loadWideNode(mesh.wide_nodes, node_index);
                    const float4 n0 = nd.n0, n1 = nd.n1, n2 = nd.n2, n3 = nd.n3, n4 = nd.n4;
                    tally.node();

                    const uint32_t hitmask = testWideNode(w, w.ray.tmin, cullLimit(w.ray), n0, n1, n2, n3, n4);

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

// Ray-pool variant (product path). The lock-step kernel above keeps one ray per lane, so a NODE step only uses the lanes whose ray
// happens to want a node test (measured: 18 - 22 of 32) and a TRIANGLE step 9 - 11. Here a warp keeps `kPoolSlots` rays in flight in
// shared memory (ray, traversal groups, hit so far: 80 bytes per ray) and every step first hands the rays that are ready for the
// chosen kind to the lanes, so the test code runs on (nearly) full warps whenever the pool holds 32 such rays. The per-ray
// traversal stacks live in a global scratch array (cached like local memory). Same tests, same node / triangle order per ray up
// to the postponement of triangle groups, so results are the lock-step kernel's: t, u, v bit-identical, primitives identical up
// to equal-t ties — which are now resolved by primitive id (the larger BVH-order index wins, the reference's "later hit wins"
// within a leaf), independent of the schedule.
constexpr uint32_t kPoolSlots = 64;

struct PoolTuning {
    uint32_t fetch_free;  // refill when at least this many slots are free
    uint32_t tri_num;     // triangle step when ready_tri * tri_den >= ready_node * tri_num
    uint32_t tri_den;
    uint32_t prefetch;    // 1: prefetch the node / triangle record a ray's next step reads
};

struct RayPool {  // one per warp
    float4   a[kPoolSlots];  // origin xyz | min_t
    float4   b[kPoolSlots];  // direction xyz | max_t (the closest hit so far)
    float4   c[kPoolSlots];  // inverse direction xyz | u of the hit
    uint4    g[kPoolSlots];  // node group | triangle group
    uint4    h[kPoolSlots];  // stack depth | ray index (kEnd: the slot is free) | primitive | v of the hit
    float    t0[kPoolSlots];  // the max_t the ray started with: the limit of the leaf gates (gateBox)
    uint32_t assign[32];
};

template <bool AnyHit, bool Count>
__global__ void __launch_bounds__(128)
    traceWidePool(MeshDevice mesh, const RayIn* __restrict__ rays, void* __restrict__ out, uint32_t n, TraceCounters* counters,
                  uint32_t* __restrict__ work_counter, uint2* __restrict__ stacks, PoolTuning tune) {
    constexpr uint32_t kFull = 0xffffffffu;
    __shared__ RayPool pools[4];
    RayPool&           pool = pools[threadIdx.x >> 5];
    const uint32_t     lane = threadIdx.x & 31u;
    const uint32_t     lt   = (1u << lane) - 1u;
    uint2* __restrict__ stk = stacks + size_t(blockIdx.x * 4u + (threadIdx.x >> 5)) * kPoolSlots * kWideStack;
    Tally<Count>       tally;

    for (uint32_t s = lane; s < kPoolSlots; s += 32) {
        pool.g[s] = make_uint4(0u, 0u, 0u, 0u);
        pool.h[s] = make_uint4(0u, kEnd, kEnd, 0u);
    }
    __syncwarp();

    uint32_t occupied  = 0;  // warp-uniform
    uint32_t pool_next = 0, pool_end = 0;
    bool     exhausted = false;

    for (;;) {
        // ---- refill free slots
        if (!exhausted && kPoolSlots - occupied >= tune.fetch_free) {
            for (uint32_t half = 0; half < kPoolSlots / 32; ++half) {
                const uint32_t slot = lane + 32u * half;
                uint32_t       need = __ballot_sync(kFull, kEnd == pool.h[slot].y);
                while (0 != need && !exhausted) {
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
                    const bool     take  = 0 != ((need >> lane) & 1u) && __popc(need & lt) < avail;
                    if (take) {
                        const uint32_t index = pool_next + __popc(need & lt);
                        const RayT     r     = loadRay(rays, index);
                        pool.a[slot]         = make_float4(r.o.x, r.o.y, r.o.z, r.tmin);
                        pool.b[slot]         = make_float4(r.d.x, r.d.y, r.d.z, r.tmax);
                        pool.c[slot]         = make_float4(r.inv_d.x, r.inv_d.y, r.inv_d.z, 0.f);
                        pool.g[slot]         = make_uint4(0u, 0x80000000u, 0u, 0u);  // root as the only hit child of a virtual parent
                        pool.h[slot]         = make_uint4(0u, index, kEnd, 0u);
                        pool.t0[slot]        = r.tmax;
                    }
                    const uint32_t taken = __ballot_sync(kFull, take);
                    pool_next += __popc(taken);
                    occupied += __popc(taken);
                    need &= ~taken;
                }
            }
            __syncwarp();
        }
        if (0 == occupied) {
            if (exhausted) break;
            continue;
        }

        // ---- which rays are ready for which kind of step
        uint32_t m_node[kPoolSlots / 32], m_tri[kPoolSlots / 32];
        uint32_t cn = 0, ct = 0;
#pragma unroll
        for (uint32_t half = 0; half < kPoolSlots / 32; ++half) {
            const uint4 g = pool.g[lane + 32u * half];
            m_node[half]  = __ballot_sync(kFull, g.y > 0x00FFFFFFu);
            m_tri[half]   = __ballot_sync(kFull, 0 != g.w);
            cn += __popc(m_node[half]);
            ct += __popc(m_tri[half]);
        }
        const bool tri_step = 0 != ct && (0 == cn || ct * tune.tri_den >= cn * tune.tri_num);

        // ---- hand the first 32 ready rays to the lanes
        uint32_t rank = 0;
#pragma unroll
        for (uint32_t half = 0; half < kPoolSlots / 32; ++half) {
            const uint32_t m = tri_step ? m_tri[half] : m_node[half];
            if (0 != ((m >> lane) & 1u)) {
                const uint32_t r = rank + __popc(m & lt);
                if (r < 32u) pool.assign[r] = lane + 32u * half;
            }
            rank += __popc(m);
        }
        __syncwarp();
        const bool     busy = lane < rank;
        const uint32_t slot = busy ? pool.assign[lane] : 0u;

        uint32_t retired = 0;
        if (busy) {
            const float4 ra = pool.a[slot];
            const float4 rb = pool.b[slot];
            const float4 rc = pool.c[slot];
            uint4        g  = pool.g[slot];
            uint4        h  = pool.h[slot];

            WideRay w;
            w.ray.o     = {ra.x, ra.y, ra.z};
            w.ray.tmin  = ra.w;
            w.ray.d     = {rb.x, rb.y, rb.z};
            w.ray.tmax  = rb.w;
            w.ray.inv_d = {rc.x, rc.y, rc.z};
            setupWideRay(w);
            uint2* __restrict__ my_stack = stk + size_t(slot) * kWideStack;

            uint2    node_group = make_uint2(g.x, g.y);
            uint2    tri_group  = make_uint2(g.z, g.w);
            uint32_t sp         = h.x;

            if (tri_step) {
                const uint32_t bit = 31u - __clz(tri_group.y);
                tri_group.y &= ~(1u << bit);
                tally.tri();
                float    t, u, v;
                uint32_t prim;
                if (testWideTriangle(mesh, w.ray, pool.t0[slot], tri_group.x + bit, t, u, v, prim)) {
                    if (AnyHit) {
                        h.z          = 0;  // occluded: drop all remaining work of this ray
                        sp           = 0;
                        node_group.y = 0;
                        tri_group.y  = 0;
                    } else if (kEnd == h.z || t < w.ray.tmax || prim > h.z) {  // equal t: the larger primitive id wins
                        w.ray.tmax      = t;
                        pool.b[slot].w  = t;
                        pool.c[slot].w  = u;
                        h.w             = __float_as_uint(v);
                        h.z             = prim;
                    }
                }
            } else {
                const uint32_t hits  = node_group.y;
                const uint32_t gmask = hits & 0xffu;
                const uint32_t bit   = 31u - __clz(hits);
                node_group.y         = hits & ~(1u << bit);
                const uint32_t cslot = (bit - 24u) ^ w.octinv;
                const uint32_t crank = __popc(gmask & ((1u << cslot) - 1u));
                const uint32_t node_index = node_group.x + crank;
                if (node_group.y > 0x00FFFFFFu) my_stack[sp++] = node_group;
                if (0 != tri_group.y) my_stack[sp++] = tri_group;  // postponed triangles
                tally.depth(sp);

                const WideNodeRegs nd = loadWideNode(mesh.wide_nodes, node_index);
                const float4 n0 = nd.n0, n1 = nd.n1, n2 = nd.n2, n3 = nd.n3, n4 = nd.n4;
                tally.node();

                const uint32_t hitmask = testWideNode(w, w.ray.tmin, cullLimit(w.ray), n0, n1, n2, n3, n4);

                node_group.x = __float_as_uint(n1.x);
                node_group.y = (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
                tri_group.x  = __float_as_uint(n1.y);
                tri_group.y  = hitmask & 0x00FFFFFFu;
            }

            // ---- a ray that ran dry pops its stack or retires
            if (node_group.y <= 0x00FFFFFFu && 0 == tri_group.y) {
                if (0 == sp) {
                    if (AnyHit) {
                        reinterpret_cast<uint32_t*>(out)[h.y] = kEnd == h.z ? 0u : 1u;
                    } else {
                        float4 r;
                        r.x = w.ray.tmax;  // the hit's t, or max_t on a miss
                        r.y = kEnd == h.z ? 0.f : pool.c[slot].w;
                        r.z = kEnd == h.z ? 0.f : __uint_as_float(h.w);
                        r.w = __uint_as_float(h.z);
                        reinterpret_cast<float4*>(out)[h.y] = r;
                    }
                    h.y     = kEnd;
                    retired = 1;
                } else {
                    const uint2 e = my_stack[--sp];
                    if (e.y > 0x00FFFFFFu) {
                        node_group = e;
                    } else {
                        tri_group = e;
                    }
                }
            }
            h.x          = sp;
            pool.g[slot] = make_uint4(node_group.x, node_group.y, tri_group.x, tri_group.y);
            pool.h[slot] = h;

            // the ray waits in the pool for at least one step of the warp: start fetching what its next step will read
            if (0 != tune.prefetch) {
                if (node_group.y > 0x00FFFFFFu) {
                    const uint32_t nbit  = 31u - __clz(node_group.y);
                    const uint32_t nslot = (nbit - 24u) ^ w.octinv;
                    const float4*  np    = mesh.wide_nodes + kWideNodeWords * size_t(node_group.x + __popc(node_group.y & 0xffu & ((1u << nslot) - 1u)));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 4));
                }
                if (0 != tri_group.y) {
                    const float4* tp = mesh.wide_tris + 4 * size_t(tri_group.x + (31u - __clz(tri_group.y)));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(tp));
                }
            }
        }
        occupied -= __popc(__ballot_sync(kFull, 0 != retired));
        __syncwarp();
    }
    tally.flush(counters);
}

int envInt(const char* name, int fallback) {
    const char* v = getenv(name);
    return v ? atoi(v) : fallback;
}

struct WideLaunchConfig {
    int        variant;  // 0: one thread per ray, 1: persistent lock-step (one ray per lane), 2: persistent ray pool (default)
    WideTuning tune;
    PoolTuning pool;
    int        blocks_per_sm;
    int        num_sms;
};

const WideLaunchConfig& wideConfig() {
    static const WideLaunchConfig cfg = [] {
        WideLaunchConfig c;
        c.variant         = envInt("ZYGPU_WIDE_VARIANT", 2);
        c.pool.fetch_free = uint32_t(envInt("ZYGPU_POOL_FETCH_FREE", 32));
        c.pool.tri_num    = uint32_t(envInt("ZYGPU_POOL_TRI_NUM", 1));
        c.pool.tri_den    = uint32_t(envInt("ZYGPU_POOL_TRI_DEN", 1));
        c.pool.prefetch   = uint32_t(envInt("ZYGPU_POOL_PREFETCH", 0));  // measured: prefetch.global.L1 of the next record halves the incoherent rate (more L1 requests, the bound)
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
                      TraceCounters* counters, uint32_t* work_counter, uint2* stacks, uint32_t max_stack_blocks, cudaStream_t stream) {
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
    if (2 == cfg.variant && stacks) {
        static int pool_resident = 0;  // per template instance
        if (0 == pool_resident) {
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, traceWidePool<AnyHit, Count>, int(block), 0);
            if (cfg.blocks_per_sm > 0) per_sm = std::min(per_sm, cfg.blocks_per_sm);
            pool_resident = std::max(per_sm, 1) * cfg.num_sms;
        }
        const uint32_t groups = (n + kPoolSlots - 1) / kPoolSlots;  // never more warps than there are pools to fill
        const uint32_t pgrid  = std::min<uint32_t>(std::min<uint32_t>(uint32_t(pool_resident), max_stack_blocks), (groups + 3) / 4);
        cudaError_t    perr   = cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), stream);
        if (cudaSuccess != perr) return perr;
        traceWidePool<AnyHit, Count><<<pgrid, block, 0, stream>>>(mesh, rays, out, n, counters, work_counter, stacks, cfg.pool);
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

size_t traceStackBytesPerBlock() { return size_t(4) * kPoolSlots * kWideStack * sizeof(uint2); }

cudaError_t launchTrace(const MeshDevice& mesh, int mode, const RayIn* rays, void* out, uint32_t n,
                        TraceCounters* counters, uint32_t* work_counter, void* stack_scratch, size_t stack_scratch_bytes, cudaStream_t stream) {
    uint2*         stacks           = static_cast<uint2*>(stack_scratch);
    const uint32_t max_stack_blocks = uint32_t(stack_scratch_bytes / traceStackBytesPerBlock());
    const bool wide = kClosestWide == mode || kAnyWide == mode;
    const bool any  = kAnyWide == mode || kAnyBinary == mode;
    if (mode < 0 || mode > 3) return cudaErrorInvalidValue;

    cudaError_t err;
    if (counters) {
        addRays<<<1, 1, 0, stream>>>(counters, n);
        err = any ? launchOne<true, true>(mesh, wide, rays, out, n, counters, work_counter, stacks, max_stack_blocks, stream)
                  : launchOne<false, true>(mesh, wide, rays, out, n, counters, work_counter, stacks, max_stack_blocks, stream);
    } else {
        err = any ? launchOne<true, false>(mesh, wide, rays, out, n, nullptr, work_counter, stacks, max_stack_blocks, stream)
                  : launchOne<false, false>(mesh, wide, rays, out, n, nullptr, work_counter, stacks, max_stack_blocks, stream);
    }
    return err;
}

}  // namespace zygpu
