// Traversal stages of the forward pass: Scene.intersect / Scene.visibility for the queued rays.
#include "render_common.cuh"

namespace zygpu {

namespace {

// Item ids: closest-hit rays are identified by their path slot, shadow rays by their record (slot * stride + k).
#ifndef ZYGPU_TOP_BLOCKS
#define ZYGPU_TOP_BLOCKS 6  // 80 registers: measured +4 % on the instanced scene over the unbounded 96-register build
#endif
template <bool AnyHit>
__global__ void __launch_bounds__(kBlock, ZYGPU_TOP_BLOCKS) topKernel(SceneDevice sc, PathState st) {
    const uint32_t stride = st.shadow_stride;
    const uint32_t* __restrict__ closest_queue = st.lanes > 1 ? st.queue_t : st.queue_a;
    const bool     compact = AnyHit && nullptr != st.queue_r;  // shadow records listed in queue_r instead of stride per slot
    const uint64_t total   = AnyHit ? (compact ? uint64_t(st.counters[10]) : uint64_t(st.counters[1]) * stride)
                                    : uint64_t(st.counters[st.lanes > 1 ? 7 : 0]);
    const uint32_t count  = uint32_t(total < 0xFFFFFFFFull ? total : 0xFFFFFFFFull);
    const uint32_t iters  = (count + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    uint32_t       traced = 0;

    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t i       = it * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
        bool           to_mesh = false;
        uint32_t       item    = 0;
        bool           valid   = i < count;
        if (valid) {
            if (compact) {
                item = st.queue_r[i];
            } else if (AnyHit) {
                const uint32_t slot = st.queue_b[i / stride];
                const uint32_t k    = i % stride;
                valid               = k < st.sh_n[slot];
                item                = slot * stride + k;
            } else {
                item = closest_queue[i];
            }
        }
        if (valid) {
            traced += 1;
            uint32_t depth_surface, flags = 0;
            RayT     ray = loadTraceRay<AnyHit>(st, item, depth_surface, &flags);
            if (!AnyHit) clipToMedium(sc, st, item, flags, ray);

            uint32_t stack[kPropStack];
            uint32_t end = 0;
            uint32_t n   = 0 == sc.num_solid_nodes ? kEnd : 0;

            HitD     isec       = {0.f, 0.f, 0.f, 0};
            uint32_t hit_prop   = kEnd;
            bool     occluded   = false;
            uint32_t candidates = 0;

            while (kEnd != n && !(AnyHit && occluded)) {
                const float4 nmin = __ldg(sc.solid_nodes + 2 * size_t(n));
                const float4 nmax = __ldg(sc.solid_nodes + 2 * size_t(n) + 1);

                const uint32_t num = __float_as_uint(nmax.w);
                if (0 != num) {
                    const uint32_t start = __float_as_uint(nmin.w);
                    for (uint32_t li = start; li < start + num; ++li) {
                        const uint32_t  p    = __ldg(sc.solid_indices + li);
                        const ZygpuProp prop = sc.props[p];
                        if (ZYG_SHAPE_TRIANGLE_MESH == prop.shape && candidates < kMeshCandidates) {
                            // Prop.intersect / Prop.visibility up to the shape call, prop.zig:176-183, 212-218
                            if (AnyHit ? 0 == (prop.flags & ZYG_PROP_VISIBLE_IN_SHADOW) : !propVisible(prop.flags, depth_surface)) continue;
                            if (!aabbIntersect(sc.aabbs, p, ray)) continue;
                            st.ml_props[size_t(item) * kMeshCandidates + candidates] = p;
                            candidates += 1;
                            continue;
                        }
                        if (AnyHit) {
                            if (!propVisibility(sc, p, ray)) {
                                occluded = true;
                                break;
                            }
                        } else {
                            HitD h;
                            if (propIntersect(sc, p, ray, depth_surface, h)) {
                                ray.tmax = h.t;
                                isec     = h;
                                hit_prop = p;
                            }
                        }
                    }
                    n = 0 == end ? kEnd : stack[--end];
                    continue;
                }

                uint32_t a = __float_as_uint(nmin.w);
                uint32_t b = a + 1;

                float dista = intersectNode(__ldg(sc.solid_nodes + 2 * size_t(a)), __ldg(sc.solid_nodes + 2 * size_t(a) + 1), ray);
                float distb = intersectNode(__ldg(sc.solid_nodes + 2 * size_t(b)), __ldg(sc.solid_nodes + 2 * size_t(b) + 1), ray);
                if (dista > distb) {
                    const uint32_t tn = a;
                    a                 = b;
                    b                 = tn;
                    const float td    = dista;
                    dista             = distb;
                    distb             = td;
                }
                if (FLT_MAX == dista) {
                    n = 0 == end ? kEnd : stack[--end];
                } else {
                    n = a;
                    if (FLT_MAX != distb) stack[end++] = b;
                }
            }

            if (AnyHit) {
                st.sh_wi[item].w = occluded ? 0.f : 1.f;
                to_mesh          = !occluded && 0 != candidates;
            } else {
                st.ray_d[item].w = ray.tmax;
                st.hit[item]     = make_float4(isec.u, isec.v, __uint_as_float(isec.primitive), __uint_as_float(hit_prop));
                to_mesh          = 0 != candidates;
            }
            if (to_mesh) st.ml_count[item] = candidates;
        }
        queuePush(st.queue_m, &st.counters[2], to_mesh, item);
    }
    for (int o = 16; o > 0; o >>= 1) traced += __shfl_down_sync(0xffffffffu, traced, o);
    if (0 == (threadIdx.x & 31u) && 0 != traced) atomicAdd(&st.counters[AnyHit ? 6 : 5], traced);
}

template <bool AnyHit, int MinBlocks>
__global__ void __launch_bounds__(128, MinBlocks) meshTracePersistent(SceneDevice sc, PathState st, uint32_t* __restrict__ work_counter,
                                                           SceneTraceTuning tune) {
    constexpr uint32_t kFull = 0xffffffffu;
    const uint32_t     lane  = threadIdx.x & 31u;
    const uint32_t     n     = st.counters[2];

    // Queue items are handed out in pools: large pools keep the atomic cold on big queues, small pools spread a short
    // queue (late bounces) over all resident warps instead of leaving it to a few.
    const uint32_t warps      = gridDim.x * (blockDim.x / 32u);
    const uint32_t pool_items = max(32u, min(kScenePoolItems, (n / (warps * 4u)) & ~31u));

    uint32_t pool_next = 0, pool_end = 0;
    bool     exhausted = false;

    bool     has_ray = false;  // the lane owns a ray (between candidates or inside a mesh)
    bool     in_mesh = false;
    uint32_t item    = 0;
    float    tmax    = 0.f;  // world max_t == object max_t
    float    tmax0   = 0.f;  // the max_t the mesh walk started with (after the analytic props): the limit of the leaf gates
    uint32_t cand_i = 0, cand_n = 0;
    uint32_t cur_prop = 0;
    uint32_t hit_prop = kEnd;
    bool     occluded = false;
    MeshDevice mesh;  // of the mesh the lane is inside: only the two wide arrays are read
    mesh.wide_nodes = nullptr;
    mesh.wide_tris  = nullptr;

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
                if (0 == lane) base = atomicAdd(work_counter, pool_items);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n) {
                    exhausted = true;
                    break;
                }
                pool_next = base;
                pool_end  = min(base + pool_items, n);
            }
            const uint32_t avail = pool_end - pool_next;
            const uint32_t rank  = __popc(idle & ((1u << lane) - 1u));
            if (!has_ray && rank < avail) {
                item     = st.queue_m[pool_next + rank];
                cand_i   = 0;
                cand_n   = st.ml_count[item];
                has_ray  = true;
                in_mesh  = false;
                hit_prop = kEnd;
                occluded = false;
                tmax     = AnyHit ? 0.f : st.ray_d[item].w;
                tmax0    = tmax;
            }
            pool_next += min(avail, (uint32_t)__popc(idle));
            idle = __ballot_sync(kFull, !has_ray);
        }
        if (kFull == idle) break;

        // ---- lanes between candidates: enter the next mesh or retire
        while (has_ray && !in_mesh) {
            if (cand_i == cand_n || (AnyHit && occluded)) {
                if (AnyHit) {
                    if (occluded) st.sh_wi[item].w = 0.f;
                } else if (kEnd != hit_prop) {
                    st.ray_d[item].w = tmax;
                    st.hit[item]     = make_float4(hu, hv, __uint_as_float(primitive), __uint_as_float(hit_prop));
                }
                has_ray = false;
                break;
            }
            const uint32_t p = st.ml_props[size_t(item) * kMeshCandidates + cand_i];
            cand_i += 1;

            uint32_t depth_surface;
            RayT     ray = loadTraceRay<AnyHit>(st, item, depth_surface);
            if (!AnyHit) ray.tmax = tmax;
            // (the first candidate was tested by the top kernel)
            if (0 != cand_i - 1 && !gateBox(__ldg(sc.aabbs + 2 * size_t(p)), __ldg(sc.aabbs + 2 * size_t(p) + 1), ray, AnyHit ? ray.tmax : tmax0)) continue;

            const TrafoD trafo = loadTrafo(sc.trafos, p);
            w.ray              = worldToObjectRay(trafo, ray);  // triangle_tree.zig:49: t is shared with world space
            setupWideRay(w);
            cur_prop = p;
            {
                const MeshDevice* m = sc.meshes + sc.props[p].mesh;
                mesh.wide_nodes     = m->wide_nodes;
                mesh.wide_tris      = m->wide_tris;
            }
            in_mesh    = true;
            sp         = 0;
            node_group = make_uint2(0u, 0x80000000u);
            tri_group  = make_uint2(0u, 0u);
        }

        // ---- lock-step NODE / TRIANGLE steps over the lanes inside a mesh
        for (;;) {
            const bool     ready_node = in_mesh && node_group.y > 0x00FFFFFFu;
            const bool     ready_tri  = in_mesh && 0 != tri_group.y;
            const uint32_t mn         = __ballot_sync(kFull, ready_node);
            const uint32_t mt         = __ballot_sync(kFull, ready_tri);
            const uint32_t cn = __popc(mn), ct = __popc(mt);
            if (0 == cn && 0 == ct) break;

            if (0 != ct && (0 == cn || ct * tune.tri_den >= cn * tune.tri_num)) {
                if (ready_tri) {
                    const uint32_t bit = 31u - __clz(tri_group.y);
                    tri_group.y &= ~(1u << bit);
                    float    t, u, v;
                    uint32_t prim;
                    if (testWideTriangle(mesh, w.ray, AnyHit ? w.ray.tmax : tmax0, tri_group.x + bit, t, u, v, prim)) {
                        if (AnyHit) {
                            occluded     = true;
                            sp           = 0;
                            node_group.y = 0;
                            tri_group.y  = 0;
                        } else if (closerOrLater(t, w.ray.tmax, cur_prop, prim, hit_prop, primitive)) {
                            w.ray.tmax = t;
                            tmax       = t;  // probe.ray.max_t = isec.t, prop_tree.zig:77
                            hu         = u;
                            hv         = v;
                            primitive  = prim;
                            hit_prop   = cur_prop;
                        }
                    }
                }
            } else if (ready_node) {
                const uint32_t hits  = node_group.y;
                const uint32_t gmask = hits & 0xffu;
                const uint32_t bit   = 31u - __clz(hits);
                node_group.y         = hits & ~(1u << bit);
                const uint32_t slot  = (bit - 24u) ^ w.octinv;
                const uint32_t rank  = __popc(gmask & ((1u << slot) - 1u));
                const uint32_t node_index = node_group.x + rank;
                if (node_group.y > 0x00FFFFFFu) stack[sp++] = node_group;
                if (0 != tri_group.y) stack[sp++] = tri_group;

                const WideNodeRegs nd = loadWideNode(mesh.wide_nodes, node_index);
                const float4 n0 = nd.n0, n1 = nd.n1, n2 = nd.n2, n3 = nd.n3, n4 = nd.n4;

                const uint32_t hitmask = testWideNode(w, w.ray.tmin, cullLimit(w.ray), n0, n1, n2, n3, n4);

                node_group.x = __float_as_uint(n1.x);
                node_group.y = (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
                tri_group.x  = __float_as_uint(n1.y);
                tri_group.y  = hitmask & 0x00FFFFFFu;
            }

            // lanes that ran dry pop their stack or leave the mesh
            if (in_mesh && node_group.y <= 0x00FFFFFFu && 0 == tri_group.y) {
                if (0 == sp) {
                    in_mesh = false;
                } else {
                    const uint2 e = stack[--sp];
                    if (e.y > 0x00FFFFFFu) {
                        node_group = e;
                    } else {
                        tri_group = e;
                    }
                }
            }

            const uint32_t inside = __popc(__ballot_sync(kFull, in_mesh));
            if (0 == inside) break;
            if (32u - inside >= tune.fetch_idle) {
                // enough lanes left their mesh: let them move on / be refilled, unless nothing is left for them to do
                const uint32_t waiting = __ballot_sync(kFull, has_ray && !in_mesh);
                if (0 != waiting || !exhausted) break;
            }
        }
    }
}

// ---- fused two-level traversal -------------------------------------------------------------------------------------
//
// One persistent kernel walks both levels of the "two-level layout for prop instances": the prop tree (PropBvh, prop_tree.zig:
// 56-240) collapsed on upload into the same 80-byte 8-wide quantised nodes as the mesh trees, its leaf slots pointing at prop
// records {prop id, exact box of the reference leaf}. A lane owns a ray from the trace queue until the ray is done; the warp
// runs lock-step steps of three kinds, each over the lanes that have that kind of work:
//
//   NODE      test the eight quantised child boxes of one wide node — the same code for a lane in the prop tree and a lane
//             inside a mesh, only the node array differs
//   TRIANGLE  one gated triangle test (lanes inside a mesh)
//   PROP      one prop record (lanes in the prop tree): reference leaf gate, visibility flags, the prop's world box against the
//             current max_t (Prop.intersect up to the shape call, prop.zig:163-197), then an analytic shape in place or entry into
//             a mesh: the ray goes to object space, the prop-tree work still pending is left on the lane's stack below the mesh's
//
// Children are visited front to back by octant, every test uses the ray's current max_t, so instances behind the closest hit
// so far are culled at the node or at their world box; nothing but the result goes back to HBM (the former top kernel wrote
// 8 candidate props per ray and the mesh kernel read them back). Relative to the reference only the order in which props are
// tested changes: the closest hit is identical except for equal-t ties between different props.

// Conservative: false only if the segment [tmin, tmax] of the ray cannot touch the sphere (xyz centre, w radius).
__device__ __forceinline__ bool segmentMeetsSphere(const RayT& ray, float4 sphere) {
    if (FLT_MAX == sphere.w) return true;
    const V3    oc = {sphere.x - ray.o.x, sphere.y - ray.o.y, sphere.z - ray.o.z};
    const float dd = dot3(ray.d, ray.d);
    const float b  = dot3(oc, ray.d);
    const float r2 = sphere.w * sphere.w;
    const float oo = dot3(oc, oc);
    if (oo <= r2) return true;  // the origin is inside
    if (b <= 0.f) return false;  // outside and heading away
    const float tc = __fdividef(b, dd);  // parameter of the closest approach
    const V3    pv = {oc.x - tc * ray.d.x, oc.y - tc * ray.d.y, oc.z - tc * ray.d.z};
    if (dot3(pv, pv) > r2 * 1.0001f) return false;
    // the entry point is no nearer than tc - r / |d|
    return tc - sphere.w * rsqrtf(dd) * 1.0001f <= cullLimit(ray);
}

// Prefetches. The kernel waits on memory, not on bandwidth (DRAM 6 % busy, half of the stall samples on the first use of a loaded
// record): a lane asks for the records it is going to read as soon as it knows which, without holding registers for them.
__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// mode 1: the one record of the lane's next step, to L1; 2: every 64-byte record of the leaf slots just hit, to L2; 3: those and the
// inner children that go to the stack; 4: like 2, to L1
__device__ __forceinline__ void prefetchNext(uint32_t mode, const float4* nodes, const float4* recs, uint2 node_group, uint2 tri_group,
                                             uint32_t octinv) {
    if (1 == mode) {
        if (node_group.y > 0x00FFFFFFu) {
            const uint32_t bit  = 31u - __clz(node_group.y);
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t rank = __popc(node_group.y & 0xffu & ((1u << slot) - 1u));
            const char*    p    = reinterpret_cast<const char*>(nodes + kWideNodeWords * size_t(node_group.x + rank));
            prefetchL1(p);
            prefetchL1(p + 64);
        } else if (0 != tri_group.y) {
            prefetchL1(recs + 4 * size_t(tri_group.x + (31u - __clz(tri_group.y))));
        }
        return;
    }
    for (uint32_t m = tri_group.y; 0 != m; m &= m - 1u) {
        const float4* p = recs + 4 * size_t(tri_group.x + uint32_t(__ffs(int(m))) - 1u);
        if (4 == mode) {
            prefetchL1(p);
        } else {
            prefetchL2(p);
        }
    }
    if (3 == mode && node_group.y > 0x00FFFFFFu) {
        uint32_t m = node_group.y >> 24;
        m &= ~(1u << (31u - __clz(m)));  // the nearest child is read in the lane's next NODE step
        for (; 0 != m; m &= m - 1u) {
            const uint32_t slot = (uint32_t(__ffs(int(m))) - 1u) ^ octinv;
            const uint32_t rank = __popc(node_group.y & 0xffu & ((1u << slot) - 1u));
            const char*    p    = reinterpret_cast<const char*>(nodes + kWideNodeWords * size_t(node_group.x + rank));
            prefetchL2(p);
            prefetchL2(p + 64);
        }
    }
}

// ---- ray sort ---------------------------------------------------------------------------------------------------------------------------
// After the first bounce the rays of a warp start anywhere and point anywhere: they walk different instances, the lock-step loop runs
// half empty and every node comes from DRAM. A counting sort of the trace items by (origin cell, direction octant) — two passes with one
// counter per key, no host round trip: the queue lengths stay on the device — puts rays that walk the same part of the scene side by side.
// Hits are written per item and ties are resolved by ids (closerOrLater), so the order of the queue does not change any result.

struct TraceItems {
    uint32_t n, stride;
    bool     compact;
};

template <bool AnyHit>
__device__ __forceinline__ TraceItems traceItems(const PathState& st) {
    TraceItems t;
    t.stride             = st.shadow_stride;
    t.compact            = AnyHit && nullptr != st.queue_r;
    const uint64_t total = AnyHit ? (t.compact ? uint64_t(st.counters[10]) : uint64_t(st.counters[1]) * t.stride)
                                  : uint64_t(st.counters[st.lanes > 1 ? 7 : 0]);
    t.n                  = uint32_t(total < 0xFFFFFFFFull ? total : 0xFFFFFFFFull);
    return t;
}

template <bool AnyHit>
__device__ __forceinline__ bool traceItem(const PathState& st, const TraceItems& t, uint32_t i, uint32_t& item) {
    if (t.compact) {
        item = st.queue_r[i];
        return true;
    }
    if (AnyHit) {
        const uint32_t slot = st.queue_b[i / t.stride];
        const uint32_t k    = i % t.stride;
        item                = slot * t.stride + k;
        return k < st.sh_n[slot];
    }
    item = (st.lanes > 1 ? st.queue_t : st.queue_a)[i];
    return true;
}

__device__ __forceinline__ uint32_t spreadBits5(uint32_t v) {  // abcde -> 0a0b0c0d0e
    v = (v | (v << 4)) & 0x10Fu;
    v = (v | (v << 2)) & 0x133u;
    v = (v | (v << 1)) & 0x155u;
    return v;
}

// the lane's rank among the lanes of the warp that hold its key, their number, and the first of them (the one that touches the counter)
__device__ __forceinline__ uint32_t warpKeyRank(uint32_t key, uint32_t& leader, uint32_t& count) {
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    leader               = uint32_t(__ffs(int(peers))) - 1u;
    count                = uint32_t(__popc(peers));
    return uint32_t(__popc(peers & ((1u << (threadIdx.x & 31u)) - 1u)));
}

template <bool AnyHit>
__global__ void __launch_bounds__(256) sortKeyKernel(SceneDevice sc, PathState st, uint32_t mode) {
    const TraceItems t     = traceItems<AnyHit>(st);
    const uint32_t   round = (t.n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < round; i += gridDim.x * blockDim.x) {
        uint32_t key = 0xFFFFFFFFu, item = 0;
        if (i < t.n && traceItem<AnyHit>(st, t, i, item)) {
            V3 o, d;
            if (AnyHit) {
                const float4 so = st.sh_o[item], sp = st.sh_p[item];
                o = {so.x, so.y, so.z};
                if (0 != (__float_as_uint(sp.w) & 0x80000000u)) {
                    const float4 wi = st.sh_wi[item];
                    d = {wi.x, wi.y, wi.z};
                } else {
                    d = {sp.x - so.x, sp.y - so.y, sp.z - so.z};
                }
            } else {
                const float4 ro = st.ray_o[item], rd = st.ray_d[item];
                o = {ro.x, ro.y, ro.z};
                d = {rd.x, rd.y, rd.z};
            }
            const uint32_t cx   = uint32_t(fminf(fmaxf((o.x - sc.world_lo.x) * sc.world_cells.x, 0.f), 31.f));
            const uint32_t cy   = uint32_t(fminf(fmaxf((o.y - sc.world_lo.y) * sc.world_cells.y, 0.f), 7.f));
            const uint32_t cz   = uint32_t(fminf(fmaxf((o.z - sc.world_lo.z) * sc.world_cells.z, 0.f), 31.f));
            const uint32_t oct  = (d.x < 0.f ? 1u : 0u) | (d.y < 0.f ? 2u : 0u) | (d.z < 0.f ? 4u : 0u);
            const uint32_t cell = (cy << 10) | spreadBits5(cx) | (spreadBits5(cz) << 1);
            key                 = 2 == mode ? ((oct << 13) | cell) : ((cell << 3) | oct);
        }
        if (i < t.n) st.ml_count[i] = key;
        uint32_t       leader, count;
        const uint32_t rank = warpKeyRank(key, leader, count);
        if (0 == rank && 0xFFFFFFFFu != key) atomicAdd(st.sort_bins + key, count);
    }
}

// exclusive scan of the kSortBins counters in place; the total becomes the length of the sorted queue
__global__ void __launch_bounds__(1024) sortScanKernel(PathState st) {
    __shared__ uint32_t sums[1024];
    constexpr uint32_t  kPer  = kSortBins / 1024;
    uint32_t*           bins  = st.sort_bins + threadIdx.x * kPer;
    uint32_t            local = 0;
    for (uint32_t k = 0; k < kPer; ++k) local += bins[k];
    sums[threadIdx.x] = local;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        const uint32_t v = threadIdx.x >= o ? sums[threadIdx.x - o] : 0u;
        __syncthreads();
        sums[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t running = sums[threadIdx.x] - local;
    for (uint32_t k = 0; k < kPer; ++k) {
        const uint32_t c = bins[k];
        bins[k]          = running;
        running += c;
    }
    if (1023 == threadIdx.x) st.counters[15] = sums[1023];
}

template <bool AnyHit>
__global__ void __launch_bounds__(256) sortScatterKernel(PathState st) {
    const TraceItems t     = traceItems<AnyHit>(st);
    const uint32_t   round = (t.n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < round; i += gridDim.x * blockDim.x) {
        uint32_t key = 0xFFFFFFFFu, item = 0;
        if (i < t.n) {
            key = st.ml_count[i];
            if (0xFFFFFFFFu != key) traceItem<AnyHit>(st, t, i, item);
        }
        uint32_t       leader, count;
        const uint32_t rank = warpKeyRank(key, leader, count);
        uint32_t       base = 0;
        if (0 == rank && 0xFFFFFFFFu != key) base = atomicAdd(st.sort_bins + key, count);
        base = __shfl_sync(0xffffffffu, base, leader);
        if (0xFFFFFFFFu != key) st.queue_m[base + rank] = item;
    }
}

struct SceneStepTuning {
    uint32_t fetch_idle;  // refill when at least this many lanes are idle
    uint32_t weight[4];   // NODE, TRIANGLE, PROP, ENTER: the step kind with the largest (ready lanes x weight) runs
    uint32_t prefetch;    // 1: a lane asks L1 for the node / record it will read in its next step as soon as it knows which
    uint32_t sorted;      // 1: the items come from queue_m in ray-sort order (counters[15] of them)
    uint32_t repeat, repeat_lanes;  // ray-pool kernel: up to `repeat` steps of one kind per pick while `repeat_lanes` rays want another
    uint32_t debug_item;  // diagnostics (ZYGPU_DEBUG_TRACE_ITEM): the instrumented (zygpu_set_counting) one-ray-per-lane kernel prints what
                          // this trace item does; kEnd = off. Only the counting instances carry the printf (it costs 200 bytes of stack).
};

template <bool AnyHit, bool Count, int MinBlocks>
__global__ void __launch_bounds__(128, MinBlocks) sceneTracePersistent(SceneDevice sc, PathState st, uint32_t* __restrict__ work_counter,
                                                            SceneStepTuning tune, unsigned long long* __restrict__ tally) {
    constexpr uint32_t kFull = 0xffffffffu;
    const uint32_t     lane  = threadIdx.x & 31u;

    // the trace queue: closest-hit rays are vertex ids, shadow rays are records (compact list, or `stride` slots per path)
    TraceItems items = traceItems<AnyHit>(st);
    if (0 != tune.sorted) items.n = st.counters[15];
    const uint32_t n = items.n;

    const uint32_t warps      = gridDim.x * (blockDim.x / 32u);
    const uint32_t pool_items = max(32u, min(kScenePoolItems, (n / (warps * 4u)) & ~31u));

    uint32_t pool_next = 0, pool_end = 0;
    bool     exhausted = false;

    bool     has_ray = false;
    bool     in_mesh = false;
    uint32_t item    = 0;
    uint32_t depth_surface = 0;
    float    tmax0      = 0.f;   // the max_t the ray started with: the limit of the reference's box gates (gateBox)
    uint32_t enter_prop = kEnd;  // a mesh prop that passed the culling tests and waits for its ENTER step
    uint32_t cur_prop = 0, hit_prop = kEnd;
    bool     occluded = false;
    const float4* nodes = sc.tlas_nodes;  // of the level the lane is in
    const float4* recs  = sc.tlas_recs;

    WideRay  w;
    uint2    stack[kWideStack];
    uint32_t sp = 0, sp_mesh = 0;  // sp_mesh: stack depth at mesh entry (the world ray and the prop-tree entries lie below)
    uint2    node_group = make_uint2(0u, 0u);
    uint2    tri_group  = make_uint2(0u, 0u);
    float    hu = 0.f, hv = 0.f;
    uint32_t primitive = 0;
    uint32_t traced = 0, count_nodes = 0, count_tris = 0, count_props = 0;
    uint32_t steps[3] = {0, 0, 0};  // instrumented build: warp-level NODE / TRIANGLE / PROP + ENTER steps

    for (;;) {
        // ---- refill idle lanes
        uint32_t idle = __ballot_sync(kFull, !has_ray);
        while (0 != idle && !exhausted) {
            if (pool_next >= pool_end) {
                uint32_t base = 0;
                if (0 == lane) base = atomicAdd(work_counter, pool_items);
                base = __shfl_sync(kFull, base, 0);
                if (base >= n) {
                    exhausted = true;
                    break;
                }
                pool_next = base;
                pool_end  = min(base + pool_items, n);
            }
            const uint32_t avail = pool_end - pool_next;
            const uint32_t rank  = __popc(idle & ((1u << lane) - 1u));
            if (!has_ray && rank < avail) {
                const uint32_t i     = pool_next + rank;
                bool           valid = true;
                if (0 != tune.sorted) {
                    item = st.queue_m[i];
                } else {
                    valid = traceItem<AnyHit>(st, items, i, item);
                }
                if (valid) {
                    uint32_t flags = 0;
                    w.ray          = loadTraceRay<AnyHit>(st, item, depth_surface, &flags);
                    if (!AnyHit) clipToMedium(sc, st, item, flags, w.ray);
                    tmax0 = w.ray.tmax;
                    setupWideRay(w);
                    has_ray    = true;
                    in_mesh    = false;
                    enter_prop = kEnd;
                    hit_prop   = kEnd;
                    occluded   = false;
                    nodes      = sc.tlas_nodes;
                    recs       = sc.tlas_recs;
                    sp         = 0;
                    sp_mesh    = 0;
                    node_group = make_uint2(0u, 0 != sc.num_solid_nodes ? 0x80000000u : 0u);  // root of the prop tree
                    tri_group  = make_uint2(0u, 0u);
                    hu = hv    = 0.f;
                    primitive  = 0;
                    traced += 1;
                }
            }
            pool_next += min(avail, (uint32_t)__popc(idle));
            idle = __ballot_sync(kFull, !has_ray);
        }
        if (kFull == idle) break;

        // ---- lock-step steps until enough lanes ran out of work
        for (;;) {
            const bool     ready_enter = has_ray && kEnd != enter_prop;
            const bool     ready_node  = has_ray && !ready_enter && node_group.y > 0x00FFFFFFu;
            const bool     ready_leaf  = has_ray && !ready_enter && 0 != tri_group.y;
            const uint32_t cn = __popc(__ballot_sync(kFull, ready_node)) * tune.weight[0];
            const uint32_t ct = __popc(__ballot_sync(kFull, ready_leaf && in_mesh)) * tune.weight[1];
            const uint32_t cp = __popc(__ballot_sync(kFull, ready_leaf && !in_mesh)) * tune.weight[2];
            const uint32_t ce = __popc(__ballot_sync(kFull, ready_enter)) * tune.weight[3];
            const uint32_t most = max(max(cn, ct), max(cp, ce));

            if (0 != ce && ce == most) {
                // ENTER step: the ray goes to the object space of the mesh (triangle_tree.zig:49: t is shared). The world ray and the
                // prop-tree work still pending stay on the stack below the mesh's entries.
                if (Count) steps[2] += 1;
                if (ready_enter) {
                    if (node_group.y > 0x00FFFFFFu) stack[sp++] = node_group;
                    if (0 != tri_group.y) stack[sp++] = tri_group;
                    stack[sp++] = make_uint2(__float_as_uint(w.ray.o.x), __float_as_uint(w.ray.o.y));
                    stack[sp++] = make_uint2(__float_as_uint(w.ray.o.z), __float_as_uint(w.ray.d.x));
                    stack[sp++] = make_uint2(__float_as_uint(w.ray.d.y), __float_as_uint(w.ray.d.z));
                    stack[sp++] = make_uint2(__float_as_uint(w.ray.inv_d.x), __float_as_uint(w.ray.inv_d.y));
                    stack[sp++] = make_uint2(__float_as_uint(w.ray.inv_d.z), 0u);
                    sp_mesh     = sp;
                    const TrafoD      trafo = loadTrafo(sc.trafos, enter_prop);
                    const MeshDevice* m     = sc.meshes + sc.props[enter_prop].mesh;
                    nodes                   = m->wide_nodes;
                    recs                    = m->wide_tris;
                    w.ray                   = worldToObjectRay(trafo, w.ray);
                    setupWideRay(w);
                    if (Count && !AnyHit && item == tune.debug_item) {
                        printf("[trace] ENTER prop %u mesh %u sp %u | world o %.9g %.9g %.9g d %.9g %.9g %.9g tmax %.9g | object o %.9g %.9g %.9g d %.9g %.9g %.9g\n",
                               enter_prop, sc.props[enter_prop].mesh, sp, __uint_as_float(stack[sp - 5].x), __uint_as_float(stack[sp - 5].y),
                               __uint_as_float(stack[sp - 4].x), __uint_as_float(stack[sp - 4].y), __uint_as_float(stack[sp - 3].x),
                               __uint_as_float(stack[sp - 3].y), w.ray.tmax, w.ray.o.x, w.ray.o.y, w.ray.o.z, w.ray.d.x, w.ray.d.y, w.ray.d.z);
                    }
                    cur_prop   = enter_prop;
                    enter_prop = kEnd;
                    in_mesh    = true;
                    node_group = make_uint2(0u, 0x80000000u);
                    tri_group  = make_uint2(0u, 0u);
                }
            } else if (0 != cp && cp == most) {
                // PROP step: a lane works through its pending prop records until a mesh prop survives the culling tests
                if (Count) steps[2] += 1;
                if (ready_leaf && !in_mesh) {
                    do {
                        const uint32_t bit = 31u - __clz(tri_group.y);
                        tri_group.y &= ~(1u << bit);
                        if (Count) count_props += 1;
                        const float4* rp = recs + 4 * size_t(tri_group.x + bit);
                        const F8      rr = ldg256(rp);
                        const float4  r0 = rr.lo, r1 = rr.hi;
                        const uint32_t  p    = __float_as_uint(r0.w);
                        const ZygpuProp prop = sc.props[p];
                        // the reference reaches a prop through its leaf's box (prop_tree.zig:86-104) ...
                        bool enter = gateBox(make_float4(r0.x, r0.y, r0.z, 0.f), make_float4(r1.x, r1.y, r1.z, 0.f), w.ray, tmax0);
                        // ... then Prop.intersect / Prop.visibility: flags, world box (prop.zig:176-183, 212-218)
                        enter = enter && (AnyHit ? 0 != (prop.flags & ZYG_PROP_VISIBLE_IN_SHADOW) : propVisible(prop.flags, depth_surface));
                        enter = enter && gateBox(__ldg(sc.aabbs + 2 * size_t(p)), __ldg(sc.aabbs + 2 * size_t(p) + 1), w.ray, tmax0);
                        if (!enter) continue;
                        if (ZYG_SHAPE_TRIANGLE_MESH == prop.shape) {
                            // culling only: every triangle of the instance lies inside its bounding sphere
                            if (segmentMeetsSphere(w.ray, __ldg(rp + 2))) {
                                enter_prop = p;
                                break;
                            }
                            continue;
                        }
                        const TrafoD trafo = loadTrafo(sc.trafos, p);
                        if (AnyHit) {
                            bool hit = false;
                            HitD unused;
                            switch (prop.shape) {
                                case ZYG_SHAPE_CUBE: hit = cubeIntersectP(w.ray, trafo); break;
                                case ZYG_SHAPE_RECTANGLE: hit = rectangleIntersect(w.ray, trafo, unused); break;
                                case ZYG_SHAPE_DISK: hit = diskIntersect(w.ray, trafo, unused); break;
                                case ZYG_SHAPE_SPHERE: hit = sphereIntersect(w.ray, trafo, unused); break;
                                default: break;
                            }
                            if (hit) {
                                occluded     = true;
                                sp           = 0;
                                node_group.y = 0;
                                tri_group.y  = 0;
                            }
                        } else {
                            HitD h;
                            bool hit = false;
                            switch (prop.shape) {
                                case ZYG_SHAPE_CUBE: hit = cubeIntersect(w.ray, trafo, h); break;
                                case ZYG_SHAPE_RECTANGLE: hit = rectangleIntersect(w.ray, trafo, h); break;
                                case ZYG_SHAPE_DISK: hit = diskIntersect(w.ray, trafo, h); break;
                                case ZYG_SHAPE_SPHERE: hit = sphereIntersect(w.ray, trafo, h); break;
                                default: break;
                            }
                            if (hit && closerOrLater(h.t, w.ray.tmax, p, h.primitive, hit_prop, primitive)) {
                                w.ray.tmax = h.t;
                                hu         = h.u;
                                hv         = h.v;
                                primitive  = h.primitive;
                                hit_prop   = p;
                            }
                        }
                    } while (0 != tri_group.y);
                }
            } else if (0 != ct && ct == most) {
                // TRIANGLE step
                if (Count) steps[1] += 1;
                if (ready_leaf && in_mesh) {
                    const uint32_t bit = 31u - __clz(tri_group.y);
                    tri_group.y &= ~(1u << bit);
                    if (Count) count_tris += 1;
                    MeshDevice mesh;
                    mesh.wide_tris = recs;
                    float    t, u, v;
                    uint32_t prim;
                    if (testWideTriangle(mesh, w.ray, tmax0, tri_group.x + bit, t, u, v, prim)) {
                        if (AnyHit) {
                            occluded     = true;
                            in_mesh      = false;
                            sp           = 0;
                            node_group.y = 0;
                            tri_group.y  = 0;
                        } else if (closerOrLater(t, w.ray.tmax, cur_prop, prim, hit_prop, primitive)) {
                            if (Count && item == tune.debug_item) {
                                printf("[trace] HIT prop %u record %u prim %u t %.9g (max_t was %.9g) u %.9g v %.9g sp %u sp_mesh %u\n", cur_prop,
                                       tri_group.x + bit, prim, t, w.ray.tmax, u, v, sp, sp_mesh);
                            }
                            w.ray.tmax = t;  // probe.ray.max_t = isec.t, prop_tree.zig:77
                            hu         = u;
                            hv         = v;
                            primitive  = prim;
                            hit_prop   = cur_prop;
                        }
                    }
                }
            } else {
                // NODE step: the same code for a lane in the prop tree and a lane inside a mesh
                if (Count && 0 != cn) steps[0] += 1;
                if (ready_node) {
                    const uint32_t hits  = node_group.y;
                    const uint32_t gmask = hits & 0xffu;
                    const uint32_t bit   = 31u - __clz(hits);
                    node_group.y         = hits & ~(1u << bit);
                    const uint32_t slot  = (bit - 24u) ^ w.octinv;
                    const uint32_t rank  = __popc(gmask & ((1u << slot) - 1u));
                    const uint32_t node_index = node_group.x + rank;
                    if (node_group.y > 0x00FFFFFFu) stack[sp++] = node_group;
                    if (0 != tri_group.y) stack[sp++] = tri_group;
                    if (Count) count_nodes += 1;

                    const WideNodeRegs nd = loadWideNode(nodes, node_index);
                    const float4 n0 = nd.n0, n1 = nd.n1, n2 = nd.n2, n3 = nd.n3, n4 = nd.n4;

                    const uint32_t hitmask = testWideNode(w, w.ray.tmin, cullLimit(w.ray), n0, n1, n2, n3, n4);

                    node_group.x = __float_as_uint(n1.x);
                    node_group.y = (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
                    tri_group.x  = __float_as_uint(n1.y);
                    tri_group.y  = hitmask & 0x00FFFFFFu;
                    if (0 != tune.prefetch) prefetchNext(tune.prefetch, nodes, recs, node_group, tri_group, w.octinv);
                }
            }

            // ---- lanes that ran dry pop their stack, leave the mesh or retire their ray
            if (has_ray && kEnd == enter_prop && node_group.y <= 0x00FFFFFFu && 0 == tri_group.y) {
                if (in_mesh && sp == sp_mesh) {
                    // back to the prop tree: the world ray comes off the stack, max_t is the one found so far
                    in_mesh        = false;
                    const uint2 e4 = stack[--sp], e3 = stack[--sp], e2 = stack[--sp], e1 = stack[--sp], e0 = stack[--sp];
                    w.ray.o        = {__uint_as_float(e0.x), __uint_as_float(e0.y), __uint_as_float(e1.x)};
                    w.ray.d        = {__uint_as_float(e1.y), __uint_as_float(e2.x), __uint_as_float(e2.y)};
                    w.ray.inv_d    = {__uint_as_float(e3.x), __uint_as_float(e3.y), __uint_as_float(e4.x)};
                    if (Count && !AnyHit && item == tune.debug_item) {
                        printf("[trace] LEAVE prop %u sp %u | world o %.9g %.9g %.9g d %.9g %.9g %.9g tmax %.9g\n", cur_prop, sp, w.ray.o.x, w.ray.o.y,
                               w.ray.o.z, w.ray.d.x, w.ray.d.y, w.ray.d.z, w.ray.tmax);
                    }
                    setupWideRay(w);
                    nodes = sc.tlas_nodes;
                    recs  = sc.tlas_recs;
                }
                if (0 == sp) {
                    if (AnyHit) {
                        st.sh_wi[item].w = occluded ? 0.f : 1.f;
                    } else {
                        st.ray_d[item].w = w.ray.tmax;
                        st.hit[item]     = make_float4(hu, hv, __uint_as_float(primitive), __uint_as_float(hit_prop));
                    }
                    has_ray = false;
                } else if (!in_mesh || sp > sp_mesh) {
                    const uint2 e = stack[--sp];
                    if (e.y > 0x00FFFFFFu) {
                        node_group = e;
                    } else {
                        tri_group = e;
                    }
                    if (1 == tune.prefetch) prefetchNext(1, nodes, recs, node_group, tri_group, w.octinv);
                }
            }

            const uint32_t active = __ballot_sync(kFull, has_ray);
            if (0 == active) break;
            if (!exhausted && 32u - __popc(active) >= tune.fetch_idle) break;
        }
    }

    for (int o = 16; o > 0; o >>= 1) traced += __shfl_down_sync(kFull, traced, o);
    if (0 == lane && 0 != traced) atomicAdd(&st.counters[AnyHit ? 6 : 5], traced);
    if (Count) {
        unsigned long long cnt[3] = {count_nodes, count_tris, count_props};
        for (int k = 0; k < 3; ++k) {
            for (int o = 16; o > 0; o >>= 1) cnt[k] += __shfl_down_sync(kFull, cnt[k], o);
            if (0 == lane) atomicAdd(tally + (AnyHit ? 6 : 0) + k, cnt[k]);
        }
        if (0 == lane) {
            for (int k = 0; k < 3; ++k) atomicAdd(tally + (AnyHit ? 6 : 0) + 3 + k, (unsigned long long)steps[k]);
        }
    }
}

// ---- ray-pool variant of the fused kernel -------------------------------------------------------------------------------------------------
// sceneTracePersistent keeps one ray per lane, so a step only uses the lanes whose ray wants that kind of step: 14 of 32 on the instanced
// scene with its four kinds (profiles/r02_d_config3_sceneTrace_64reg.md). Here a warp keeps kScenePoolSlots rays in shared memory (the ray
// in its current space, traversal groups, hit so far, level bookkeeping: 100 bytes per ray) and every step first hands the rays that are
// ready for the chosen kind to the lanes, like traceWidePool does for ray batches (device/trace.cu). Per-ray stacks live in a global scratch
// array. Same tests per ray, ties resolved by ids: the results are sceneTracePersistent's, bit for bit.
constexpr uint32_t kScenePoolSlots = 64;

struct ScenePool {  // one per warp
    float4   a[kScenePoolSlots];   // origin xyz (world, or object space inside a mesh) | the max_t the ray started with
    float4   b[kScenePoolSlots];   // direction xyz | max_t (the closest hit so far)
    float4   c[kScenePoolSlots];   // inverse direction xyz | u of the hit
    uint4    g[kScenePoolSlots];   // node group | triangle / prop-record group
    uint4    h[kScenePoolSlots];   // stack depth | trace item (kEnd: the slot is free) | primitive (any hit: 1 = occluded) | v of the hit
    uint4    s[kScenePoolSlots];   // stack depth at mesh entry | prop the ray is inside of (kEnd: prop tree) | prop waiting for ENTER | prop of the hit
    uint32_t dm[kScenePoolSlots];  // depth.surface of the ray (bits 0-7) | mesh of the prop the ray is inside of (bits 8-31)
    uint32_t ready[kScenePoolSlots];  // what the ray's next step is ready for: 1 NODE, 2 TRIANGLE, 4 PROP, 8 ENTER; 0: the slot is free
    uint32_t assign[32];
};

template <bool AnyHit>
__global__ void __launch_bounds__(128, 7) scenePoolTrace(SceneDevice sc, PathState st, uint32_t* __restrict__ work_counter, SceneStepTuning tune,
                                                         uint2* __restrict__ stacks) {
    constexpr uint32_t kFull = 0xffffffffu;
    __shared__ ScenePool pools[4];
    ScenePool&           pool = pools[threadIdx.x >> 5];
    const uint32_t       lane = threadIdx.x & 31u;
    const uint32_t       lt   = (1u << lane) - 1u;
    uint2* __restrict__  stk  = stacks + size_t(blockIdx.x * 4u + (threadIdx.x >> 5)) * kScenePoolSlots * kScenePoolStack;

    TraceItems items = traceItems<AnyHit>(st);
    if (0 != tune.sorted) items.n = st.counters[15];
    const uint32_t n = items.n;

    const uint32_t warps      = gridDim.x * (blockDim.x / 32u);
    const uint32_t pool_items = max(32u, min(kScenePoolItems, (n / (warps * 4u)) & ~31u));

    for (uint32_t s = lane; s < kScenePoolSlots; s += 32) {
        pool.g[s] = make_uint4(0u, 0u, 0u, 0u);
        pool.h[s] = make_uint4(0u, kEnd, 0u, 0u);
        pool.s[s] = make_uint4(0u, kEnd, kEnd, kEnd);
        pool.ready[s] = 0u;
    }
    __syncwarp();

    uint32_t occupied  = 0;  // warp-uniform
    uint32_t pool_next = 0, pool_end = 0;
    bool     exhausted = false;
    uint32_t traced    = 0;

    for (;;) {
        // ---- refill free slots
        if (!exhausted && kScenePoolSlots - occupied >= tune.fetch_idle) {
            for (uint32_t half = 0; half < kScenePoolSlots / 32; ++half) {
                const uint32_t slot = lane + 32u * half;
                uint32_t       need = __ballot_sync(kFull, 0u == pool.ready[slot]);
                while (0 != need && !exhausted) {
                    if (pool_next >= pool_end) {
                        uint32_t base = 0;
                        if (0 == lane) base = atomicAdd(work_counter, pool_items);
                        base = __shfl_sync(kFull, base, 0);
                        if (base >= n) {
                            exhausted = true;
                            break;
                        }
                        pool_next = base;
                        pool_end  = min(base + pool_items, n);
                    }
                    const uint32_t avail = pool_end - pool_next;
                    const bool     take  = 0 != ((need >> lane) & 1u) && uint32_t(__popc(need & lt)) < avail;
                    bool           valid = false;
                    if (take) {
                        const uint32_t i    = pool_next + uint32_t(__popc(need & lt));
                        uint32_t       item = 0;
                        if (0 != tune.sorted) {
                            item  = st.queue_m[i];
                            valid = true;
                        } else {
                            valid = traceItem<AnyHit>(st, items, i, item);
                        }
                        if (valid) {
                            uint32_t depth_surface = 0, flags = 0;
                            RayT     r             = loadTraceRay<AnyHit>(st, item, depth_surface, &flags);
                            if (!AnyHit) clipToMedium(sc, st, item, flags, r);
                            pool.a[slot]  = make_float4(r.o.x, r.o.y, r.o.z, r.tmax);
                            pool.b[slot]  = make_float4(r.d.x, r.d.y, r.d.z, r.tmax);
                            pool.c[slot]  = make_float4(r.inv_d.x, r.inv_d.y, r.inv_d.z, 0.f);
                            pool.g[slot]  = make_uint4(0u, 0x80000000u, 0u, 0u);  // root of the prop tree
                            pool.h[slot]  = make_uint4(0u, item, 0u, 0u);
                            pool.s[slot]  = make_uint4(0u, kEnd, kEnd, kEnd);
                            pool.dm[slot] = depth_surface & 0xffu;
                            pool.ready[slot] = 1u;
                            traced += 1;
                        }
                    }
                    const uint32_t taken = __ballot_sync(kFull, take);
                    pool_next += __popc(taken);
                    occupied += __popc(__ballot_sync(kFull, valid));
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
        uint32_t m_node[kScenePoolSlots / 32], m_tri[kScenePoolSlots / 32], m_prop[kScenePoolSlots / 32], m_enter[kScenePoolSlots / 32];
        uint32_t cn = 0, ct = 0, cp = 0, ce = 0;
#pragma unroll
        for (uint32_t half = 0; half < kScenePoolSlots / 32; ++half) {
            const uint32_t r = pool.ready[lane + 32u * half];
            m_node[half]     = __ballot_sync(kFull, 0 != (r & 1u));
            m_tri[half]      = __ballot_sync(kFull, 0 != (r & 2u));
            m_prop[half]     = __ballot_sync(kFull, 0 != (r & 4u));
            m_enter[half]    = __ballot_sync(kFull, 0 != (r & 8u));
            cn += __popc(m_node[half]);
            ct += __popc(m_tri[half]);
            cp += __popc(m_prop[half]);
            ce += __popc(m_enter[half]);
        }
        const uint32_t wn = min(cn, 32u) * tune.weight[0], wt = min(ct, 32u) * tune.weight[1], wp = min(cp, 32u) * tune.weight[2],
                       we = min(ce, 32u) * tune.weight[3];
        const uint32_t most = max(max(wn, wt), max(wp, we));
        const uint32_t kind = (0 != we && we == most) ? 3u : ((0 != wp && wp == most) ? 2u : ((0 != wt && wt == most) ? 1u : 0u));

        // ---- hand the first 32 ready rays to the lanes
        uint32_t rank = 0;
#pragma unroll
        for (uint32_t half = 0; half < kScenePoolSlots / 32; ++half) {
            const uint32_t m = 3u == kind ? m_enter[half] : (2u == kind ? m_prop[half] : (1u == kind ? m_tri[half] : m_node[half]));
            if (0 != ((m >> lane) & 1u)) {
                const uint32_t r = rank + uint32_t(__popc(m & lt));
                if (r < 32u) pool.assign[r] = lane + 32u * half;
            }
            rank += __popc(m);
        }
        __syncwarp();
        const bool     busy = lane < rank;
        const uint32_t slot = busy ? pool.assign[lane] : 0u;

        // The rays stay in the lanes' registers while most of them want another step of the same kind (a NODE step usually leaves a ray
        // with the next node group, a TRIANGLE step with the rest of its leaf): picking and moving rays costs about half a step.
        uint32_t retired = 0;
        // (declared for all lanes: the repetition vote below is warp-wide)
        WideRay  w;
        float    tmax0 = 0.f, hu = 0.f;
        uint2*   stack = stk;
        uint2    node_group = make_uint2(0u, 0u), tri_group = make_uint2(0u, 0u);
        uint32_t sp = 0, dm = 0;
        bool     in_mesh = false, ray_dirty = false, hit_dirty = false, level_dirty = false;
        uint4    h = make_uint4(0u, kEnd, 0u, 0u), sv = make_uint4(0u, kEnd, kEnd, kEnd);
        uint32_t ready = 0;
        if (busy) {
            const float4 ra = pool.a[slot];
            const float4 rb = pool.b[slot];
            const float4 rc = pool.c[slot];
            const uint4  g  = pool.g[slot];
            h               = pool.h[slot];
            sv              = pool.s[slot];
            dm              = pool.dm[slot];

            w.ray.o     = {ra.x, ra.y, ra.z};
            w.ray.tmin  = 0.f;
            w.ray.d     = {rb.x, rb.y, rb.z};
            w.ray.tmax  = rb.w;
            w.ray.inv_d = {rc.x, rc.y, rc.z};
            tmax0 = ra.w;
            hu    = rc.w;
            stack = stk + size_t(slot) * kScenePoolStack;

            node_group = make_uint2(g.x, g.y);
            tri_group  = make_uint2(g.z, g.w);
            sp         = h.x;
            in_mesh    = kEnd != sv.y;
            // what has to go back to shared memory: the groups and the stack depth always, the rest when it changed

        }
        bool part = busy;
        for (uint32_t rep = 0;; ++rep) {
        if (part) {
            setupWideRay(w);
            const float4* nodes = sc.tlas_nodes;
            const float4* recs  = sc.tlas_recs;
            if (in_mesh) {
                const MeshDevice* m = sc.meshes + (dm >> 8);
                nodes               = m->wide_nodes;
                recs                = m->wide_tris;
            }

            if (3u == kind) {
                // ENTER: the world ray and the prop-tree work still pending stay on the stack below the mesh's entries
                const uint32_t enter_prop = sv.z;
                if (node_group.y > 0x00FFFFFFu) stack[sp++] = node_group;
                if (0 != tri_group.y) stack[sp++] = tri_group;
                stack[sp++] = make_uint2(__float_as_uint(w.ray.o.x), __float_as_uint(w.ray.o.y));
                stack[sp++] = make_uint2(__float_as_uint(w.ray.o.z), __float_as_uint(w.ray.d.x));
                stack[sp++] = make_uint2(__float_as_uint(w.ray.d.y), __float_as_uint(w.ray.d.z));
                stack[sp++] = make_uint2(__float_as_uint(w.ray.inv_d.x), __float_as_uint(w.ray.inv_d.y));
                stack[sp++] = make_uint2(__float_as_uint(w.ray.inv_d.z), 0u);
                sv.x        = sp;
                const TrafoD   trafo = loadTrafo(sc.trafos, enter_prop);
                const uint32_t mesh  = sc.props[enter_prop].mesh;
                w.ray                = worldToObjectRay(trafo, w.ray);
                sv.y                 = enter_prop;
                sv.z                 = kEnd;
                dm                   = (dm & 0xffu) | (mesh << 8);
                in_mesh              = true;
                ray_dirty = level_dirty = true;
                node_group           = make_uint2(0u, 0x80000000u);
                tri_group            = make_uint2(0u, 0u);
            } else if (2u == kind) {
                // PROP: the ray works through its pending prop records until a mesh prop survives the culling tests
                do {
                    const uint32_t bit = 31u - __clz(tri_group.y);
                    tri_group.y &= ~(1u << bit);
                    const float4* rp = recs + 4 * size_t(tri_group.x + bit);
                    const F8      rr = ldg256(rp);
                    const float4  r0 = rr.lo, r1 = rr.hi;
                    const uint32_t  p    = __float_as_uint(r0.w);
                    const ZygpuProp prop = sc.props[p];
                    bool enter = gateBox(make_float4(r0.x, r0.y, r0.z, 0.f), make_float4(r1.x, r1.y, r1.z, 0.f), w.ray, tmax0);
                    enter = enter && (AnyHit ? 0 != (prop.flags & ZYG_PROP_VISIBLE_IN_SHADOW) : propVisible(prop.flags, dm & 0xffu));
                    enter = enter && gateBox(__ldg(sc.aabbs + 2 * size_t(p)), __ldg(sc.aabbs + 2 * size_t(p) + 1), w.ray, tmax0);
                    if (!enter) continue;
                    if (ZYG_SHAPE_TRIANGLE_MESH == prop.shape) {
                        if (segmentMeetsSphere(w.ray, __ldg(rp + 2))) {
                            sv.z        = p;
                            level_dirty = true;
                            break;
                        }
                        continue;
                    }
                    const TrafoD trafo = loadTrafo(sc.trafos, p);
                    if (AnyHit) {
                        bool hit = false;
                        HitD unused;
                        switch (prop.shape) {
                            case ZYG_SHAPE_CUBE: hit = cubeIntersectP(w.ray, trafo); break;
                            case ZYG_SHAPE_RECTANGLE: hit = rectangleIntersect(w.ray, trafo, unused); break;
                                case ZYG_SHAPE_DISK: hit = diskIntersect(w.ray, trafo, unused); break;
                            case ZYG_SHAPE_SPHERE: hit = sphereIntersect(w.ray, trafo, unused); break;
                            default: break;
                        }
                        if (hit) {
                            h.z          = 1u;
                            hit_dirty    = true;
                            sp           = 0;
                            node_group.y = 0;
                            tri_group.y  = 0;
                        }
                    } else {
                        HitD hd;
                        bool hit = false;
                        switch (prop.shape) {
                            case ZYG_SHAPE_CUBE: hit = cubeIntersect(w.ray, trafo, hd); break;
                            case ZYG_SHAPE_RECTANGLE: hit = rectangleIntersect(w.ray, trafo, hd); break;
                            case ZYG_SHAPE_DISK: hit = diskIntersect(w.ray, trafo, hd); break;
                            case ZYG_SHAPE_SPHERE: hit = sphereIntersect(w.ray, trafo, hd); break;
                            default: break;
                        }
                        if (hit && closerOrLater(hd.t, w.ray.tmax, p, hd.primitive, sv.w, h.z)) {
                            w.ray.tmax = hd.t;
                            hu         = hd.u;
                            h.w        = __float_as_uint(hd.v);
                            h.z        = hd.primitive;
                            sv.w       = p;
                            hit_dirty = level_dirty = true;
                        }
                    }
                } while (0 != tri_group.y);
            } else if (1u == kind) {
                // TRIANGLE
                const uint32_t bit = 31u - __clz(tri_group.y);
                tri_group.y &= ~(1u << bit);
                MeshDevice mesh;
                mesh.wide_tris = recs;
                float    t, u, v;
                uint32_t prim;
                if (testWideTriangle(mesh, w.ray, tmax0, tri_group.x + bit, t, u, v, prim)) {
                    if (AnyHit) {
                        h.z          = 1u;
                        hit_dirty = level_dirty = true;
                        in_mesh      = false;
                        sv.y         = kEnd;
                        sp           = 0;
                        node_group.y = 0;
                        tri_group.y  = 0;
                    } else if (closerOrLater(t, w.ray.tmax, sv.y, prim, sv.w, h.z)) {
                        w.ray.tmax = t;
                        hu         = u;
                        h.w        = __float_as_uint(v);
                        h.z        = prim;
                        sv.w       = sv.y;
                        hit_dirty = level_dirty = true;
                    }
                }
            } else {
                // NODE: the same code for a ray in the prop tree and a ray inside a mesh
                const uint32_t hits  = node_group.y;
                const uint32_t gmask = hits & 0xffu;
                const uint32_t bit   = 31u - __clz(hits);
                node_group.y         = hits & ~(1u << bit);
                const uint32_t cslot = (bit - 24u) ^ w.octinv;
                const uint32_t crank = __popc(gmask & ((1u << cslot) - 1u));
                const uint32_t node_index = node_group.x + crank;
                if (node_group.y > 0x00FFFFFFu) stack[sp++] = node_group;
                if (0 != tri_group.y) stack[sp++] = tri_group;

                const WideNodeRegs nd = loadWideNode(nodes, node_index);
                const float4 n0 = nd.n0, n1 = nd.n1, n2 = nd.n2, n3 = nd.n3, n4 = nd.n4;
                const uint32_t hitmask = testWideNode(w, w.ray.tmin, cullLimit(w.ray), n0, n1, n2, n3, n4);

                node_group.x = __float_as_uint(n1.x);
                node_group.y = (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
                tri_group.x  = __float_as_uint(n1.y);
                tri_group.y  = hitmask & 0x00FFFFFFu;
            }

            // ---- a ray that ran dry pops its stack, leaves the mesh or retires
            if (kEnd == sv.z && node_group.y <= 0x00FFFFFFu && 0 == tri_group.y) {
                if (in_mesh && sp == sv.x) {
                    in_mesh        = false;
                    sv.y           = kEnd;
                    ray_dirty = level_dirty = true;
                    const uint2 e4 = stack[--sp], e3 = stack[--sp], e2 = stack[--sp], e1 = stack[--sp], e0 = stack[--sp];
                    w.ray.o        = {__uint_as_float(e0.x), __uint_as_float(e0.y), __uint_as_float(e1.x)};
                    w.ray.d        = {__uint_as_float(e1.y), __uint_as_float(e2.x), __uint_as_float(e2.y)};
                    w.ray.inv_d    = {__uint_as_float(e3.x), __uint_as_float(e3.y), __uint_as_float(e4.x)};
                }
                if (0 == sp) {
                    const uint32_t item = h.y;
                    if (AnyHit) {
                        st.sh_wi[item].w = 0 != h.z ? 0.f : 1.f;
                    } else {
                        st.ray_d[item].w = w.ray.tmax;
                        st.hit[item]     = make_float4(hu, __uint_as_float(h.w), __uint_as_float(h.z), __uint_as_float(sv.w));
                    }
                    h.y     = kEnd;
                    retired = 1;
                } else if (!in_mesh || sp > sv.x) {
                    const uint2 e = stack[--sp];
                    if (e.y > 0x00FFFFFFu) {
                        node_group = e;
                    } else {
                        tri_group = e;
                    }
                }
            }
            ready = 0;
            if (0 == retired) {
                if (kEnd != sv.z) {
                    ready = 8u;
                } else {
                    ready = (node_group.y > 0x00FFFFFFu ? 1u : 0u) | (0 != tri_group.y ? (in_mesh ? 2u : 4u) : 0u);
                }
            }
        }
        // another step of the same kind while enough of the lanes' rays are ready for one
        const bool     again = part && 0 != (ready & (1u << kind));
        const uint32_t more  = __ballot_sync(kFull, again);
        if (rep + 1u >= tune.repeat || uint32_t(__popc(more)) < tune.repeat_lanes) break;
        part = again;
        }
        if (busy) {
            if (ray_dirty) {
                pool.a[slot]  = make_float4(w.ray.o.x, w.ray.o.y, w.ray.o.z, tmax0);
                pool.b[slot]  = make_float4(w.ray.d.x, w.ray.d.y, w.ray.d.z, w.ray.tmax);
                pool.c[slot]  = make_float4(w.ray.inv_d.x, w.ray.inv_d.y, w.ray.inv_d.z, hu);
                pool.dm[slot] = dm;
            } else if (hit_dirty) {
                pool.b[slot].w = w.ray.tmax;
                pool.c[slot].w = hu;
            }
            pool.g[slot] = make_uint4(node_group.x, node_group.y, tri_group.x, tri_group.y);
            if (hit_dirty || 0 != retired) {
                h.x          = sp;
                pool.h[slot] = h;
            } else {
                pool.h[slot].x = sp;
            }
            if (level_dirty) pool.s[slot] = sv;
            pool.ready[slot] = ready;
        }
        occupied -= __popc(__ballot_sync(kFull, 0 != retired));
        __syncwarp();
    }

    for (int o = 16; o > 0; o >>= 1) traced += __shfl_down_sync(kFull, traced, o);
    if (0 == lane && 0 != traced) atomicAdd(&st.counters[AnyHit ? 6 : 5], traced);
}


// Context.nextEvent -> Scene.intersect, context.zig:54-69, scene.zig:225-227
__global__ void __launch_bounds__(kBlock) extendKernel(SceneDevice sc, PathState st) {
    const uint32_t count = st.counters[st.lanes > 1 ? 7 : 0];
    const uint32_t* __restrict__ queue = st.lanes > 1 ? st.queue_t : st.queue_a;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t slot = queue[i];  // vertex id
        const float4   o    = st.ray_o[slot];
        float4         d    = st.ray_d[slot];

        RayT           ray           = makeRay({o.x, o.y, o.z}, {d.x, d.y, d.z}, 0.f, d.w);
        const uint32_t depth_surface = (__float_as_uint(o.w) >> 8) & 0xffu;
        clipToMedium(sc, st, slot, __float_as_uint(o.w), ray);

        HitD           isec = {0.f, 0.f, 0.f, 0};
        const uint32_t prop = sceneIntersect(sc, ray, depth_surface, isec);

        d.w             = ray.tmax;  // probe.ray.max_t = isec.t (prop_tree.zig:77); unchanged on a miss
        st.ray_d[slot]  = d;
        st.hit[slot]    = make_float4(isec.u, isec.v, __uint_as_float(isec.primitive), __uint_as_float(prop));
    }
    if (0 == blockIdx.x && 0 == threadIdx.x) atomicAdd(&st.counters[5], count);  // statistics: closest-hit rays
}

// Scene.visibility for every shadow-ray record of the surviving paths, scene.zig:229-235 (no volume props)
__global__ void __launch_bounds__(kBlock) shadowKernel(SceneDevice sc, PathState st) {
    const uint32_t count  = st.counters[1];
    const uint32_t stride = st.shadow_stride;
    const uint64_t items  = uint64_t(count) * stride;
    uint32_t       traced = 0;
    for (uint64_t i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint32_t slot = st.queue_b[uint32_t(i / stride)];
        const uint32_t k    = uint32_t(i % stride);
        if (k >= st.sh_n[slot]) continue;
        const size_t rec = size_t(slot) * stride + k;

        const float4 o = st.sh_o[rec];
        const float4 p = st.sh_p[rec];

        uint32_t   unused;
        const RayT ray = loadTraceRay<true>(st, uint32_t(rec), unused);

        st.sh_wi[rec].w = sceneVisibility(sc, ray) ? 1.f : 0.f;
        traced += 1;
    }
    for (int o = 16; o > 0; o >>= 1) traced += __shfl_down_sync(0xffffffffu, traced, o);
    if (0 == (threadIdx.x & 31u) && 0 != traced) atomicAdd(&st.counters[6], traced);  // statistics: shadow rays
}



struct SceneTraceConfig {
    int              variant;  // 0: one thread per ray (extendKernel / shadowKernel), 1: top kernel + persistent mesh kernel,
                               // 2: fused two-level persistent kernel for scenes with meshes (default)
    SceneTraceTuning tune;
    SceneStepTuning  step;
    int              blocks_per_sm;
    int              min_blocks;
    int              mesh_min_blocks;
    int              pool, pool_fetch_free;
    int              sort, sort_from, sort_mode;
};

const SceneTraceConfig& sceneTraceConfig() {
    static const SceneTraceConfig cfg = [] {
        SceneTraceConfig c;
        c.variant         = envInt("ZYGPU_SCENE_TRACE", 2);
        c.step.fetch_idle = uint32_t(envInt("ZYGPU_FUSED_FETCH_IDLE", 10));
        c.step.weight[0]  = uint32_t(envInt("ZYGPU_W_NODE", 1));
        c.step.weight[1]  = uint32_t(envInt("ZYGPU_W_TRI", 2));
        c.step.weight[2]  = uint32_t(envInt("ZYGPU_W_PROP", 2));
        c.step.weight[3]  = uint32_t(envInt("ZYGPU_W_ENTER", 2));
        c.tune.fetch_idle = uint32_t(envInt("ZYGPU_SCENE_FETCH_IDLE", 10));  // measured: 10 beats 6 by 1 - 2 % on the sphere and instanced scenes
        c.tune.tri_num    = uint32_t(envInt("ZYGPU_TRI_NUM", 1));
        c.tune.tri_den    = uint32_t(envInt("ZYGPU_TRI_DEN", 2));
        c.blocks_per_sm   = envInt("ZYGPU_SCENE_BLOCKS_PER_SM", 0);
        c.min_blocks      = envInt("ZYGPU_SCENE_MIN_BLOCKS", 0);
        c.mesh_min_blocks = envInt("ZYGPU_MESH_MIN_BLOCKS", 0);
        c.pool            = envInt("ZYGPU_SCENE_POOL", -1);  // the ray-pool variant of the fused kernel: 1 on, 0 off, -1 by prop-tree size
        c.pool_fetch_free = envInt("ZYGPU_POOL_FETCH_FREE", 16);
        c.step.repeat       = uint32_t(std::max(1, envInt("ZYGPU_POOL_REPEAT", 4)));
        c.step.repeat_lanes = uint32_t(envInt("ZYGPU_POOL_REPEAT_LANES", 16));
        c.sort            = envInt("ZYGPU_RAY_SORT", 0);       // bit 0: closest-hit rays, bit 1: shadow rays
        c.sort_from       = envInt("ZYGPU_RAY_SORT_FROM", 1);  // first bounce that sorts
        c.sort_mode       = envInt("ZYGPU_RAY_SORT_MODE", 1);  // 1: (cell, octant), 2: (octant, cell)
        c.step.sorted     = 0;
        c.step.debug_item = uint32_t(envInt("ZYGPU_DEBUG_TRACE_ITEM", -1));
        c.step.prefetch   = uint32_t(envInt("ZYGPU_SCENE_PREFETCH", 0));
        return c;
    }();
    return cfg;
}

template <bool AnyHit>
cudaError_t launchSceneTrace(const SceneDevice& scene, const PathState& st, uint32_t max_items, bool has_meshes, uint32_t bounce, cudaStream_t stream) {
    const SceneTraceConfig& cfg = sceneTraceConfig();
    // a prop tree that is a single leaf (a mesh and a few analytic props) gains nothing from the fused walk: the thread-per-ray
    // top kernel deals with it at full lane occupancy (measured on the 1M-triangle sphere scene: 48.2 ms against 51.4 ms fused)
    // (an instrumented pass always takes the fused kernel, the one that counts its fetches; the results are the same)
    // (a scene whose trees are too deep for the fused kernels' stacks takes the two-kernel path, whose levels have a stack each)
    const bool pool_fits  = scene.trace_stack_bound <= kScenePoolStack && nullptr != st.trace_stacks && nullptr == st.tally && 0 != cfg.pool;
    const bool fused_fits = scene.trace_stack_bound <= kWideStack || pool_fits;
    if ((2 == cfg.variant && has_meshes && scene.num_solid_nodes > 1 && fused_fits) || (nullptr != st.tally && scene.trace_stack_bound <= kWideStack)) {
        static int resident = 0, resident_counted = 0;
        // resident blocks per SM the kernel is compiled for (ZYGPU_SCENE_MIN_BLOCKS): 5 -> 96 registers, 6 -> 80, 7 -> 72, 8 -> 64.
        // Measured on config 3 (1920 x 1080 x 4 spp): 188.5 / 179.6 / 169.9 / 168.9 ms - the walk waits on memory (5 M triangles do not
        // fit the L2), more resident warps cover more of it; config 4 does not care (profiles/r02_sweeps.md)
        const int  mb = 0 != cfg.min_blocks ? cfg.min_blocks : 8;
        const auto fn = 8 == mb   ? sceneTracePersistent<AnyHit, false, 8>
                        : 7 == mb ? sceneTracePersistent<AnyHit, false, 7>
                        : 6 == mb ? sceneTracePersistent<AnyHit, false, 6>
                                  : sceneTracePersistent<AnyHit, false, 5>;
        if (0 == resident) {
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 128, 0);
            if (cfg.blocks_per_sm > 0) per_sm = std::min(per_sm, cfg.blocks_per_sm);
            resident = std::max(per_sm, 1) * numSms();
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sceneTracePersistent<AnyHit, true, 5>, 128, 0);
            resident_counted = std::max(per_sm, 1) * numSms();
        }
        const bool     counted = nullptr != st.tally;
        const uint32_t needed  = (max_items + 127) / 128;
        const uint32_t grid    = std::max(1u, std::min<uint32_t>(uint32_t(counted ? resident_counted : resident), needed));
        cudaError_t    err     = cudaMemsetAsync(st.counters + 8, 0, sizeof(uint32_t), stream);
        if (cudaSuccess != err) return err;
        SceneStepTuning step = cfg.step;
        // camera rays and their shadow rays arrive in pixel order: coherent as they are
        step.sorted = (nullptr != st.sort_bins && bounce >= uint32_t(cfg.sort_from) && 0 != (cfg.sort & (AnyHit ? 2 : 1))) ? 1u : 0u;
        if (0 != step.sorted) {
            err = cudaMemsetAsync(st.sort_bins, 0, (kSortBins + 1) * sizeof(uint32_t), stream);
            if (cudaSuccess != err) return err;
            const uint32_t sort_grid = std::max(1u, std::min<uint32_t>((max_items + 255) / 256, uint32_t(numSms()) * 8));
            sortKeyKernel<AnyHit><<<sort_grid, 256, 0, stream>>>(scene, st, uint32_t(cfg.sort_mode));
            sortScanKernel<<<1, 1024, 0, stream>>>(st);
            sortScatterKernel<AnyHit><<<sort_grid, 256, 0, stream>>>(st);
        }
        // Measured (profiles/r02_sweeps.md): 20.6 instead of 14.0 active lanes per instruction on config 3, but 30 % more thread-level work for
        // picking and moving the rays - 339.6 -> 322.3 ms per 8-spp frame there, 190.8 -> 200.2 ms on config 4, whose rays see few props. The
        // variant is taken for large prop trees (ZYGPU_SCENE_POOL = 0 / 1 overrides).
        const bool pooled = pool_fits && (scene.trace_stack_bound > kWideStack || (-1 == cfg.pool ? scene.num_solid_nodes >= 4096u : true));
        if (!counted && pooled && nullptr != st.trace_stacks && scene.num_solid_nodes > 1) {
            // the ray-pool variant: rays ready for the same kind of step are handed to the lanes first
            static int pool_resident = 0;
            if (0 == pool_resident) {
                int per_sm = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scenePoolTrace<AnyHit>, 128, 0);
                pool_resident = std::max(per_sm, 1) * numSms();
            }
            const uint32_t pool_grid = std::max(1u, std::min<uint32_t>(std::min<uint32_t>(uint32_t(pool_resident), st.trace_stack_blocks), needed));
            step.fetch_idle          = uint32_t(cfg.pool_fetch_free);
            scenePoolTrace<AnyHit><<<pool_grid, 128, 0, stream>>>(scene, st, st.counters + 8, step, st.trace_stacks);
        } else if (counted) {
            sceneTracePersistent<AnyHit, true, 5><<<grid, 128, 0, stream>>>(scene, st, st.counters + 8, step, st.tally);
        } else {
            fn<<<grid, 128, 0, stream>>>(scene, st, st.counters + 8, step, nullptr);
        }
        return cudaGetLastError();
    }
    // counters[2] = mesh queue length, counters[8] = work counter of the persistent kernel
    cudaError_t err = cudaMemsetAsync(st.counters + 2, 0, sizeof(uint32_t), stream);
    if (cudaSuccess != err) return err;
    topKernel<AnyHit><<<gridFor(max_items, walkGrid()), kBlock, 0, stream>>>(scene, st);
    err = cudaGetLastError();
    if (cudaSuccess != err || !has_meshes) return err;

    // resident blocks per SM the mesh kernel is compiled for (ZYGPU_MESH_MIN_BLOCKS): 6 -> 80 registers, 7 -> 72, 8 -> 64
    const int  mmb = 0 != cfg.mesh_min_blocks ? cfg.mesh_min_blocks : 8;  // sphere scene: 48.9 (6 / 7) -> 47.1 ms (8)
    const auto mfn = 8 == mmb ? meshTracePersistent<AnyHit, 8> : (7 == mmb ? meshTracePersistent<AnyHit, 7> : meshTracePersistent<AnyHit, 6>);
    static int resident = 0;
    if (0 == resident) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mfn, 128, 0);
        if (cfg.blocks_per_sm > 0) per_sm = std::min(per_sm, cfg.blocks_per_sm);
        resident = std::max(per_sm, 1) * numSms();
    }
    const uint32_t needed = (max_items + 127) / 128;
    const uint32_t grid   = std::max(1u, std::min<uint32_t>(uint32_t(resident), needed));
    err                   = cudaMemsetAsync(st.counters + 8, 0, sizeof(uint32_t), stream);
    if (cudaSuccess != err) return err;
    mfn<<<grid, 128, 0, stream>>>(scene, st, st.counters + 8, cfg.tune);
    return cudaGetLastError();
}

}  // namespace

int numSmsOfCurrentDevice() { return numSms(); }

uint32_t sceneTraceLaunches(bool has_meshes, uint32_t num_solid_nodes) {  // kernels per extend / shadow stage
    const int v = sceneTraceConfig().variant;
    if (0 == v || !has_meshes) return 1u;
    return (2 == v && num_solid_nodes > 1) ? 1u : 2u;  // fused kernel, or top kernel + mesh kernel
}

cudaError_t launchExtend(const SceneDevice& scene, const PathState& st, uint32_t max_items, bool has_meshes, uint32_t bounce, cudaStream_t stream) {
    if (0 != sceneTraceConfig().variant) return launchSceneTrace<false>(scene, st, max_items, has_meshes, bounce, stream);
    extendKernel<<<gridFor(max_items, walkGrid()), kBlock, 0, stream>>>(scene, st);
    return cudaGetLastError();
}
cudaError_t launchExtendReference(const SceneDevice& scene, const PathState& st, uint32_t max_items, cudaStream_t stream) {
    extendKernel<<<gridFor(max_items, walkGrid()), kBlock, 0, stream>>>(scene, st);
    return cudaGetLastError();
}

namespace {
__global__ void compareHitsKernel(PathState st, const float4* __restrict__ ray_d_before, const float4* __restrict__ ray_d_a,
                                  const float4* __restrict__ hit_a, uint32_t bounce) {
    const uint32_t count = st.counters[st.lanes > 1 ? 7 : 0];
    const uint32_t* __restrict__ queue = st.lanes > 1 ? st.queue_t : st.queue_a;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t item = queue[i];
        const float4   ha = hit_a[item], hb = st.hit[item];
        const float    ta = ray_d_a[item].w, tb = st.ray_d[item].w;
        if (__float_as_uint(ta) != __float_as_uint(tb) || __float_as_uint(ha.z) != __float_as_uint(hb.z) || __float_as_uint(ha.w) != __float_as_uint(hb.w)) {
            const float4 o = st.ray_o[item], d = ray_d_before[item];
            printf("[verify] bounce %u item %u o %.9g %.9g %.9g d %.9g %.9g %.9g max_t %.9g | fused t %.9g prim %u prop %u | reference t %.9g prim %u prop %u\n",
                   bounce, item, o.x, o.y, o.z, d.x, d.y, d.z, d.w, ta, __float_as_uint(ha.z), __float_as_uint(ha.w), tb, __float_as_uint(hb.z),
                   __float_as_uint(hb.w));
        }
    }
}
}  // namespace

cudaError_t launchCompareHits(const PathState& st, const float4* ray_d_before, const float4* ray_d_a, const float4* hit_a, uint32_t max_items,
                              uint32_t bounce, cudaStream_t stream) {
    compareHitsKernel<<<gridFor(max_items, 16), kBlock, 0, stream>>>(st, ray_d_before, ray_d_a, hit_a, bounce);
    return cudaGetLastError();
}

cudaError_t launchShadow(const SceneDevice& scene, const PathState& st, uint32_t max_items, bool has_meshes, uint32_t bounce, cudaStream_t stream) {
    if (0 != sceneTraceConfig().variant) return launchSceneTrace<true>(scene, st, max_items * st.shadow_stride, has_meshes, bounce, stream);
    shadowKernel<<<gridFor(max_items, walkGrid()), kBlock, 0, stream>>>(scene, st);
    return cudaGetLastError();
}

}  // namespace zygpu
