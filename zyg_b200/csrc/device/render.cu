#include "shading.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace zygpu {

namespace {

constexpr uint32_t kBlock = 128;
#ifndef ZYGPU_SHADE_BLOCKS
#define ZYGPU_SHADE_BLOCKS 4  // resident blocks per SM the shade kernels are compiled for (128 registers)
#endif

// ---- state packing ---------------------------------------------------------------------------

enum : uint32_t {  // Vertex.State, vertex.zig:19-28
    kPrimaryRay      = 1u << 0,
    kTransparent     = 1u << 1,
    kSingular        = 1u << 2,
    kSpecular        = 1u << 3,
    kTranslucent     = 1u << 4,
    kStartedSpecular = 1u << 5,
};

__device__ __forceinline__ uint32_t packFlags(uint32_t state, uint32_t probe_depth, uint32_t vertex_depth, uint32_t path_count_log2 = 0,
                                              uint32_t num_media = 0) {
    return state | (probe_depth << 8) | (vertex_depth << 16) | (path_count_log2 << 24) | (num_media << 26);
}

// ---- vertex pool of one camera sample (Pool, vertex.zig:215-310) -------------------------------
//
// One word per slot: bits 0-7 lanes of the current generation in processing order (2 bits each), 8-10 their number,
// 11-18 / 19-21 the same for the next generation, 22-25 lanes in use. path_count bounds the live vertices by 4.

__device__ __forceinline__ uint32_t poolCurCount(uint32_t m) { return (m >> 8) & 7u; }
__device__ __forceinline__ uint32_t poolCurLane(uint32_t m, uint32_t k) { return (m >> (2 * k)) & 3u; }
__device__ __forceinline__ uint32_t poolNextCount(uint32_t m) { return (m >> 19) & 7u; }
__device__ __forceinline__ uint32_t poolSwap(uint32_t m) { return (m & 0x03C00000u) | ((m >> 11) & 0x7FFu); }
__device__ __forceinline__ uint32_t poolFree(uint32_t m, uint32_t lane) { return m & ~(1u << (22 + lane)); }
__device__ __forceinline__ uint32_t poolAlloc(uint32_t m) { return uint32_t(__ffs(int(~(m >> 22) & 0xFu))) - 1u; }  // 0xFFFFFFFF when full
__device__ __forceinline__ uint32_t poolAppendNext(uint32_t m, uint32_t lane) {
    const uint32_t n = poolNextCount(m);
    return (m | (lane << (11 + 2 * n)) | (1u << (22 + lane))) + (1u << 19);
}
constexpr uint32_t kPoolFirst = (1u << 19) | (1u << 22);  // after generate: lane 0 is the next generation

// ---- medium stack (Stack, prop/medium.zig:30-153) ----------------------------------------------

struct MediaD {
    uint32_t count;
    uint32_t prop[3];  // Num_entries - 1 entries can be pushed (:117-131)
    uint32_t part[3];
};

__device__ __forceinline__ MediaD unpackMedia(uint4 w, uint32_t count) {
    return {count, {w.x, w.y, w.z}, {w.w & 0xffu, (w.w >> 8) & 0xffu, (w.w >> 16) & 0xffu}};
}
__device__ __forceinline__ uint4 packMedia(const MediaD& m) {
    return make_uint4(m.prop[0], m.prop[1], m.prop[2], m.part[0] | (m.part[1] << 8) | (m.part[2] << 16));
}
__device__ __forceinline__ void mediaPush(MediaD& m, uint32_t prop, uint32_t part) {
    if (m.count < 3) {
        m.prop[m.count] = prop;
        m.part[m.count] = part;
        m.count += 1;
    }
}
__device__ __forceinline__ void mediaRemove(MediaD& m, uint32_t prop, uint32_t part) {
    for (int i = int(m.count) - 1; i >= 0; --i) {
        if (m.prop[i] == prop && m.part[i] == part) {
            for (int j = i; j < int(m.count) - 1; ++j) {
                m.prop[j] = m.prop[j + 1];
                m.part[j] = m.part[j + 1];
            }
            m.count -= 1;
            return;
        }
    }
}

// ray_offset.zig:29-31
__device__ __forceinline__ float offsetF(float t) {
    return t < (1.f / 32.f) ? t + (1.f / 65536.f) : __int_as_float(int(uint32_t(__float_as_int(t)) + 256u));
}

constexpr float kLowThreshold = 0.00000001f;  // helper.zig:29

__device__ __forceinline__ float splitThreshold(float threshold, uint32_t total_depth) {  // helper.zig:33-39
    return zmin(total_depth < 4 ? threshold : kLowThreshold, threshold);
}
__device__ __forceinline__ float powerHeuristic(float f_pdf, float g_pdf) {  // helper.zig:64-67
    const float f2 = f_pdf * f_pdf;
    return __fdiv_rn(f2, __fmaf_rn(g_pdf, g_pdf, f2));
}
__device__ __forceinline__ float predividedPowerHeuristic(float f_pdf, float g_pdf) {  // helper.zig:70-73
    const float f2 = f_pdf * f_pdf;
    return __fdiv_rn(f_pdf, __fmaf_rn(g_pdf, g_pdf, f2));
}

// ---- per-slot sampler state ------------------------------------------------------------------

struct SlotId {
    uint32_t pixel_id;   // over the padded resolution, worker.zig:127-141
    uint32_t iteration;  // absolute sample number
};

__device__ __forceinline__ SlotId slotId(uint32_t slot, const PassParams& pass) {
    const uint32_t padded = pass.padded_w * pass.padded_h;
    const uint32_t s      = slot / padded;
    return {slot - s * padded, pass.iteration + s};
}

// worker.zig:143-149 with num_samples = 1 per iteration (Driver.renderIterations(iteration, 1))
__device__ __forceinline__ void seedSamplers(const SlotId id, const PassParams& pass, uint32_t spp_total, SobolD& sobol, PcgD& rng) {
    const uint32_t a = pass.padded_w * pass.padded_h;
    const uint64_t o = uint64_t(id.iteration) * a;
    rng.start(0, uint64_t(id.pixel_id) + o);

    const uint64_t sample_index = uint64_t(id.pixel_id) * uint64_t(spp_total) + uint64_t(id.iteration);
    const uint32_t tsi          = uint32_t(sample_index);
    const uint32_t seed         = uint32_t(sample_index >> 32) + id.iteration / spp_total;
    sobol.startPixel(tsi, seed);
}

__device__ __forceinline__ void loadSampler(const PathState& st, uint32_t slot, uint4 s, const PassParams& pass, uint32_t spp_total,
                                            uint32_t total_depth, SamplerD& sampler) {
    const SlotId id = slotId(slot, pass);
    sampler.use_sobol = total_depth < 3;  // pickSampler; a Random take sampler is handled by the caller (view.sampler)
    const uint64_t sample_index = uint64_t(id.pixel_id) * uint64_t(spp_total) + uint64_t(id.iteration);
    if (sampler.use_sobol) {
        sampler.sobol.restore(uint32_t(sample_index), s.x, s.y, s.z);
    } else {
        sampler.sobol.sample     = uint32_t(sample_index);
        sampler.sobol.block_seed = s.x;
        sampler.sobol.run_seed   = s.y;
        sampler.sobol.dimension  = s.z;
    }
    const uint2 r     = st.rng[slot];
    sampler.rng.state = (uint64_t(r.y) << 32) | r.x;
    const uint64_t a  = uint64_t(pass.padded_w) * pass.padded_h;
    sampler.rng.inc   = ((uint64_t(id.pixel_id) + uint64_t(id.iteration) * a) << 1) | 1;
}

__device__ __forceinline__ void storeSampler(const PathState& st, uint32_t slot, const SamplerD& sampler, uint32_t aux) {
    st.smp[slot] = make_uint4(sampler.sobol.block_seed, sampler.sobol.run_seed, sampler.sobol.dimension, aux);
    st.rng[slot] = make_uint2(uint32_t(sampler.rng.state), uint32_t(sampler.rng.state >> 32));
}

// ---- queues ----------------------------------------------------------------------------------

// Warp-aggregated append: one atomic per warp.
__device__ __forceinline__ void queuePush(uint32_t* queue, uint32_t* counter, bool push, uint32_t value) {
    const uint32_t mask = __ballot_sync(0xffffffffu, push);
    if (0 == mask) return;
    const uint32_t lane   = threadIdx.x & 31u;
    const uint32_t leader = __ffs(mask) - 1;
    uint32_t       base   = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (push) queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

// ---- scene queries ---------------------------------------------------------------------------

__device__ __forceinline__ bool propVisible(uint32_t flags, uint32_t depth_surface) {  // prop.zig:38-48
    return 0 == depth_surface ? 0 != (flags & ZYG_PROP_VISIBLE_IN_CAMERA) : 0 != (flags & ZYG_PROP_VISIBLE_IN_REFLECTION);
}

__device__ __forceinline__ bool aabbIntersect(const float4* aabbs, uint32_t i, const RayT& ray) {  // aabb.zig:46-60
    return FLT_MAX != intersectNode(__ldg(aabbs + 2 * size_t(i)), __ldg(aabbs + 2 * size_t(i) + 1), ray);
}

// AABB.intersectP, aabb.zig:62-84
__device__ __forceinline__ float aabbIntersectP(float4 mi, float4 ma, const RayT& ray) {
    const float lx = (mi.x - ray.o.x) * ray.inv_d.x, ly = (mi.y - ray.o.y) * ray.inv_d.y, lz = (mi.z - ray.o.z) * ray.inv_d.z;
    const float ux = (ma.x - ray.o.x) * ray.inv_d.x, uy = (ma.y - ray.o.y) * ray.inv_d.y, uz = (ma.z - ray.o.z) * ray.inv_d.z;

    const float imin = zmax(zmax(zmin(lx, ux), zmin(ly, uy)), zmin(lz, uz));
    const float imax = zmin(zmin(zmax(lx, ux), zmax(ly, uy)), zmax(lz, uz));

    const float tboxmin = zmax(imin, ray.tmin);
    const float tboxmax = zmin(imax, ray.tmax);

    if (tboxmin <= tboxmax) return imin < ray.tmin ? imax : imin;
    return FLT_MAX;
}

// VolumeIntegrator.integrate, volume_integrator.zig:97-99: a vertex inside a medium only looks as far as the medium prop's box
__device__ __forceinline__ void clipToMedium(const SceneDevice& sc, const PathState& st, uint32_t vertex_id, uint32_t flags, RayT& ray) {
    const uint32_t num_media = (flags >> 26) & 3u;
    if (nullptr == st.med || 0 == num_media) return;
    const uint4    w    = st.med[vertex_id];
    const uint32_t prop = 1 == num_media ? w.x : (2 == num_media ? w.y : w.z);
    const float    limit = aabbIntersectP(__ldg(sc.aabbs + 2 * size_t(prop)), __ldg(sc.aabbs + 2 * size_t(prop) + 1), ray);
    ray.tmax             = zmin(offsetF(limit), ray.tmax);
}

__device__ __forceinline__ float shapeArea(uint32_t shape, V3 scale) {  // shape.zig:143-156
    switch (shape) {
        case ZYG_SHAPE_RECTANGLE: return scale.x * scale.y;
        case ZYG_SHAPE_SPHERE: return (4.f * kPi) * ((0.5f * scale.x) * (0.5f * scale.x));
        case ZYG_SHAPE_DISTANT: return distantSolidAngle(scale.x);
        case ZYG_SHAPE_CANOPY: return 2.f * kPi;
        default: return 0.f;
    }
}

// Prop.intersect + Shape.intersect, prop.zig:163-197, shape.zig:165-179
__device__ __forceinline__ bool propIntersect(const SceneDevice& sc, uint32_t entity, RayT& ray, uint32_t depth_surface, HitD& isec) {
    const ZygpuProp prop = sc.props[entity];
    if (!propVisible(prop.flags, depth_surface)) return false;
    if (!aabbIntersect(sc.aabbs, entity, ray)) return false;
    const TrafoD trafo = loadTrafo(sc.trafos, entity);
    switch (prop.shape) {
        case ZYG_SHAPE_CUBE: return cubeIntersect(ray, trafo, isec);
        case ZYG_SHAPE_RECTANGLE: return rectangleIntersect(ray, trafo, isec);
        case ZYG_SHAPE_SPHERE: return sphereIntersect(ray, trafo, isec);
        case ZYG_SHAPE_TRIANGLE_MESH: {
            // TriangleTree.intersect, triangle_tree.zig:46-109: the ray goes to object space un-normalised, so t is shared
            WideRay w;
            w.ray = worldToObjectRay(trafo, ray);
            setupWideRay(w);
            float    ht, hu, hv;
            uint32_t prim;
            if (traverseWide<false>(sc.meshes[prop.mesh], w, ht, hu, hv, prim)) {
                isec = {ht, hu, hv, prim};
                return true;
            }
            return false;
        }
        default: return false;
    }
}

// Prop.visibility, prop.zig:199-237 (no masks): true = unoccluded
__device__ __forceinline__ bool propVisibility(const SceneDevice& sc, uint32_t entity, const RayT& ray) {
    const ZygpuProp prop = sc.props[entity];
    if (0 == (prop.flags & ZYG_PROP_VISIBLE_IN_SHADOW)) return true;
    if (!aabbIntersect(sc.aabbs, entity, ray)) return true;
    const TrafoD trafo = loadTrafo(sc.trafos, entity);
    switch (prop.shape) {
        case ZYG_SHAPE_CUBE: return !cubeIntersectP(ray, trafo);
        case ZYG_SHAPE_RECTANGLE: {
            HitD unused;
            return !rectangleIntersect(ray, trafo, unused);
        }
        case ZYG_SHAPE_SPHERE: {
            HitD unused;
            return !sphereIntersect(ray, trafo, unused);
        }
        case ZYG_SHAPE_TRIANGLE_MESH: {
            WideRay w;
            w.ray = worldToObjectRay(trafo, ray);
            setupWideRay(w);
            float    ht, hu, hv;
            uint32_t prim;
            return !traverseWide<true>(sc.meshes[prop.mesh], w, ht, hu, hv, prim);
        }
        default: return true;
    }
}

constexpr uint32_t kPropStack = 64;  // prop trees are shallow; the reference's NodeStack holds 127

// PropBvh.intersect, prop_tree.zig:56-116: reference order, so equal-t ties resolve like the reference.
__device__ __forceinline__ uint32_t sceneIntersect(const SceneDevice& sc, RayT& ray, uint32_t depth_surface, HitD& isec) {
    uint32_t stack[kPropStack];
    uint32_t end = 0;
    uint32_t n   = 0 == sc.num_solid_nodes ? kEnd : 0;

    uint32_t prop = kEnd;

    while (kEnd != n) {
        const float4 nmin = __ldg(sc.solid_nodes + 2 * size_t(n));
        const float4 nmax = __ldg(sc.solid_nodes + 2 * size_t(n) + 1);

        const uint32_t num = __float_as_uint(nmax.w);
        if (0 != num) {
            const uint32_t start = __float_as_uint(nmin.w);
            for (uint32_t i = start; i < start + num; ++i) {
                const uint32_t p = __ldg(sc.solid_indices + i);
                HitD           h;
                if (propIntersect(sc, p, ray, depth_surface, h)) {
                    ray.tmax = h.t;
                    isec     = h;
                    prop     = p;
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
    return prop;
}

// PropBvh.visibility, prop_tree.zig:185-240
__device__ __forceinline__ bool sceneVisibility(const SceneDevice& sc, const RayT& ray) {
    uint32_t stack[kPropStack];
    uint32_t end = 0;
    uint32_t n   = 0 == sc.num_solid_nodes ? kEnd : 0;

    while (kEnd != n) {
        const float4 nmin = __ldg(sc.solid_nodes + 2 * size_t(n));
        const float4 nmax = __ldg(sc.solid_nodes + 2 * size_t(n) + 1);

        const uint32_t num = __float_as_uint(nmax.w);
        if (0 != num) {
            const uint32_t start = __float_as_uint(nmin.w);
            for (uint32_t i = start; i < start + num; ++i) {
                if (!propVisibility(sc, __ldg(sc.solid_indices + i), ray)) return false;
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
    return true;
}

// ---- two-level traversal, product path ---------------------------------------------------------
//
// The extend and shadow stages run as two kernels each:
//
//   top    one thread per ray walks the prop tree in the reference's order (binary nodes, near child first, leaf props in
//          order: prop_tree.zig:56-116, 185-240). Analytic props are tested where they are met; a triangle-mesh prop whose
//          world box the ray hits is appended to the ray's candidate list instead of being entered. Rays with candidates
//          go to the mesh queue. All threads do the same short walk, so the warps stay full.
//   mesh   persistent kernel over the mesh queue: a lane takes a ray, moves it into the object space of its next candidate
//          (re-testing the world box against the shrunken max_t first) and traverses the 8-wide BVH. Warps run the
//          lock-step loop of trace.cu's persistent kernel — NODE steps and TRIANGLE steps over the lanes that have that
//          kind of work, postponing triangle groups — and lanes whose ray ran out of candidates are refilled from the
//          queue (one global atomic per 1024 items), so incoherent bounces keep their lanes busy.
//
// Relative to the reference only the order in which props are tested changes (all analytic props of the walk first, then
// the meshes in walk order): the closest hit is the same except for equal-t ties between different props.

// Equal-t ties. The reference accepts `hit_t <= max_t`, so of two hits at the same t the one tested later wins (triangle.zig:47,
// prop_tree.zig:76-79) — later in ITS traversal order. The device visits nodes in another order, and in the lock-step kernels
// the order even depends on the warp's votes; resolving ties by (prop id, primitive id), larger wins, makes the result
// independent of the schedule (renders are bit-reproducible) and agrees with the reference inside a leaf, where later = larger.
__device__ __forceinline__ bool closerOrLater(float t, float tmax, uint32_t prop, uint32_t prim, uint32_t hit_prop, uint32_t hit_prim) {
    return kEnd == hit_prop || t < tmax || prop > hit_prop || (prop == hit_prop && prim > hit_prim);
}

constexpr uint32_t kMeshCandidates = 8;  // per ray; further meshes are traversed inline by the top kernel
constexpr uint32_t kScenePoolItems = 1024;

struct SceneTraceTuning {
    uint32_t fetch_idle;  // refill when at least this many lanes are idle
    uint32_t tri_num, tri_den;
};

template <bool AnyHit>
__device__ __forceinline__ RayT loadTraceRay(const PathState& st, uint32_t item, uint32_t& depth_surface, uint32_t* flags_out = nullptr) {
    if (AnyHit) {  // Shape.shadowRay, shape.zig:401-416: the record holds both end points
        const float4 o           = st.sh_o[item];
        const float4 p           = st.sh_p[item];
        const V3     origin      = {o.x, o.y, o.z};
        depth_surface            = 0;
        if (0 != (__float_as_uint(p.w) & 0x80000000u)) {  // Shape.shadowRay for Canopy / Distant / Dome
            const float4 wi = st.sh_wi[item];
            return makeRay(origin, {wi.x, wi.y, wi.z}, 0.f, kRayMaxT);
        }
        const V3     shadow_axis = sub3({p.x, p.y, p.z}, origin);
        const float  shadow_len  = length3(shadow_axis);
        depth_surface            = 0;
        return makeRay(origin, divs3(shadow_axis, shadow_len), 0.f, shadow_len);
    }
    const float4 o = st.ray_o[item];
    const float4 d = st.ray_d[item];
    depth_surface  = (__float_as_uint(o.w) >> 8) & 0xffu;
    if (flags_out) *flags_out = __float_as_uint(o.w);
    return makeRay({o.x, o.y, o.z}, {d.x, d.y, d.z}, 0.f, d.w);
}

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

template <bool AnyHit>
__global__ void __launch_bounds__(128) meshTracePersistent(SceneDevice sc, PathState st, uint32_t* __restrict__ work_counter,
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

                const uint32_t hitmask = testWideNode(w, w.ray.tmin, w.ray.tmax, n0, n1, n2, n3, n4);

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
    return tc - sphere.w * rsqrtf(dd) * 1.0001f <= ray.tmax;
}

struct SceneStepTuning {
    uint32_t fetch_idle;  // refill when at least this many lanes are idle
    uint32_t weight[4];   // NODE, TRIANGLE, PROP, ENTER: the step kind with the largest (ready lanes x weight) runs
};

template <bool AnyHit, bool Count>
__global__ void __launch_bounds__(128) sceneTracePersistent(SceneDevice sc, PathState st, uint32_t* __restrict__ work_counter,
                                                            SceneStepTuning tune, unsigned long long* __restrict__ tally) {
    constexpr uint32_t kFull = 0xffffffffu;
    const uint32_t     lane  = threadIdx.x & 31u;

    // the trace queue: closest-hit rays are vertex ids, shadow rays are records (compact list, or `stride` slots per path)
    const uint32_t stride = st.shadow_stride;
    const uint32_t* __restrict__ closest_queue = st.lanes > 1 ? st.queue_t : st.queue_a;
    const bool     compact = AnyHit && nullptr != st.queue_r;
    const uint64_t total   = AnyHit ? (compact ? uint64_t(st.counters[10]) : uint64_t(st.counters[1]) * stride)
                                    : uint64_t(st.counters[st.lanes > 1 ? 7 : 0]);
    const uint32_t n       = uint32_t(total < 0xFFFFFFFFull ? total : 0xFFFFFFFFull);

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

                    const uint32_t hitmask = testWideNode(w, w.ray.tmin, w.ray.tmax, n0, n1, n2, n3, n4);

                    node_group.x = __float_as_uint(n1.x);
                    node_group.y = (hitmask & 0xFF000000u) | (__float_as_uint(n0.w) >> 24);
                    tri_group.x  = __float_as_uint(n1.y);
                    tri_group.y  = hitmask & 0x00FFFFFFu;
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

// Shape.fragment, shape.zig:205-219
__device__ __forceinline__ void shapeFragment(const SceneDevice& sc, uint32_t prop, const RayT& ray, const HitD& isec, FragD& frag) {
    frag.prop      = prop;
    frag.primitive = isec.primitive;
    frag.trafo     = loadTrafo(sc.trafos, prop);
    switch (sc.props[prop].shape) {
        case ZYG_SHAPE_CUBE: cubeFragment(ray, isec, frag); break;
        case ZYG_SHAPE_RECTANGLE: rectangleFragment(ray, isec, frag); break;
        case ZYG_SHAPE_SPHERE: sphereFragment(ray, isec, frag); break;
        // Distant and Canopy are infinite props: only met on escape, where shade_a calls their fragment functions itself
        case ZYG_SHAPE_TRIANGLE_MESH: meshFragment(sc.mesh_shading[sc.props[prop].mesh], isec, frag); break;
        default: break;
    }
}

// ---- lights ----------------------------------------------------------------------------------

struct VertexD {  // the parts of Vertex (vertex.zig:45-64) the light code reads
    RayT     ray;
    V3       origin, geo_n;
    float    bxdf_pdf, light_split_threshold;
    uint32_t state, probe_depth;
};

__device__ __forceinline__ uint32_t lightNumSamples(const ZygpuLight& l, float split_threshold) {  // shape_sampler.zig:35-41
    return split_threshold <= kLowThreshold ? 1u : l.num_samples;
}

// C. Ureña, M. Fajardo, A. King: An Area-Preserving Parametrization for Spherical Rectangles. rectangle.zig:199-303
struct SphQuadD {
    V3    o, x, y, z;
    float z0, x0, y0, x1, y1, b0, b1, k, S;

    __device__ void init(V3 scale, V3 origin) {
        const V3 s  = {-0.5f * scale.x, -0.5f * scale.y, 0.f};
        const V3 ex = {scale.x, 0.f, 0.f};
        const V3 ey = {0.f, scale.y, 0.f};

        o               = origin;
        const float exl = length3(ex);
        const float eyl = length3(ey);
        x               = divs3(ex, exl);
        y               = divs3(ey, eyl);
        z               = cross3(x, y);
        const V3 d      = sub3(s, o);
        z0              = dot3(d, z);
        if (z0 > 0.f) {
            z  = neg3(z);
            z0 = -z0;
        }
        x0 = dot3(d, x);
        y0 = dot3(d, y);
        x1 = x0 + exl;
        y1 = y0 + eyl;

        const V3 v00 = {x0, y0, z0}, v01 = {x0, y1, z0}, v10 = {x1, y0, z0}, v11 = {x1, y1, z0};

        const V3 n0 = normalize3(cross3(v00, v10));
        const V3 n1 = normalize3(cross3(v10, v11));
        const V3 n2 = normalize3(cross3(v11, v01));
        const V3 n3 = normalize3(cross3(v01, v00));

        const float g0 = acosf(-dot3(n0, n1));
        const float g1 = acosf(-dot3(n1, n2));
        const float g2 = acosf(-dot3(n2, n3));
        const float g3 = acosf(-dot3(n3, n0));

        b0 = n0.z;
        b1 = n2.z;
        k  = 2.f * kPi - g2 - g3;
        S  = g0 + g1 - k;
    }

    __device__ V3 sample(float u0, float u1) const {
        const float au = u0 * S + k;
        float       sa, ca;
        sincosf(au, &sa, &ca);
        const float fu = __fdiv_rn(ca * b0 - b1, sa);
        float       cu = __fdiv_rn(1.f, __fsqrt_rn(fu * fu + b0 * b0)) * (fu > 0.f ? 1.f : -1.f);
        cu             = cu < -1.f ? -1.f : (cu > 1.f ? 1.f : cu);

        float xu = __fdiv_rn(-(cu * z0), __fsqrt_rn(1.f - cu * cu));
        xu       = xu < x0 ? x0 : (xu > x1 ? x1 : xu);

        const float d   = __fsqrt_rn(xu * xu + z0 * z0);
        const float h0  = __fdiv_rn(y0, __fsqrt_rn(d * d + y0 * y0));
        const float h1  = __fdiv_rn(y1, __fsqrt_rn(d * d + y1 * y1));
        const float hv  = h0 + u1 * (h1 - h0);
        const float hv2 = hv * hv;
        const float eps = __uint_as_float(0x35800000u);
        const float yv  = hv2 < 1.f - eps ? __fdiv_rn(hv * d, __fsqrt_rn(1.f - hv2)) : y1;

        return add3(add3(add3(o, scale3(xu, x)), scale3(yv, y)), scale3(z0, z));
    }

    __device__ float pdf(V3 scale) const {
        const float sqr_dist = squaredLength3(o);
        const float area     = scale.x * scale.y;
        const float numer    = area * fabsf(o.z);
        const float denom    = sqr_dist * __fsqrt_rn(sqr_dist);
        return numer > denom * kDotMin ? __fdiv_rn(1.f, S) : __fdiv_rn(denom, numer);
    }
};

struct LightPropsD {  // light.zig:25-30, scene.zig:664-674
    V3    center;
    float radius;
    V3    cone_axis;
    float cone_cos;
    float power;
    bool  two_sided;
};

__device__ __forceinline__ LightPropsD lightProperties(const SceneDevice& sc, uint32_t light_id) {
    const float4 mi   = __ldg(sc.light_aabbs + 2 * size_t(light_id));
    const float4 ma   = __ldg(sc.light_aabbs + 2 * size_t(light_id) + 1);
    const float4 cone = __ldg(sc.light_cones + light_id);
    return {{0.5f * (mi.x + ma.x), 0.5f * (mi.y + ma.y), 0.5f * (mi.z + ma.z)}, ma.w, {cone.x, cone.y, cone.z}, cone.w, mi.w,
            0 != sc.lights[light_id].two_sided};
}

__device__ __forceinline__ float clampedCosSub(float cos_a, float cos_b, float sin_a, float sin_b) {  // light_tree.zig:217-220
    const float angle = __fmaf_rn(cos_a, cos_b, sin_a * sin_b);
    return cos_a > cos_b ? 1.f : angle;
}
__device__ __forceinline__ float clampedSinSub(float cos_a, float cos_b, float sin_a, float sin_b) {  // :222-225
    const float angle = __fmaf_rn(sin_a, cos_b, -sin_b * cos_a);
    return cos_a > cos_b ? 0.f : angle;
}

// light_tree.zig:173-215. Out of line on purpose: every argument is a value (nothing forces kernel parameters into local memory), and the
// tree code calls it from a dozen places - inlined, its IEEE divisions and square roots alone were a fifth of the shade kernels' code,
// which are bound by instruction fetch (ncu: `no_instruction` next to `long_scoreboard`).
#ifndef ZYGPU_IMPORTANCE_INLINE
#define ZYGPU_IMPORTANCE_INLINE __noinline__
#endif
__device__ ZYGPU_IMPORTANCE_INLINE float lightImportance(V3 p, V3 n, V3 center, V3 cone_axis, float cos_cone, float radius, float power, bool two_sided,
                                 bool total_sphere) {
    const V3    axis = sub3(p, center);
    const float l    = length3(axis);
    const V3    na   = divs3(axis, l);

    const float sin_cu = zmin(__fdiv_rn(radius, l), 1.f);
    const float dca    = dot3(cone_axis, na);
    const float cos_a  = two_sided ? fabsf(dca) : dca;
    const float cos_n  = zmax(-dot3(n, na), 0.f);

    const float cos_cu   = __fsqrt_rn(zmax(__fmaf_rn(sin_cu, -sin_cu, 1.f), 0.f));
    const float sin_cone = __fsqrt_rn(zmax(__fmaf_rn(cos_cone, -cos_cone, 1.f), 0.f));
    const float sin_a    = __fsqrt_rn(zmax(__fmaf_rn(cos_a, -cos_a, 1.f), 0.f));
    const float sin_n    = __fsqrt_rn(zmax(__fmaf_rn(cos_n, -cos_n, 1.f), 0.f));

    const float ta = clampedCosSub(cos_a, cos_cone, sin_a, sin_cone);
    const float tb = clampedSinSub(cos_a, cos_cone, sin_a, sin_cone);
    const float tc = clampedCosSub(ta, cos_cu, tb, sin_cu);
    const float tn = clampedCosSub(cos_n, cos_cu, sin_n, sin_cu);

    const float ra = total_sphere ? 1.f : tn;
    const float rb = zmax(tc, 0.f);

    const float clamped_dist = zmax(l, 0.5f * radius);
    const float rc           = __fdiv_rn(power, clamped_dist * clamped_dist);

    return zmax(ra * rb * rc, 0.f);
}

// The tree a traversal runs over: the scene's (lights = scene lights) or the PrimitiveTree of a mesh sampler (lights = the
// emitting triangles of the part, light_tree.zig:520-719).
struct TreeD {
    const ZygpuLightNode*    nodes;
    const uint32_t*          middles;
    const uint32_t*          orders;
    const uint32_t*          mapping;
    float4                   bounds_min, bounds_max;
    const MeshSamplerDevice* sampler;  // null for the scene tree
};

__device__ __forceinline__ TreeD sceneTree(const SceneDevice& sc) {
    return {sc.lt_nodes, sc.lt_middles, sc.lt_orders, sc.lt_mapping, sc.lt_bounds_min, sc.lt_bounds_max, nullptr};
}
__device__ __forceinline__ TreeD primitiveTree(const MeshSamplerDevice& m) {
    return {m.nodes, m.node_middles, m.light_orders, m.light_mapping, m.bounds_min, m.bounds_max, &m};
}

__device__ __forceinline__ V3 meshPosition(const MeshDevice& mesh, uint32_t index) {
    return {__ldg(mesh.positions + 3 * size_t(index)), __ldg(mesh.positions + 3 * size_t(index) + 1), __ldg(mesh.positions + 3 * size_t(index) + 2)};
}
__device__ __forceinline__ void meshTriangle(const MeshDevice& mesh, uint32_t t, V3& a, V3& b, V3& c) {
    a = meshPosition(mesh, __ldg(mesh.triangles + 3 * size_t(t)));
    b = meshPosition(mesh, __ldg(mesh.triangles + 3 * size_t(t) + 1));
    c = meshPosition(mesh, __ldg(mesh.triangles + 3 * size_t(t) + 2));
}

// MeshImpl.lightProperties, shape_sampler.zig:198-226
__device__ __forceinline__ LightPropsD meshLightProperties(const SceneDevice& sc, const MeshSamplerDevice& m, uint32_t light) {
    V3 a, b, c;
    meshTriangle(sc.meshes[m.mesh], __ldg(m.triangle_mapping + light), a, b, c);
    const V3    center = divs3(add3(add3(a, b), c), 3.f);
    const float sra    = squaredLength3(sub3(a, center));
    const float srb    = squaredLength3(sub3(b, center));
    const float src    = squaredLength3(sub3(c, center));
    const float radius = __fsqrt_rn(zmax(sra, zmax(srb, src)));
    const V3    nn     = normalize3(cross3(sub3(b, a), sub3(c, a)));
    return {center, radius, nn, 1.f, __ldg(m.triangle_pdfs + light), 0 != m.two_sided};
}

__device__ __forceinline__ float lightWeight(const SceneDevice& sc, const TreeD& tr, V3 p, V3 n, bool total_sphere, uint32_t light) {  // light_tree.zig:227-233
    const LightPropsD lp = tr.sampler ? meshLightProperties(sc, *tr.sampler, light) : lightProperties(sc, light);
    return lightImportance(p, n, lp.center, lp.cone_axis, lp.cone_cos, lp.radius, lp.power, lp.two_sided, total_sphere);
}

struct LightNodeD {
    V3       center;
    float    radius;
    V3       cone_axis;
    float    cone_cos;
    float    power, variance;
    uint32_t meta, num_lights;
};

__device__ __forceinline__ LightNodeD loadLightNode(const TreeD& sc, uint32_t i) {  // light_tree.zig:25-42
    const uint4* p  = reinterpret_cast<const uint4*>(sc.nodes + i);
    const uint4  a  = __ldg(p);
    const uint4  b  = __ldg(p + 1);
    const float  ku = 1.f / 65535.f;
    const float  tx = float(a.x & 0xffffu) * ku, ty = float(a.x >> 16) * ku, tz = float(a.y & 0xffffu) * ku, tw = float(a.y >> 16) * ku;
    LightNodeD   n;
    n.center    = {zlerp(sc.bounds_min.x, sc.bounds_max.x, tx), zlerp(sc.bounds_min.y, sc.bounds_max.y, ty),
                   zlerp(sc.bounds_min.z, sc.bounds_max.z, tz)};
    n.radius    = zlerp(sc.bounds_min.w, sc.bounds_max.w, tw);
    n.cone_axis = {__fmaf_rn(float(a.z & 0xffffu), 1.f / 32768.f, -1.f), __fmaf_rn(float(a.z >> 16), 1.f / 32768.f, -1.f),
                   __fmaf_rn(float(a.w & 0xffffu), 1.f / 32768.f, -1.f)};
    n.cone_cos  = __fmaf_rn(float(a.w >> 16), 1.f / 32768.f, -1.f);
    n.power     = __uint_as_float(b.x);
    n.variance  = __uint_as_float(b.y);
    n.meta      = b.z;
    n.num_lights = b.w;
    return n;
}

__device__ __forceinline__ float lightNodeWeight(const LightNodeD& node, V3 p, V3 n, bool total_sphere) {  // light_tree.zig:57-63
    return lightImportance(p, n, node.center, node.cone_axis, node.cone_cos, node.radius, node.power, 0 != (node.meta & 2u), total_sphere);
}

__device__ __forceinline__ bool lightNodeSplit(const LightNodeD& node, V3 p, float threshold) {  // light_tree.zig:65-89
    const float r = node.radius;
    const float d = zmin(length3(sub3(p, node.center)), 1.0e6f);
    const float a = zmax(d - r, 0.001f);
    const float b = d + r;

    const float eg  = __fdiv_rn(1.f, a * b);
    const float eg2 = eg * eg;
    const float a3  = a * a * a;
    const float b3  = b * b * b;
    const float e2g = __fdiv_rn(b3 - a3, 3.f * (b - a) * a3 * b3);
    const float vg  = e2g - eg2;

    const float ve = node.variance;
    const float ee = node.power;
    const float s2 = zmax(ve * vg + ve * eg2 + ee * ee * vg, 0.f);
    const float ns = __fdiv_rn(1.f, 1.f + __fsqrt_rn(s2));
    return ns < threshold;
}

struct LightPickD {
    uint32_t offset;
    float    pdf;
};

// Node.randomLight, light_tree.zig:91-145
__device__ __forceinline__ LightPickD lightNodeRandomLight(const SceneDevice& sc, const TreeD& tr, const LightNodeD& node, V3 p, V3 n, bool total_sphere,
                                           float random) {
    const uint32_t num_lights = node.num_lights;
    const uint32_t light      = node.meta >> 2;
    if (1 == num_lights) return {__ldg(tr.mapping + light), 1.f};

    uint32_t front = light;
    uint32_t back  = light + num_lights - 1;

    float w_front = lightWeight(sc, tr, p, n, total_sphere, __ldg(tr.mapping + front));
    float w_back  = lightWeight(sc, tr, p, n, total_sphere, __ldg(tr.mapping + back));

    float w_sum_front = w_front;
    float w_sum_back  = w_back;
    float w_sum       = 0.f;

    while (front != back) {
        w_sum = w_sum_front + w_sum_back;
        if (w_sum_front <= random * w_sum) {
            front += 1;
            if (front != back) {
                w_front = lightWeight(sc, tr, p, n, total_sphere, __ldg(tr.mapping + front));
                w_sum_front += w_front;
            } else {
                w_front = w_back;
            }
        } else {
            back -= 1;
            if (front != back) {
                w_back = lightWeight(sc, tr, p, n, total_sphere, __ldg(tr.mapping + back));
                w_sum_back += w_back;
            }
        }
    }
    if (0.f == w_sum) return {0, 0.f};
    return {__ldg(tr.mapping + front), __fdiv_rn(w_front, w_sum)};
}

// Node.pdf, light_tree.zig:147-170
__device__ __forceinline__ float lightNodePdf(const SceneDevice& sc, const TreeD& tr, const LightNodeD& node, V3 p, V3 n, bool total_sphere, uint32_t id) {
    const uint32_t num_lights = node.num_lights;
    if (1 == num_lights) return 1.f;
    const uint32_t light = node.meta >> 2;
    const uint32_t end   = light + num_lights;
    float          w_id = 0.f, sum = 0.f;
    for (uint32_t i = light; i < end; ++i) {
        const float lw = lightWeight(sc, tr, p, n, total_sphere, __ldg(tr.mapping + i));
        sum += lw;
        if (id == i) w_id = lw;
    }
    if (0.f == sum) return 0.f;
    return __fdiv_rn(w_id, sum);
}

constexpr uint32_t kMaxLightPicks = 64;  // Tree.MaxLights

// Tree.randomLight, light_tree.zig:346-447. `emit` is called for every pick in the reference's order.
template <typename Emit>
__device__ __forceinline__ void lightTreeRandomLight(const SceneDevice& sc, V3 p, V3 n, bool total_sphere, float random, float split_threshold, Emit&& emit) {
    float       ip    = 0.f;
    const bool  split = split_threshold > 0.f;
    const TreeD tr    = sceneTree(sc);

    if (split && sc.lt_num_infinite < kMaxLightPicks - 1) {
        for (uint32_t i = 0; i < sc.lt_num_infinite; ++i) emit(LightPickD{__ldg(sc.lt_mapping + i), 1.f});
    } else {
        ip = sc.lt_infinite_weight;
        if (random < sc.lt_infinite_guard) {
            // infinite_light_distribution.sampleDiscrete(random), distribution_1d.zig:56-60
            const uint32_t l = dist1dSample(sc.lt_infinite_cdf, sc.lt_num_infinite + 1, random);
            emit(LightPickD{__ldg(sc.lt_mapping + l), (__ldg(sc.lt_infinite_cdf + l + 1) - __ldg(sc.lt_infinite_cdf + l)) * ip});
            return;
        }
    }
    if (0 == sc.lt_num_nodes) return;

    const float    pd              = 1.f - ip;
    const uint32_t max_split_depth = sc.lt_max_split_depth;

    struct Value {
        float    pdf, random;
        uint32_t node, depth;
    };
    Value    stack[12];
    uint32_t end = 0;

    Value t{pd, __fdiv_rn(random - ip, pd), 0, split ? 0 : max_split_depth};
    stack[end++] = t;

    while (end > 0) {
        const LightNodeD node = loadLightNode(tr, t.node);
        if (0 != (node.meta & 1u)) {
            const bool     do_split = t.depth < max_split_depth && lightNodeSplit(node, p, split_threshold);
            const uint32_t c0       = node.meta >> 2;
            const uint32_t c1       = c0 + 1;
            if (do_split) {
                t.depth += 1;
                t.node       = c0;
                stack[end++] = {t.pdf, t.random, c1, t.depth};
            } else {
                t.depth = max_split_depth;

                float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);

                const float pt = p0 + p1;
                if (0.f == pt) {
                    t = stack[--end];
                    continue;
                }
                p0 = __fdiv_rn(p0, pt);
                p1 = __fdiv_rn(p1, pt);
                if (t.random < p0) {
                    t.node = c0;
                    t.pdf *= p0;
                    t.random = __fdiv_rn(t.random, p0);
                } else {
                    t.node = c1;
                    t.pdf *= p1;
                    t.random = zmin(__fdiv_rn(t.random - p0, p1), 1.f);
                }
            }
        } else {
            const LightPickD pick = lightNodeRandomLight(sc, tr, node, p, n, total_sphere, t.random);
            if (pick.pdf > 0.f) emit(LightPickD{pick.offset, pick.pdf * t.pdf});
            t = stack[--end];
        }
    }
}

// PrimitiveTree.randomLight, light_tree.zig:577-650. `emit` receives (part triangle, pdf) in the reference's order.
template <typename Emit>
__device__ __forceinline__ void primitiveTreeRandomLight(const SceneDevice& sc, const MeshSamplerDevice& m, V3 p, V3 n, bool total_sphere, float random,
                                         float split_threshold, Emit&& emit) {
    constexpr uint32_t kMaxSplitDepth = 6;
    const TreeD        tr             = primitiveTree(m);
    const bool         split          = split_threshold > 0.f;

    struct Value {
        float    pdf, random;
        uint32_t node, depth;
    };
    Value    stack[kMaxSplitDepth + 1];
    uint32_t end = 0;

    Value t{1.f, random, 0, split ? 0 : kMaxSplitDepth};
    stack[end++] = t;

    while (end > 0) {
        const LightNodeD node = loadLightNode(tr, t.node);
        if (0 != (node.meta & 1u)) {
            const bool     do_split = t.depth < kMaxSplitDepth && lightNodeSplit(node, p, split_threshold);
            const uint32_t c0       = node.meta >> 2;
            const uint32_t c1       = c0 + 1;
            if (do_split) {
                t.depth += 1;
                t.node       = c0;
                stack[end++] = {t.pdf, t.random, c1, t.depth};
            } else {
                t.depth = kMaxSplitDepth;

                float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);

                const float pt = p0 + p1;
                if (0.f == pt) {
                    t = stack[--end];
                    continue;
                }
                p0 = __fdiv_rn(p0, pt);
                p1 = __fdiv_rn(p1, pt);
                if (t.random < p0) {
                    t.node = c0;
                    t.pdf *= p0;
                    t.random = __fdiv_rn(t.random, p0);
                } else {
                    t.node = c1;
                    t.pdf *= p1;
                    t.random = zmin(__fdiv_rn(t.random - p0, p1), 1.f);
                }
            }
        } else {
            const LightPickD pick = lightNodeRandomLight(sc, tr, node, p, n, total_sphere, t.random);
            if (pick.pdf > 0.f) emit(LightPickD{pick.offset, pick.pdf * t.pdf});
            t = stack[--end];
        }
    }
}

// PrimitiveTree.pdf, light_tree.zig:652-719
__device__ __forceinline__ float primitiveTreePdf(const SceneDevice& sc, const MeshSamplerDevice& m, V3 p, V3 n, bool total_sphere, float split_threshold,
                                  uint32_t id) {
    constexpr uint32_t kMaxSplitDepth = 6;
    const TreeD        tr             = primitiveTree(m);
    const uint32_t     lo             = __ldg(tr.orders + id);
    const bool         split          = split_threshold > 0.f;

    float    pd    = 1.f;
    uint32_t nid   = 0;
    uint32_t depth = split ? 0 : kMaxSplitDepth;
    for (;;) {
        const LightNodeD node = loadLightNode(tr, nid);
        if (0 != (node.meta & 1u)) {
            const bool     do_split = depth < kMaxSplitDepth && lightNodeSplit(node, p, split_threshold);
            const uint32_t c0       = node.meta >> 2;
            const uint32_t c1       = c0 + 1;
            const uint32_t middle   = __ldg(tr.middles + nid);
            if (do_split) {
                depth += 1;
                nid = lo < middle ? c0 : c1;
            } else {
                depth          = kMaxSplitDepth;
                const float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                const float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);
                const float pt = p0 + p1;
                if (0.f == pt) return 0.f;
                if (lo < middle) {
                    nid = c0;
                    pd *= __fdiv_rn(p0, pt);
                } else {
                    nid = c1;
                    pd *= __fdiv_rn(p1, pt);
                }
            }
        } else {
            return pd * lightNodePdf(sc, tr, node, p, n, total_sphere, lo);
        }
    }
}

// Tree.pdf, light_tree.zig:449-517
__device__ __forceinline__ float lightTreePdf(const SceneDevice& sc, V3 p, V3 n, bool total_sphere, float split_threshold, uint32_t id) {
    const TreeD    tr             = sceneTree(sc);
    const uint32_t lo             = __ldg(sc.lt_orders + id);
    const bool     split          = split_threshold > 0.f;
    const bool     split_infinite = split && sc.lt_num_infinite < kMaxLightPicks - 1;

    if (lo < sc.lt_infinite_end) {  // infinite_weight * infinite_light_distribution.pdfI(lo)
        return split_infinite ? 1.f : sc.lt_infinite_weight * (__ldg(sc.lt_infinite_cdf + lo + 1) - __ldg(sc.lt_infinite_cdf + lo));
    }
    if (0 == sc.lt_num_nodes) return 0.f;

    const float    ip              = split_infinite ? 0.f : sc.lt_infinite_weight;
    const uint32_t max_split_depth = sc.lt_max_split_depth;

    float    pd    = 1.f - ip;
    uint32_t nid   = 0;
    uint32_t depth = split ? 0 : max_split_depth;
    for (;;) {
        const LightNodeD node = loadLightNode(tr, nid);
        if (0 != (node.meta & 1u)) {
            const bool     do_split = depth < max_split_depth && lightNodeSplit(node, p, split_threshold);
            const uint32_t c0       = node.meta >> 2;
            const uint32_t c1       = c0 + 1;
            const uint32_t middle   = __ldg(sc.lt_middles + nid);
            if (do_split) {
                depth += 1;
                nid = lo < middle ? c0 : c1;
            } else {
                depth          = max_split_depth;
                const float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                const float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);
                const float pt = p0 + p1;
                if (0.f == pt) return 0.f;
                if (lo < middle) {
                    nid = c0;
                    pd *= __fdiv_rn(p0, pt);
                } else {
                    nid = c1;
                    pd *= __fdiv_rn(p1, pt);
                }
            }
        } else {
            return pd * lightNodePdf(sc, tr, node, p, n, total_sphere, lo);
        }
    }
}

// Mesh.pdf, triangle_mesh.zig:662-703. Inlined: only the MeshLights instances of the kernels contain it, and an out-of-line copy
// taking the scene by reference makes every thread copy the kernel's parameter structs to local memory.
__device__ __forceinline__ float meshLightPdf(const SceneDevice& sc, const MeshSamplerDevice& m, const VertexD& vertex, const FragD& frag) {
    const float n_dot_dir = fabsf(dot3(frag.geo_n, vertex.ray.d));

    const V3 op = frag.trafo.worldToObjectPoint(vertex.origin);
    const V3 on = frag.trafo.worldToObjectNormal(vertex.geo_n);

    const uint32_t pm      = __ldg(m.primitive_mapping + frag.primitive);
    const float    tri_pdf = primitiveTreePdf(sc, m, op, on, 0 != (vertex.state & kTranslucent), vertex.light_split_threshold, pm);

    V3 a, b, c;
    meshTriangle(sc.meshes[m.mesh], frag.primitive, a, b, c);
    const V3    ca       = mul3(mul3(frag.trafo.scale, frag.trafo.scale), cross3(sub3(b, a), sub3(c, a)));
    const float tri_area = 0.5f * length3(ca);
    const V3    center   = divs3(add3(add3(a, b), c), 3.f);

    if (__fdiv_rn(tri_area, length3(sub3(center, op))) > kAreaDistanceRatio) return tri_pdf * pdfSpherical(op, a, b, c);
    const float sl = squaredLength3(sub3(vertex.origin, frag.p));
    return __fdiv_rn(tri_pdf * sl, n_dot_dir * tri_area);
}

// Mesh.sampleTo, triangle_mesh.zig:492-608: appends the shadow records of one picked mesh light, returns the new record count.
// Inlined for the same reason.
__device__ __forceinline__ uint32_t meshLightSampleTo(const SceneDevice& sc, const PathState& st, uint32_t slot, const ZygpuLight& light,
                                                   LightPickD pick, const TrafoD& trafo, const FragD& frag, V3 n, bool translucent,
                                                   float split_threshold, SamplerD& sampler, uint32_t num_records) {
    const MeshSamplerDevice& m  = sc.mesh_samplers[light.sampler];
    const V3                 p  = frag.p;
    const V3                 op = trafo.worldToObjectPoint(p);
    const V3                 on = trafo.worldToObjectNormal(n);
    const V3    scale_squared   = mul3(trafo.scale, trafo.scale);
    const float r1              = sampler.sample1D();
    primitiveTreeRandomLight(sc, m, op, on, translucent, r1, split_threshold, [&](LightPickD sp) {
        V3 a, b, c;
        meshTriangle(sc.meshes[m.mesh], __ldg(m.triangle_mapping + sp.offset), a, b, c);

        const V3    ca  = mul3(scale_squared, cross3(sub3(b, a), sub3(c, a)));
        const float lca = length3(ca);
        V3          wn  = trafo.objectToWorldNormal(divs3(ca, lca));

        const float tri_area = 0.5f * lca;
        const V3    center   = divs3(add3(add3(a, b), c), 3.f);

        float u0, u1;
        sampler.sample2D(u0, u1);

        V3    dir, v;
        float sample_pdf, n_dot_dir;
        if (__fdiv_rn(tri_area, length3(sub3(center, op))) > kAreaDistanceRatio) {
            V3    sdir;
            float bu, bv, spdf;
            if (!sampleSpherical(op, a, b, c, u0, u1, sdir, bu, bv, spdf)) return;
            if (dot3(sdir, on) <= 0.f && !translucent) return;
            dir        = trafo.objectToWorldNormal(sdir);
            v          = trafo.objectToWorldPoint(interpolate3(a, b, c, bu, bv));
            sample_pdf = sp.pdf * spdf;
            if (0 != light.two_sided && dot3(wn, dir) > 0.f) wn = neg3(wn);
            n_dot_dir = -dot3(wn, dir);
        } else {
            float bu, bv;
            triangleUniform(u0, u1, bu, bv);
            v = trafo.objectToWorldPoint(interpolate3(a, b, c, bu, bv));

            const V3    axis = sub3(v, p);
            const float sl   = squaredLength3(axis);
            const float d    = __fsqrt_rn(sl);
            dir              = divs3(axis, d);
            if (dot3(dir, n) <= 0.f && !translucent) return;
            if (0 != light.two_sided && dot3(wn, dir) > 0.f) wn = neg3(wn);
            n_dot_dir  = -dot3(wn, dir);
            sample_pdf = __fdiv_rn(sp.pdf * sl, n_dot_dir * tri_area);
        }
        if (n_dot_dir < kDotMin) return;

        if (num_records < st.shadow_stride) {
            const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
            const V3     origin    = frag.offsetP(dir);
            const V3     light_pos = offsetRay(v, wn);
            st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, sample_pdf * pick.pdf);
            st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
            st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
            num_records += 1;
        } else {
            st.counters[3] = 1;
        }
    });
    return num_records;
}

// Scene.lightPdf, scene.zig:624-634 (+ Light.pdf -> Shape.pdf, light.zig:149-157, shape.zig:469-492, rectangle.zig:554-575)
// `MeshLights`: the scene has triangle-mesh lights (their sampling code is compiled out otherwise)
template <bool MeshLights>
__device__ __forceinline__ float sceneLightPdf(const SceneDevice& sc, const VertexD& vertex, const FragD& frag) {
    const uint32_t light_id = __ldg(sc.light_ids + sc.props[frag.prop].parts_start + frag.part);
    if (0 != (vertex.state & kSingular) || ZYGPU_NULL == light_id) return 1.f;

    const float select_pdf =
        lightTreePdf(sc, vertex.origin, vertex.geo_n, 0 != (vertex.state & kTranslucent), vertex.light_split_threshold, light_id);

    const ZygpuLight l          = sc.lights[light_id];
    float            sample_pdf = 0.f;
    if (ZYG_SHAPE_RECTANGLE == sc.props[l.prop].shape && ZYG_LIGHT_PROP_IMAGE == l.light_class) {  // Rectangle.materialPdf
        const float c            = fabsf(dot3(frag.trafo.r2, vertex.ray.d));
        const float area         = frag.trafo.scale.x * frag.trafo.scale.y;
        const float sl           = squaredLength3(sub3(vertex.origin, frag.p));
        const float material_pdf = imagePdf(sc.image_samplers[l.sampler], frag.u, frag.v) * float(lightNumSamples(l, vertex.light_split_threshold));
        sample_pdf               = __fdiv_rn(material_pdf * sl, c * area);
    } else if (ZYG_SHAPE_RECTANGLE == sc.props[l.prop].shape) {
        const float nsf = float(lightNumSamples(l, vertex.light_split_threshold));
        SphQuadD    squad;
        squad.init(frag.trafo.scale, frag.trafo.worldToFramePoint(vertex.origin));
        sample_pdf = nsf * squad.pdf(frag.trafo.scale);
    } else if (ZYG_SHAPE_DISTANT == sc.props[l.prop].shape) {  // Distant.pdf, distant.zig:139-141
        sample_pdf = __fdiv_rn(1.f, distantSolidAngle(frag.trafo.scale.x));
    } else if (ZYG_SHAPE_SPHERE == sc.props[l.prop].shape) {  // Sphere.pdf, sphere.zig:472-487
        sample_pdf = float(lightNumSamples(l, vertex.light_split_threshold)) * sphereLightPdf(frag.trafo, vertex.origin);
    } else if (ZYG_SHAPE_CANOPY == sc.props[l.prop].shape) {  // Light.propMaterialPdf -> Shape.materialPdf, shape.zig:519
        if (ZYG_LIGHT_PROP_IMAGE == l.light_class) sample_pdf = __fdiv_rn(imagePdf(sc.image_samplers[l.sampler], frag.u, frag.v), 2.f * kPi);
    } else if (MeshLights && ZYG_SHAPE_TRIANGLE_MESH == sc.props[l.prop].shape && ZYGPU_NULL != l.sampler) {
        sample_pdf = meshLightPdf(sc, sc.mesh_samplers[l.sampler], vertex, frag);
    }
    return powerHeuristic(vertex.bxdf_pdf, sample_pdf * select_pdf);
}

// Vertex.evaluateRadiance, vertex.zig:183-212
template <bool MeshLights>
__device__ __forceinline__ V3 evaluateRadiance(const SceneDevice& sc, const VertexD& vertex, const FragD& frag, SamplerD& sampler) {
    const V3            wo = neg3(vertex.ray.d);
    const ZygpuMaterial m  = sc.materials[__ldg(sc.material_ids + sc.props[frag.prop].parts_start + frag.part)];
    if (0 == (m.flags & ZYG_MATERIAL_EMISSIVE) || (0 == (m.flags & ZYG_MATERIAL_TWO_SIDED) && !frag.sameHemisphere(wo))) {
        return splat3(0.f);
    }
    const float stochastic_r = sampler.sample1D();  // rs.stochastic_r

    const bool  in_camera = 0 == vertex.probe_depth;
    const float area      = 0.f != m.emission_normalize ? shapeArea(sc.props[frag.prop].shape, frag.trafo.scale) : 1.f;
    const V3    energy    = ZYGPU_NULL != m.emission_map
                                ? emittanceRadianceMapped(m, wo, frag.trafo, area, in_camera,
                                                          imageTexel(sc.image_samplers[m.emission_map], frag.u, frag.v, stochastic_r))
                                : emittanceRadiance(m, wo, frag.trafo, area, in_camera);
    const float weight    = sceneLightPdf<MeshLights>(sc, vertex, frag);
    return scale3(weight, energy);
}

// Mesh.emission -> Tree.emission, triangle_mesh.zig:379-388, triangle_tree.zig:405-477: every triangle of an un-occluding mesh
// emitter the segment crosses, visited in the reference's order (binary tree, near child first) because each hit draws from
// the sampler.
template <bool MeshLights>
__device__ __forceinline__ V3 meshEmission(const SceneDevice& sc, uint32_t entity, const ZygpuProp& prop, const VertexD& vertex, SamplerD& sampler) {
    FragD frag;
    frag.prop  = entity;
    frag.trafo = loadTrafo(sc.trafos, entity);
    frag.t = frag.b = frag.n = splat3(0.f);
    frag.u = frag.v = 0.f;

    const MeshDevice&  mesh  = sc.meshes[prop.mesh];
    const MeshShading& shade = sc.mesh_shading[prop.mesh];
    const RayT         ray   = worldToObjectRay(frag.trafo, vertex.ray);

    uint32_t stack[64];
    uint32_t end = 0;
    uint32_t n   = 0;
    V3       energy = splat3(0.f);

    while (kEnd != n) {
        const float4 nmin = __ldg(mesh.binary_nodes + 2 * size_t(n));
        const float4 nmax = __ldg(mesh.binary_nodes + 2 * size_t(n) + 1);
        const uint32_t num = __float_as_uint(nmax.w);
        if (0 != num) {
            const uint32_t start = __float_as_uint(nmin.w);
            for (uint32_t i = start; i < start + num; ++i) {
                V3 a, b, c;
                meshTriangle(mesh, i, a, b, c);
                float ht, hu, hv;
                if (!intersectTriangle(ray, a, sub3(b, a), sub3(c, a), ht, hu, hv)) continue;
                frag.primitive = i;
                frag.part      = __ldg(shade.parts + i);
                frag.p         = frag.trafo.objectToWorldPoint(interpolate3(a, b, c, hu, hv));
                frag.geo_n     = frag.trafo.objectToWorldNormal(normalize3(cross3(sub3(b, a), sub3(c, a))));
                energy         = add3(energy, evaluateRadiance<MeshLights>(sc, vertex, frag, sampler));
            }
            n = 0 == end ? kEnd : stack[--end];
            continue;
        }
        uint32_t a = __float_as_uint(nmin.w);
        uint32_t b = a + 1;
        float dista = intersectNode(__ldg(mesh.binary_nodes + 2 * size_t(a)), __ldg(mesh.binary_nodes + 2 * size_t(a) + 1), ray);
        float distb = intersectNode(__ldg(mesh.binary_nodes + 2 * size_t(b)), __ldg(mesh.binary_nodes + 2 * size_t(b) + 1), ray);
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
            if (FLT_MAX != distb && end < 64) stack[end++] = b;
        }
    }
    return energy;
}

// Prop.emission + Shape.emission, prop.zig:239-264, shape.zig:283-299, rectangle.zig:188-196
template <bool MeshLights>
__device__ __forceinline__ V3 propEmission(const SceneDevice& sc, uint32_t entity, const VertexD& vertex, SamplerD& sampler) {
    const ZygpuProp prop = sc.props[entity];
    if (!propVisible(prop.flags, vertex.probe_depth)) return splat3(0.f);
    if (!aabbIntersect(sc.aabbs, entity, vertex.ray)) return splat3(0.f);
    if (MeshLights && ZYG_SHAPE_TRIANGLE_MESH == prop.shape) return meshEmission<MeshLights>(sc, entity, prop, vertex, sampler);
    if (ZYG_SHAPE_RECTANGLE != prop.shape && ZYG_SHAPE_SPHERE != prop.shape) return splat3(0.f);

    FragD frag;
    frag.prop  = entity;
    frag.trafo = loadTrafo(sc.trafos, entity);
    HitD isec;
    if (ZYG_SHAPE_SPHERE == prop.shape) {  // Sphere.emission, sphere.zig:271-279
        if (!sphereIntersect(vertex.ray, frag.trafo, isec)) return splat3(0.f);
        sphereFragment(vertex.ray, isec, frag);
    } else {
        if (!rectangleIntersect(vertex.ray, frag.trafo, isec)) return splat3(0.f);
        rectangleFragment(vertex.ray, isec, frag);
    }
    return evaluateRadiance<MeshLights>(sc, vertex, frag, sampler);
}

// Context.emission -> PropBvh.emission, prop_tree.zig:302-356: every un-occluding emitter crossed before the hit
template <bool MeshLights>
__device__ __forceinline__ V3 unoccludingEmission(const SceneDevice& sc, const VertexD& vertex, SamplerD& sampler) {
    uint32_t stack[kPropStack];
    uint32_t end = 0;
    uint32_t n   = 0 == sc.num_unocc_nodes ? kEnd : 0;

    V3 energy = splat3(0.f);

    while (kEnd != n) {
        const float4 nmin = __ldg(sc.unocc_nodes + 2 * size_t(n));
        const float4 nmax = __ldg(sc.unocc_nodes + 2 * size_t(n) + 1);

        const uint32_t num = __float_as_uint(nmax.w);
        if (0 != num) {
            const uint32_t start = __float_as_uint(nmin.w);
            for (uint32_t i = start; i < start + num; ++i) {
                energy = add3(energy, propEmission<MeshLights>(sc, __ldg(sc.unocc_indices + i), vertex, sampler));
            }
            n = 0 == end ? kEnd : stack[--end];
            continue;
        }

        uint32_t a = __float_as_uint(nmin.w);
        uint32_t b = a + 1;

        float dista = intersectNode(__ldg(sc.unocc_nodes + 2 * size_t(a)), __ldg(sc.unocc_nodes + 2 * size_t(a) + 1), vertex.ray);
        float distb = intersectNode(__ldg(sc.unocc_nodes + 2 * size_t(b)), __ldg(sc.unocc_nodes + 2 * size_t(b) + 1), vertex.ray);
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
    return energy;
}

// ---- accumulation (IValue.add, helper.zig:11-19) ---------------------------------------------

__device__ __forceinline__ void ivalueAdd(const PathState& st, uint32_t slot, V3 value, uint32_t depth, uint32_t direct_cutoff,
                                          bool is_emission, bool singular) {
    float4* target = is_emission ? st.acc_e : ((singular || depth < direct_cutoff) ? st.acc_d : st.acc_i);
    float4  a      = target[slot];
    a.x += value.x;
    a.y += value.y;
    a.z += value.z;
    target[slot] = a;
}

// ---- stage kernels ---------------------------------------------------------------------------

// Worker.render per sample: Sensor.cameraSample (sensor.zig:152-166) + Perspective.generateVertex
// (camera_perspective.zig:124-150) + Vertex.init (vertex.zig:67-85)
__global__ void __launch_bounds__(kBlock) generateKernel(ZygpuView view, PathState st, PassParams pass) {
    __shared__ __align__(16) uint32_t sobol_tables[kSobolTableWords];
    loadSobolTables(sobol_tables);
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < pass.num_paths; slot += gridDim.x * blockDim.x) {
        const SlotId id = slotId(slot, pass);

        SamplerD sampler;
        sampler.sobol.tables = sobol_tables;
        sampler.use_sobol    = ZYG_SAMPLER_SOBOL == view.sampler;
        seedSamplers(id, pass, view.spp_total, sampler.sobol, sampler.rng);

        const int32_t fr = view.filter_radius_int;
        const int32_t px = int32_t(id.pixel_id % pass.padded_w) - fr;
        const int32_t py = int32_t(id.pixel_id / pass.padded_w) - fr;

        float s4[4];
        if (sampler.use_sobol) {  // sample4D
            if (sampler.sobol.dimension >= 2) sampler.sobol.incrementSeed();
            const uint32_t d       = sampler.sobol.dimension;
            sampler.sobol.dimension = d + 4;
            s4[0] = sampler.sobol.buffer[d];
            s4[1] = sampler.sobol.buffer[d + 1];
            s4[2] = sampler.sobol.buffer[d + 2];
            s4[3] = sampler.sobol.buffer[d + 3];
        } else {
            for (int i = 0; i < 4; ++i) s4[i] = sampler.rng.randomFloat();
        }
        (void)sampler.sample1D();  // shutter time: static scenes
        sampler.incrementPadding();

        const float c0 = float(px) + s4[0];
        const float c1 = float(py) + s4[1];

        const V3 left_top = {view.left_top[0], view.left_top[1], view.left_top[2]};
        const V3 d_x      = {view.d_x[0], view.d_x[1], view.d_x[2]};
        const V3 d_y      = {view.d_y[0], view.d_y[1], view.d_y[2]};

        V3 direction = add3(add3(left_top, scale3(c0, d_x)), scale3(c1, d_y));
        V3 origin;
        if (view.aperture_radius > 0.f) {
            float lx, ly;
            diskConcentric(s4[2], s4[3], lx, ly);  // Aperture.sample, aperture.zig:46-53
            origin         = {lx * view.aperture_radius, ly * view.aperture_radius, 0.f};
            const float t  = __fdiv_rn(view.focus_distance, direction.z);
            const V3 focus = scale3(t, direction);
            direction      = sub3(focus, origin);
        } else {
            origin = {view.eye_offset[0], view.eye_offset[1], view.eye_offset[2]};
        }

        const TrafoD trafo = {{view.camera_trafo.r[0][0], view.camera_trafo.r[0][1], view.camera_trafo.r[0][2]},
                              {view.camera_trafo.r[1][0], view.camera_trafo.r[1][1], view.camera_trafo.r[1][2]},
                              {view.camera_trafo.r[2][0], view.camera_trafo.r[2][1], view.camera_trafo.r[2][2]},
                              {view.camera_trafo.r[0][3], view.camera_trafo.r[1][3], view.camera_trafo.r[2][3]},
                              {view.camera_trafo.position[0], view.camera_trafo.position[1], view.camera_trafo.position[2]}};

        const V3 origin_w    = trafo.objectToWorldPoint(origin);
        const V3 direction_w = trafo.objectToWorldVector(normalize3(direction));

        const uint32_t state = kPrimaryRay | kTransparent | kSingular;
        st.ray_o[slot]  = make_float4(origin_w.x, origin_w.y, origin_w.z, __uint_as_float(packFlags(state, 0, 0)));
        st.ray_d[slot]  = make_float4(direction_w.x, direction_w.y, direction_w.z, kRayMaxT);
        st.thr[slot]    = make_float4(1.f, 1.f, 1.f, 0.f);
        st.prev_p[slot] = make_float4(origin_w.x, origin_w.y, origin_w.z, 0.f);
        st.prev_n[slot] = make_float4(0.f, 0.f, 0.f, 1.f);  // split_weight = 1
        st.acc_e[slot]  = make_float4(0.f, 0.f, 0.f, s4[0]);
        st.acc_d[slot]  = make_float4(0.f, 0.f, 0.f, s4[1]);
        st.acc_i[slot]  = make_float4(0.f, 0.f, 0.f, 0.f);
        storeSampler(st, slot, sampler, kPoolFirst);  // the camera vertex sits in lane 0 (vertex id == slot)
        st.queue_a[slot] = slot;
        if (st.lanes > 1) st.queue_t[slot] = slot;
    }
    if (0 == blockIdx.x && 0 == threadIdx.x) {
        st.counters[0] = pass.num_paths;
        st.counters[1] = 0;
        st.counters[4] = 0;
        st.counters[7]  = pass.num_paths;
        st.counters[9]  = 0;
        st.counters[10] = 0;
        st.counters[11] = 0;
    }
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

__device__ __forceinline__ const ZygpuMaterial& propMaterial(const SceneDevice& sc, uint32_t prop, uint32_t part) {
    return sc.materials[__ldg(sc.material_ids + sc.props[prop].parts_start + part)];
}

// Vertex.iorOutside (vertex.zig:87-93) and Stack.highestPriority (medium.zig:74-82) for Vertex.sample
__device__ __forceinline__ void mediaForSample(const SceneDevice& sc, const MediaD& media, const FragD& frag, V3 wo, float& ior_outside,
                                               int& highest_priority) {
    ior_outside      = 1.f;
    highest_priority = -128;
    if (0 == media.count) return;
    for (uint32_t i = 0; i < media.count; ++i) highest_priority = max(highest_priority, propMaterial(sc, media.prop[i], media.part[i]).priority);
    const uint32_t back = media.count - 1;
    if (frag.sameHemisphere(wo)) {  // Stack.topIor
        ior_outside = propMaterial(sc, media.prop[back], media.part[back]).ior;
    } else if (media.count > 1) {  // Stack.peekIor
        const uint32_t i = (media.prop[back] == frag.prop && media.part[back] == frag.part) ? back - 1 : back;
        ior_outside      = propMaterial(sc, media.prop[i], media.part[i]).ior;
    }
}

// Pool.maxSplits, vertex.zig:306-309
__device__ __forceinline__ uint32_t maxSplits(uint32_t path_count_log2, bool primary_ray, uint32_t depth) {
    const uint32_t m = 4u >> path_count_log2;
    return m - (primary_ray ? 0u : min(depth, m - 1u));
}

struct LoadedVertex {
    VertexD  v;
    V3       throughput;
    float    reg_alpha;
    float    split_weight;
    uint32_t vertex_depth;
    uint32_t path_count_log2, num_media;
    HitD     isec;
    uint32_t prop;
};

__device__ __forceinline__ LoadedVertex loadVertex(const PathState& st, uint32_t slot /* vertex id */) {
    const float4 o  = st.ray_o[slot];
    const float4 d  = st.ray_d[slot];
    const float4 t  = st.thr[slot];
    const float4 pp = st.prev_p[slot];
    const float4 pn = st.prev_n[slot];
    const float4 h  = st.hit[slot];

    LoadedVertex    r;
    const uint32_t  flags = __float_as_uint(o.w);
    r.v.ray               = makeRay({o.x, o.y, o.z}, {d.x, d.y, d.z}, 0.f, d.w);
    r.v.origin            = {pp.x, pp.y, pp.z};
    r.v.geo_n             = {pn.x, pn.y, pn.z};
    r.v.bxdf_pdf          = t.w;
    r.v.light_split_threshold = 0.f;  // set by the stages before it is read
    r.split_weight        = pn.w;
    r.v.state             = flags & 0xffu;
    r.v.probe_depth       = (flags >> 8) & 0xffu;
    r.vertex_depth        = (flags >> 16) & 0xffu;
    r.path_count_log2     = (flags >> 24) & 3u;
    r.num_media           = (flags >> 26) & 3u;
    r.throughput          = {t.x, t.y, t.z};
    r.reg_alpha           = pp.w;
    r.isec                = {d.w, h.x, h.y, __float_as_uint(h.z)};
    r.prop                = __float_as_uint(h.w);
    return r;
}

// PathtracerMIS.li up to the shadow rays: connectLight (pathtracer_mis.zig:280-341), termination (:76-86), Russian
// roulette (:88, helper.zig:75-89), Vertex.sample (:93), sampleLights / evaluateLight up to the visibility test
// (:174-250).
// Features the scene needs of shade_a; what it does not need is compiled out (each costs registers in the hottest kernel).
// kFeatureTextured: a material reads image maps per vertex (colour / roughness / metallic / normal): instances without it hold none of
// that code (as a run-time branch it cost the map-free scenes 4 % through registers and code size)
enum : uint32_t { kFeatureSplit = 1, kFeatureMeshLights = 2, kFeatureInfiniteLights = 4, kFeatureDeferredLights = 8, kFeatureTextured = 16 };


// hlp.sampleNormal, material_helper.zig:16-79 for a UV-mapped normal map: the tangent-space normal of the map in the shading frame, then the
// adaption that keeps the reflection of `wo` above the geometric surface.
__device__ __forceinline__ V3 sampleNormal(V3 wo, V3 t, V3 b, V3 n, V3 geo_n, float nx, float ny) {
    const float nz = __fsqrt_rn(zmax(1.f - (nx * nx + ny * ny), 0.01f));
    // rs.tangentToWorld(nm), renderstate.zig:51-58
    const V3 w  = {(nx * t.x + ny * b.x) + nz * n.x, (nx * t.y + ny * b.y) + nz * n.y, (nx * t.z + ny * b.z) + nz * n.z};
    const V3 nn = normalize3(w);

    const V3    r = sub3(scale3(2.f * dot3(wo, nn), nn), wo);  // math.reflect3(n, wo), vector4.zig:94-96
    const float a = dot3(geo_n, r);
    if (a >= 0.f) return nn;
    if (dot3(geo_n, wo) < 0.0017453f) return geo_n;  // cos(89.9 degrees)

    const float bb      = dot3(geo_n, nn);
    const float epsilon = 1e-4f;
    V3          tangent = nn;
    if (bb > epsilon) {
        const float distance = __fdiv_rn(fabsf(a), bb);
        tangent              = normalize3(add3(r, scale3(distance, nn)));
    }
    tangent = add3(tangent, scale3(epsilon, geo_n));
    return normalize3(add3(wo, tangent));
}

__device__ __forceinline__ V3 surfaceMapTexel(const ImageSamplerDevice* is, float u, float v, float r) { return imageTexel(*is, u, v, r); }

// Material.sample with the image maps of a Substitute (substitute_material.zig:114-162): colour, roughness, metallic and the normal map are
// all looked up with the vertex's one stochastic_r (texture_sampler.zig:22-76), so shade_b re-reads the texels shade_a read.
template <bool Split, bool Textured>
__device__ __forceinline__ MatSampleD texturedMaterialSample(const SceneDevice& sc, ZygpuMaterial& m, const FragD& frag, V3 wo,
                                                             float stochastic_r, float reg_weight, float reg_alpha, bool caustics,
                                                             float specular_threshold, float ior_outside, int highest_priority) {
    if (Textured) {
        if (ZYGPU_NULL != m.color_map) {  // ts.sample2D_3(self.color, rs, ...), :120
            const V3 c = imageTexel(sc.image_samplers[m.color_map], frag.u, frag.v, stochastic_r);
            m.color[0] = c.x, m.color[1] = c.y, m.color[2] = c.z;
        }
        if (ZYGPU_NULL != m.roughness_map) m.roughness = surfaceMapTexel(sc.image_samplers + m.roughness_map, frag.u, frag.v, stochastic_r).x;  // :122
        if (ZYGPU_NULL != m.metallic_map) m.metallic = surfaceMapTexel(sc.image_samplers + m.metallic_map, frag.u, frag.v, stochastic_r).x;     // :123
    }
    MatSampleD r = materialSample<Split>(m, frag, wo, reg_weight, reg_alpha, caustics, specular_threshold, ior_outside, highest_priority);
    if (Textured && ZYGPU_NULL != m.normal_map && kSampleSubstitute == r.kind) {  // :157-159: result.super.frame = Frame.init(n)
        const V3 xy = surfaceMapTexel(sc.image_samplers + m.normal_map, frag.u, frag.v, stochastic_r);
        const V3 n  = sampleNormal(wo, frag.t, frag.b, r.n, r.geo_n, xy.x, xy.y);
        V3       t, b;
        orthonormalBasis3(n, t, b);
        r.frame = {t, b, n};
    }
    return r;
}

template <uint32_t Features>
__global__ void __launch_bounds__(kBlock, ZYGPU_SHADE_BLOCKS) shadeAKernel(SceneDevice sc, ZygpuView view, PathState st, PassParams pass, uint32_t round) {
    constexpr bool Split      = 0 != (Features & kFeatureSplit);
    constexpr bool MeshLights = 0 != (Features & kFeatureMeshLights);
    constexpr bool Infinite   = 0 != (Features & kFeatureInfiniteLights);
    constexpr bool Deferred   = 0 != (Features & kFeatureDeferredLights);  // light selection and sampling run in the light kernels
    constexpr bool Textured   = 0 != (Features & kFeatureTextured);
    __shared__ __align__(16) uint32_t sobol_tables[kSobolTableWords];
    const bool      later = Split && round > 0;
    const uint32_t  count = later ? st.counters[9] : st.counters[0];
    if (blockIdx.x * blockDim.x >= count) return;  // no item of any iteration falls to this block: skip the 20 KB table load
    loadSobolTables(sobol_tables);
    const uint32_t* __restrict__ queue = later ? st.queue_s : st.queue_a;
    const uint32_t  iters = (count + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t i      = it * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
        bool           alive  = false;
        bool           multi  = false;
        uint32_t       slot   = 0;
        if (i < count) {
            slot            = queue[i];
            const uint4 smp = st.smp[slot];
            uint32_t    pool = smp.w;
            uint32_t    lane = 0;
            bool        mine = true;  // the slot has a vertex for this round
            if (Split) {
                if (0 == round) {
                    pool  = poolSwap(pool);
                    multi = poolCurCount(pool) > 1;
                }
                mine = poolCurCount(pool) > round;
                lane = poolCurLane(pool, round);
            }
            if (mine) {
            const uint32_t vid = lane * st.capacity + slot;
            LoadedVertex lv = loadVertex(st, vid);
            VertexD&     vertex = lv.v;

            const uint32_t total_depth = vertex.probe_depth;
            const bool     hit         = kEnd != lv.prop;

            SamplerD sampler;
            sampler.sobol.tables = sobol_tables;
            loadSampler(st, slot, smp, pass, view.spp_total, total_depth, sampler);
            if (ZYG_SAMPLER_SOBOL != view.sampler) sampler.use_sobol = false;

            FragD frag;
            frag.prop = kEnd;
            if (hit) shapeFragment(sc, lv.prop, vertex.ray, lv.isec, frag);

            // Context.nextEvent inside a medium: VolumeIntegrator.integrate -> propScatter, the "glass" case
            // (volume_integrator.zig:51-66, 84-130): throughput *= exp(-mu_a * d); topCC = the highest-priority medium
            MediaD media{0, {0, 0, 0}, {0, 0, 0}};
            if (Split && 0 != lv.num_media) {
                media = unpackMedia(st.med[vid], lv.num_media);
                if (hit) {
                    int      priority = -128;
                    uint32_t highest  = 0;
                    for (uint32_t m = 0; m < media.count; ++m) {
                        const int lp = propMaterial(sc, media.prop[m], media.part[m]).priority;
                        if (lp >= priority) {
                            priority = lp;
                            highest  = m;
                        }
                    }
                    const ZygpuMaterial& mm = propMaterial(sc, media.prop[highest], media.part[highest]);
                    const float          nd = -(vertex.ray.tmax - vertex.ray.tmin);
                    lv.throughput = mul3(lv.throughput, {expf(nd * mm.color[0]), expf(nd * mm.color[1]), expf(nd * mm.color[2])});
                }
            }

            if (slot == pass.debug_slot) {
                printf("[gpu] depth %u pc %u sw %g state p%d s%d sg%d | hit prop %u t %.9g media %u | thr %.9g %.9g %.9g | rng %llx round %u lane %u\n",
                       total_depth, 1u << lv.path_count_log2, lv.split_weight, int(0 != (vertex.state & kPrimaryRay)),
                       int(0 != (vertex.state & kSpecular)), int(0 != (vertex.state & kSingular)), lv.prop, vertex.ray.tmax, lv.num_media,
                       lv.throughput.x, lv.throughput.y, lv.throughput.z, (unsigned long long)sampler.rng.state, round, lane);
            }

            // connectLight
            V3 this_light = splat3(0.f);
            if (!(0 == view.caustics_path && 0 != (vertex.state & kSpecular) && 0 == (vertex.state & kPrimaryRay))) {
                vertex.light_split_threshold = splitThreshold(view.split_threshold, lv.vertex_depth);
                if (hit) this_light = evaluateRadiance<MeshLights>(sc, vertex, frag, sampler);
                this_light = add3(this_light, unoccludingEmission<MeshLights>(sc, vertex, sampler));
                if (Infinite && kRayMaxT == vertex.ray.tmax) {  // the ray left the scene: infinite props, pathtracer_mis.zig:313-338
                    for (uint32_t k = 0; k < sc.num_infinite_props; ++k) {
                        const uint32_t  entity = __ldg(sc.infinite_props + k);
                        const ZygpuProp iprop  = sc.props[entity];
                        if (!propVisible(iprop.flags, vertex.probe_depth) || !aabbIntersect(sc.aabbs, entity, vertex.ray)) continue;
                        FragD light_frag;
                        light_frag.prop  = entity;
                        light_frag.trafo = loadTrafo(sc.trafos, entity);
                        HitD isec;
                        if (ZYG_SHAPE_DISTANT == iprop.shape) {
                            if (!distantIntersect(vertex.ray, light_frag.trafo, isec)) continue;
                            distantFragment(vertex.ray, isec, light_frag);
                        } else if (ZYG_SHAPE_CANOPY == iprop.shape) {
                            if (!canopyIntersect(vertex.ray, light_frag.trafo, isec)) continue;
                            canopyFragment(vertex.ray, light_frag);
                        } else {
                            continue;
                        }
                        this_light = add3(this_light, evaluateRadiance<MeshLights>(sc, vertex, light_frag, sampler));
                    }
                }
            }

            const V3 split_throughput = scale3(lv.split_weight, lv.throughput);
            ivalueAdd(st, slot, mul3(split_throughput, this_light), total_depth, 2, 0 == total_depth, 0 != (vertex.state & kSingular));

            bool terminate = !hit || vertex.probe_depth >= view.max_depth_surface || 0 >= view.max_depth_volume;

            if (!terminate) {
                // russianRoulette
                const float r  = sampler.sample1D();
                const float mx = hmax3(lv.throughput);
                const float q  = __fdiv_rn(mx, 0.1f);
                if (q < 1.f) {
                    if (r >= q) {
                        terminate = true;
                    } else {
                        lv.throughput = divs3(lv.throughput, q);
                    }
                }
            }

            if (!terminate) {
                const bool caustics = 0 == (vertex.state & kPrimaryRay) ? 0 != view.caustics_path : true;  // causticsResolve, :343-349

                const V3      wo = neg3(vertex.ray.d);
                ZygpuMaterial m  = sc.materials[__ldg(sc.material_ids + sc.props[frag.prop].parts_start + frag.part)];
                const float   stochastic_r = sampler.sample1D();  // rs.stochastic_r, vertex.zig:165
                if (Textured) st.stoch[vid] = stochastic_r;
                float ior_outside      = 1.f;
                int   highest_priority = -128;
                if (Split) mediaForSample(sc, media, frag, wo, ior_outside, highest_priority);
                const MatSampleD mat_sample = texturedMaterialSample<Split, Textured>(sc, m, frag, wo, stochastic_r, view.regularize_roughness, lv.reg_alpha,
                                                                                      caustics, view.specular_threshold, ior_outside, highest_priority);

                vertex.light_split_threshold = splitThreshold(view.split_threshold, vertex.probe_depth);

                // sampleLights. Every path owns `shadow_stride` consecutive shadow records (slot * stride + k): picks and
                // their samples are generated in the reference's order and stored in that order.
                uint32_t num_records = 0;
                bool     request     = false;
                if (Deferred && mat_sample.can_evaluate) {
                    const float    select = sampler.sample1D();
                    const uint32_t bits   = (mat_sample.translucent ? 1u : 0u) | (dot3(mat_sample.geo_n, frag.geo_n) < 0.f ? 2u : 0u) |
                                          (vertex.light_split_threshold != view.split_threshold ? 4u : 0u) | (total_depth << 8);
                    st.ls_p[slot] = make_float4(frag.p.x, frag.p.y, frag.p.z, select);
                    st.ls_g[slot] = make_float4(frag.geo_n.x, frag.geo_n.y, frag.geo_n.z, __uint_as_float(bits));
                    request       = true;
                }
                if (!Deferred && mat_sample.can_evaluate) {
                    const V3    p           = frag.p;
                    const V3    n           = mat_sample.geo_n;
                    const bool  translucent = mat_sample.translucent;
                    const float select      = sampler.sample1D();

                    // the picks first, then one loop over them: with the sampling code inside the callback the compiler keeps one
                    // out-of-line copy of it per call site of Tree.randomLight and a closure of references, which moves the scene and
                    // path records into local memory (measured: shade_a of the instanced scene 0.8 -> 1.6 ms per launch)
                    LightPickD     picks[kMaxLightPicks];
                    uint32_t       num_picks = 0;
                    lightTreeRandomLight(sc, p, n, translucent, select, vertex.light_split_threshold, [&](LightPickD pick) {
                        if (num_picks < kMaxLightPicks) picks[num_picks++] = pick;
                    });
                    for (uint32_t pi = 0; pi < num_picks; ++pi) {
                        const LightPickD pick = picks[pi];
                        const ZygpuLight light = sc.lights[pick.offset];
                        const TrafoD     trafo = loadTrafo(sc.trafos, light.prop);
                        const uint32_t   shape = sc.props[light.prop].shape;
                        if (Infinite && ZYG_SHAPE_DISTANT == shape) {  // Distant.sampleTo, distant.zig:78-107
                            const float radius = trafo.scale.x;
                            if (radius <= 0.f) continue;
                            float u0, u1;
                            sampler.sample2D(u0, u1);
                            float lx, ly;
                            diskConcentric(u0, u1, lx, ly);
                            const V3 ws  = scale3(radius, trafo.transformVector({lx, ly, 0.f}));
                            const V3 dir = normalize3(sub3(ws, trafo.r2));
                            if (dot3(dir, n) <= 0.f && !translucent) continue;
                            if (num_records < st.shadow_stride) {
                                const size_t rec    = size_t(slot) * st.shadow_stride + num_records;
                                const V3     origin = frag.offsetP(dir);
                                const float  pdf    = __fdiv_rn(1.f, distantSolidAngle(radius));
                                st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                                st.sh_p[rec]  = make_float4(0.f, 0.f, 0.f, __uint_as_float(pick.offset | 0x80000000u));
                                st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                num_records += 1;
                            } else {
                                st.counters[3] = 1;
                            }
                            continue;
                        }
                        if (Infinite && ZYG_SHAPE_CANOPY == shape) {  // Canopy.sampleMaterialTo, canopy.zig:94-131
                            if (ZYG_LIGHT_PROP_IMAGE != light.light_class) continue;
                            float u0, u1;
                            sampler.sample2D(u0, u1);
                            V3    dir;
                            float su, sv, pdf;
                            if (!canopySampleMaterialTo(sc.image_samplers[light.sampler], trafo, n, translucent, u0, u1, dir, su, sv, pdf)) continue;
                            if (num_records < st.shadow_stride) {
                                const size_t rec    = size_t(slot) * st.shadow_stride + num_records;
                                const V3     origin = frag.offsetP(dir);
                                st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                                st.sh_p[rec]  = make_float4(su, sv, 0.f, __uint_as_float(pick.offset | 0x80000000u));  // uvw of the sample
                                st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                num_records += 1;
                            } else {
                                st.counters[3] = 1;
                            }
                            continue;
                        }
                        if (MeshLights && ZYG_SHAPE_TRIANGLE_MESH == shape && ZYGPU_NULL != light.sampler) {
                            num_records = meshLightSampleTo(sc, st, slot, light, pick, trafo, frag, n, translucent,
                                                            vertex.light_split_threshold, sampler, num_records);
                            continue;
                        }
                        if (ZYG_SHAPE_SPHERE == shape) {  // Sphere.sampleTo, sphere.zig:323-393
                            SphereLightD sl;
                            sl.init(trafo, p);
                            if (!sl.valid) continue;
                            const uint32_t ns = lightNumSamples(light, vertex.light_split_threshold);
                            for (uint32_t k = 0; k < ns; ++k) {
                                float u0, u1;
                                sampler.sample2D(u0, u1);
                                V3    lp, wn, dir;
                                float pdf;
                                if (!sl.sample(trafo, p, n, translucent, u0, u1, lp, wn, dir, pdf)) continue;
                                if (num_records < st.shadow_stride) {
                                    const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                                    const V3     origin    = frag.offsetP(dir);
                                    const V3     light_pos = offsetRay(lp, wn);
                                    st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, (float(ns) * pdf) * pick.pdf);
                                    st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                                    st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                    num_records += 1;
                                } else {
                                    st.counters[3] = 1;
                                }
                            }
                            continue;
                        }
                        if (ZYG_SHAPE_RECTANGLE == shape && ZYG_LIGHT_PROP_IMAGE == light.light_class) {  // Rectangle.sampleMaterialTo
                            const uint32_t            ns   = lightNumSamples(light, vertex.light_split_threshold);
                            const ImageSamplerDevice& is   = sc.image_samplers[light.sampler];
                            const float               area = trafo.scale.x * trafo.scale.y;
                            for (uint32_t k = 0; k < ns; ++k) {
                                float u0, u1;
                                sampler.sample2D(u0, u1);
                                float su, sv, rs_pdf;
                                imageSample(is, u0, u1, su, sv, rs_pdf);
                                if (0.f == rs_pdf) continue;
                                const V3 ws   = trafo.objectToWorldPoint({-1.f * su + 0.5f, -1.f * sv + 0.5f, 0.f});
                                const V3 axis = sub3(ws, p);
                                V3       wn   = trafo.r2;
                                if (0 != light.two_sided && dot3(wn, axis) > 0.f) wn = neg3(wn);
                                const float sl  = squaredLength3(axis);
                                const float t   = __fsqrt_rn(sl);
                                const V3    dir = divs3(axis, t);
                                const float c   = -dot3(wn, dir);
                                if (c < kDotMin || (dot3(dir, n) <= 0.f && !translucent)) continue;
                                if (num_records < st.shadow_stride) {
                                    const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                                    const V3     origin    = frag.offsetP(dir);
                                    const V3     light_pos = offsetRay(ws, wn);
                                    st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, __fdiv_rn(float(ns) * rs_pdf * sl, c * area) * pick.pdf);
                                    st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                                    st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                    if (nullptr != st.sh_uv) st.sh_uv[rec] = make_float2(su, sv);
                                    num_records += 1;
                                } else {
                                    st.counters[3] = 1;
                                }
                            }
                            continue;
                        }
                        if (ZYG_SHAPE_RECTANGLE != shape) continue;

                        // Rectangle.sampleTo, rectangle.zig:305-357
                        const uint32_t ns  = lightNumSamples(light, vertex.light_split_threshold);
                        const float    nsf = float(ns);
                        SphQuadD       squad;
                        squad.init(trafo.scale, trafo.worldToFramePoint(p));
                        const float sample_pdf = nsf * squad.pdf(trafo.scale);

                        for (uint32_t k = 0; k < ns; ++k) {
                            float u0, u1;
                            sampler.sample2D(u0, u1);

                            const V3 ls  = squad.sample(u0, u1);
                            const V3 ws  = trafo.frameToWorldPoint(ls);
                            const V3 dir = normalize3(sub3(ws, p));

                            V3 wn = trafo.r2;
                            if (0 != light.two_sided && dot3(wn, dir) > 0.f) wn = neg3(wn);

                            if (-dot3(wn, dir) < kDotMin || 0.f == squad.S || (dot3(dir, n) <= 0.f && !translucent)) continue;

                            if (num_records < st.shadow_stride) {
                                const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                                const V3     origin    = frag.offsetP(dir);
                                const V3     light_pos = offsetRay(ws, wn);  // Shape.shadowRay, shape.zig:401-416
                                st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, sample_pdf * pick.pdf);
                                st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                                st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                                num_records += 1;
                            } else {
                                st.counters[3] = 1;  // more light samples than the host reserved: reported by zygpu_render
                            }
                        }
                    }
                }

                st.sh_n[slot] = num_records;
                st.thr[vid]   = make_float4(lv.throughput.x, lv.throughput.y, lv.throughput.z, vertex.bxdf_pdf);
                alive         = true;
                if (nullptr != st.queue_r && 0 != num_records) {
                    const uint32_t base = atomicAdd(&st.counters[10], num_records);
                    for (uint32_t k = 0; k < num_records; ++k) st.queue_r[base + k] = slot * st.shadow_stride + k;
                }
                if (Deferred && request) st.queue_l[atomicAdd(&st.counters[11], 1u)] = slot;
            }
            if (Split && !alive) pool = poolFree(pool, lane);
            storeSampler(st, slot, sampler, pool);
            }
        }
        queuePush(st.queue_b, &st.counters[1], alive, slot);
        if (Split && 0 == round) queuePush(st.queue_s, &st.counters[9], multi, slot);
    }
}

// ---- deferred light sampling ---------------------------------------------------------------------------------------
//
// With many lights the work of PathtracerMIS.sampleLights varies wildly between vertices: far from the lights the tree is
// descended once, near them the adaptive split returns dozens of picks, and a picked mesh light descends its own tree. Inside
// shade_a a warp would wait for its slowest lane (measured: 2.8 of 32 lanes active). So shade_a only leaves a request, and
// two persistent kernels whose lanes fetch new work as soon as they finish do the rest:
//
//   lightSelectPersistent   Tree.randomLight (light_tree.zig:346-447), one tree node per step and lane -> picks
//   lightSamplePersistent   Light.sampleTo for one pick per step and lane, in pick order -> shadow records
//
// The sampler is only touched by the second kernel, pick by pick in the reference's order, so the stream of a vertex is the
// one the inline path produces.

__device__ __forceinline__ V3 offsetPoint(V3 p, V3 geo_n, V3 w) {  // Fragment.offsetP with offset() == 0, intersection.zig:112-116
    const V3 nn = dot3(geo_n, w) > 0.f ? geo_n : neg3(geo_n);
    return offsetRay(fmas3(0.f, nn, p), nn);
}

__global__ void __launch_bounds__(128) lightSelectPersistent(SceneDevice sc, ZygpuView view, PathState st, uint32_t* __restrict__ work_counter) {
    constexpr uint32_t kFull   = 0xffffffffu;
    const uint32_t     lane    = threadIdx.x & 31u;
    const uint32_t     n_items = st.counters[11];
    const TreeD        tr      = sceneTree(sc);

    struct Value {
        float    pdf, random;
        uint32_t node, depth;
    };

    bool     active = false, exhausted = false;
    uint32_t slot = 0, num_picks = 0, end = 0;
    V3       p = {0.f, 0.f, 0.f}, n = {0.f, 0.f, 0.f};
    bool     total_sphere = false;
    float    threshold    = 0.f;
    Value    t{0.f, 0.f, 0, 0};
    Value    stack[12];

    const uint32_t max_split_depth = sc.lt_max_split_depth;

    for (;;) {
        const uint32_t idle = __ballot_sync(kFull, !active);
        if (0 != idle && !exhausted) {
            uint32_t base = 0;
            if (lane == uint32_t(__ffs(int(idle))) - 1u) base = atomicAdd(work_counter, uint32_t(__popc(idle)));
            base = __shfl_sync(kFull, base, __ffs(int(idle)) - 1);
            if (base + uint32_t(__popc(idle)) >= n_items) exhausted = true;
            const uint32_t index = base + uint32_t(__popc(idle & ((1u << lane) - 1u)));
            if (!active && index < n_items) {
                slot              = st.queue_l[index];
                const float4 lp   = st.ls_p[slot];
                const float4 lg   = st.ls_g[slot];
                const uint32_t fl = __float_as_uint(lg.w);
                p                 = {lp.x, lp.y, lp.z};
                n                 = 0 != (fl & 2u) ? V3{-lg.x, -lg.y, -lg.z} : V3{lg.x, lg.y, lg.z};
                total_sphere      = 0 != (fl & 1u);
                threshold         = 0 != (fl & 4u) ? kLowThreshold : view.split_threshold;
                const float random = lp.w;
                num_picks          = 0;
                end                = 0;

                // Tree.randomLight up to the descent, light_tree.zig:353-381
                float      ip    = 0.f;
                const bool split = threshold > 0.f;
                bool       done  = false;
                if (split && sc.lt_num_infinite < kMaxLightPicks - 1) {
                    for (uint32_t i = 0; i < sc.lt_num_infinite; ++i) {
                        st.picks[size_t(slot) * kMaxLightPicks + num_picks++] = make_uint2(__ldg(sc.lt_mapping + i), __float_as_uint(1.f));
                    }
                } else {
                    ip = sc.lt_infinite_weight;
                    if (random < sc.lt_infinite_guard) {
                        const uint32_t l  = dist1dSample(sc.lt_infinite_cdf, sc.lt_num_infinite + 1, random);
                        const float    lp = __ldg(sc.lt_infinite_cdf + l + 1) - __ldg(sc.lt_infinite_cdf + l);
                        st.picks[size_t(slot) * kMaxLightPicks + num_picks++] = make_uint2(__ldg(sc.lt_mapping + l), __float_as_uint(lp * ip));
                        done = true;
                    }
                }
                if (done || 0 == sc.lt_num_nodes) {
                    st.pick_n[slot] = num_picks;
                } else {
                    const float pd = 1.f - ip;
                    t              = {pd, __fdiv_rn(random - ip, pd), 0, split ? 0 : max_split_depth};
                    stack[end++]   = t;
                    active         = true;
                }
            }
        }
        if (0 == __ballot_sync(kFull, active)) {
            if (exhausted) break;
            continue;
        }
        if (active) {  // one iteration of the descent loop, light_tree.zig:396-444
            const LightNodeD node = loadLightNode(tr, t.node);
            if (0 != (node.meta & 1u)) {
                const bool     do_split = t.depth < max_split_depth && lightNodeSplit(node, p, threshold);
                const uint32_t c0       = node.meta >> 2;
                const uint32_t c1       = c0 + 1;
                if (do_split) {
                    t.depth += 1;
                    t.node       = c0;
                    stack[end++] = {t.pdf, t.random, c1, t.depth};
                } else {
                    t.depth = max_split_depth;

                    float p0 = lightNodeWeight(loadLightNode(tr, c0), p, n, total_sphere);
                    float p1 = lightNodeWeight(loadLightNode(tr, c1), p, n, total_sphere);

                    const float pt = p0 + p1;
                    if (0.f == pt) {
                        t = stack[--end];
                    } else {
                        p0 = __fdiv_rn(p0, pt);
                        p1 = __fdiv_rn(p1, pt);
                        if (t.random < p0) {
                            t.node = c0;
                            t.pdf *= p0;
                            t.random = __fdiv_rn(t.random, p0);
                        } else {
                            t.node = c1;
                            t.pdf *= p1;
                            t.random = zmin(__fdiv_rn(t.random - p0, p1), 1.f);
                        }
                    }
                }
            } else {
                const LightPickD pick = lightNodeRandomLight(sc, tr, node, p, n, total_sphere, t.random);
                if (pick.pdf > 0.f && num_picks < kMaxLightPicks) {
                    st.picks[size_t(slot) * kMaxLightPicks + num_picks++] = make_uint2(pick.offset, __float_as_uint(pick.pdf * t.pdf));
                }
                t = stack[--end];
            }
            if (0 == end) {  // `while (!stack.empty())`: the entry pushed first is popped last
                st.pick_n[slot] = num_picks;
                active          = false;
            }
        }
    }
}

// The kernel is long and branchy and its warps mostly wait for instruction fetch (ncu: no_instruction is the top stall, issue
// slots 13 % busy at 122 registers / 4 blocks per SM): more resident warps hide that better than registers help, so it is
// compiled for 8 blocks per SM (64 registers; measured 310 -> 265 ms per 4-spp pass of config 4, 12 and 16 blocks are slower).
#ifndef ZYGPU_LIGHT_BLOCKS
#define ZYGPU_LIGHT_BLOCKS 8
#endif
// Scenes with infinite lights run it twice: their picks come first in a vertex's list (Tree.randomLight, light_tree.zig:353-371), so
// phase 1 takes the Distant / Canopy picks of every vertex and phase 2 the finite ones, each on the sampler state the other left.
// A warp then works on one class of light at a time (and each instance holds half the code): lanes that fetch a new vertex no longer
// start with a sky sample while their neighbours are inside a mesh light. Phase 0 = every pick in one launch.
template <bool MeshLights, bool Infinite, int Phase>
__global__ void __launch_bounds__(128, ZYGPU_LIGHT_BLOCKS) lightSamplePersistent(SceneDevice sc, ZygpuView view, PathState st, PassParams pass,
                                                             uint32_t* __restrict__ work_counter) {
    constexpr bool kInfinitePicks = Infinite && 2 != Phase;  // this instance handles Distant / Canopy picks
    constexpr bool kFinitePicks   = 1 != Phase;
    __shared__ __align__(16) uint32_t sobol_tables[kSobolTableWords];
    loadSobolTables(sobol_tables);

    constexpr uint32_t kFull   = 0xffffffffu;
    const uint32_t     lane    = threadIdx.x & 31u;
    const uint32_t     n_items = st.counters[11];

    bool     active = false, exhausted = false;
    uint32_t slot = 0, pick_i = 0, pick_count = 0, num_records = 0;
    V3       p = {0.f, 0.f, 0.f}, geo_n = {0.f, 0.f, 0.f}, n = {0.f, 0.f, 0.f};
    bool     translucent = false;
    float    threshold   = 0.f;
    uint32_t pool_word   = 0;
    SamplerD sampler;
    sampler.sobol.tables = sobol_tables;

    for (;;) {
        const uint32_t idle = __ballot_sync(kFull, !active);
        if (0 != idle && !exhausted) {
            uint32_t base = 0;
            if (lane == uint32_t(__ffs(int(idle))) - 1u) base = atomicAdd(work_counter, uint32_t(__popc(idle)));
            base = __shfl_sync(kFull, base, __ffs(int(idle)) - 1);
            if (base + uint32_t(__popc(idle)) >= n_items) exhausted = true;
            const uint32_t index = base + uint32_t(__popc(idle & ((1u << lane) - 1u)));
            if (!active && index < n_items) {
                slot              = st.queue_l[index];
                const float4 lp   = st.ls_p[slot];
                const float4 lg   = st.ls_g[slot];
                const uint32_t fl = __float_as_uint(lg.w);
                p                 = {lp.x, lp.y, lp.z};
                geo_n             = {lg.x, lg.y, lg.z};
                n                 = 0 != (fl & 2u) ? neg3(geo_n) : geo_n;
                translucent       = 0 != (fl & 1u);
                threshold         = 0 != (fl & 4u) ? kLowThreshold : view.split_threshold;
                const uint32_t pn = st.pick_n[slot];
                pick_i            = 2 == Phase ? pn >> 16 : 0u;  // phase 1 leaves the index of the first finite pick there
                pick_count        = pn & 0xffffu;
                num_records       = 2 == Phase ? st.sh_n[slot] : 0u;
                const uint4 smp   = st.smp[slot];
                pool_word         = smp.w;
                loadSampler(st, slot, smp, pass, view.spp_total, (fl >> 8) & 0xffu, sampler);
                if (ZYG_SAMPLER_SOBOL != view.sampler) sampler.use_sobol = false;
                active = true;
            }
        }
        if (0 == __ballot_sync(kFull, active)) {
            if (exhausted) break;
            continue;
        }
        if (active && pick_i < pick_count) {  // Light.sampleTo for one pick, pathtracer_mis.zig:214-250
            const uint2      pk = st.picks[size_t(slot) * kMaxLightPicks + pick_i];
            const LightPickD pick{pk.x, __uint_as_float(pk.y)};
            pick_i += 1;

            const ZygpuLight light = sc.lights[pick.offset];
            const TrafoD     trafo = loadTrafo(sc.trafos, light.prop);
            const uint32_t   shape = sc.props[light.prop].shape;
            if (1 == Phase && ZYG_SHAPE_DISTANT != shape && ZYG_SHAPE_CANOPY != shape) {
                // the first finite pick ends phase 1: phase 2 resumes here
                pick_i -= 1;
                st.pick_n[slot] = pick_count | (pick_i << 16);
                st.sh_n[slot]   = num_records;
                storeSampler(st, slot, sampler, pool_word);
                active = false;
            } else if (kInfinitePicks && ZYG_SHAPE_DISTANT == shape) {  // Distant.sampleTo, distant.zig:78-107
                const float radius = trafo.scale.x;
                if (radius > 0.f) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);
                    float lx, ly;
                    diskConcentric(u0, u1, lx, ly);
                    const V3 ws  = scale3(radius, trafo.transformVector({lx, ly, 0.f}));
                    const V3 dir = normalize3(sub3(ws, trafo.r2));
                    if (!(dot3(dir, n) <= 0.f && !translucent)) {
                        if (num_records < st.shadow_stride) {
                            const size_t rec    = size_t(slot) * st.shadow_stride + num_records;
                            const V3     origin = offsetPoint(p, geo_n, dir);
                            const float  pdf    = __fdiv_rn(1.f, distantSolidAngle(radius));
                            st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                            st.sh_p[rec]  = make_float4(0.f, 0.f, 0.f, __uint_as_float(pick.offset | 0x80000000u));
                            st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                            num_records += 1;
                        } else {
                            st.counters[3] = 1;
                        }
                    }
                }
            } else if (kInfinitePicks && ZYG_SHAPE_CANOPY == shape) {  // Canopy.sampleMaterialTo, canopy.zig:94-131
                if (ZYG_LIGHT_PROP_IMAGE == light.light_class) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);
                    V3    dir;
                    float su, sv, pdf;
                    if (canopySampleMaterialTo(sc.image_samplers[light.sampler], trafo, n, translucent, u0, u1, dir, su, sv, pdf)) {
                        if (num_records < st.shadow_stride) {
                            const size_t rec    = size_t(slot) * st.shadow_stride + num_records;
                            const V3     origin = offsetPoint(p, geo_n, dir);
                            st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, pdf * pick.pdf);
                            st.sh_p[rec]  = make_float4(su, sv, 0.f, __uint_as_float(pick.offset | 0x80000000u));
                            st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                            num_records += 1;
                        } else {
                            st.counters[3] = 1;
                        }
                    }
                }
            } else if (kFinitePicks && MeshLights && ZYG_SHAPE_TRIANGLE_MESH == shape && ZYGPU_NULL != light.sampler) {
                FragD frag;  // meshLightSampleTo reads the shading point and offsets from it
                frag.p      = p;
                frag.geo_n  = geo_n;
                num_records = meshLightSampleTo(sc, st, slot, light, pick, trafo, frag, n, translucent, threshold, sampler, num_records);
            } else if (kFinitePicks && ZYG_SHAPE_SPHERE == shape) {  // Sphere.sampleTo, sphere.zig:323-393
                SphereLightD sl;
                sl.init(trafo, p);
                const uint32_t ns = sl.valid ? lightNumSamples(light, threshold) : 0;
                for (uint32_t k = 0; k < ns; ++k) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);
                    V3    lp, wn, dir;
                    float pdf;
                    if (!sl.sample(trafo, p, n, translucent, u0, u1, lp, wn, dir, pdf)) continue;
                    if (num_records < st.shadow_stride) {
                        const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                        const V3     origin    = offsetPoint(p, geo_n, dir);
                        const V3     light_pos = offsetRay(lp, wn);
                        st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, (float(ns) * pdf) * pick.pdf);
                        st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                        st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                        num_records += 1;
                    } else {
                        st.counters[3] = 1;
                    }
                }
            } else if (kFinitePicks && ZYG_SHAPE_RECTANGLE == shape && ZYG_LIGHT_PROP_IMAGE == light.light_class) {  // Rectangle.sampleMaterialTo
                const uint32_t            ns   = lightNumSamples(light, threshold);
                const ImageSamplerDevice& is   = sc.image_samplers[light.sampler];
                const float               area = trafo.scale.x * trafo.scale.y;
                for (uint32_t k = 0; k < ns; ++k) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);
                    float su, sv, rs_pdf;
                    imageSample(is, u0, u1, su, sv, rs_pdf);
                    if (0.f == rs_pdf) continue;
                    const V3 ws   = trafo.objectToWorldPoint({-1.f * su + 0.5f, -1.f * sv + 0.5f, 0.f});
                    const V3 axis = sub3(ws, p);
                    V3       wn   = trafo.r2;
                    if (0 != light.two_sided && dot3(wn, axis) > 0.f) wn = neg3(wn);
                    const float sl  = squaredLength3(axis);
                    const float t   = __fsqrt_rn(sl);
                    const V3    dir = divs3(axis, t);
                    const float c   = -dot3(wn, dir);
                    if (c < kDotMin || (dot3(dir, n) <= 0.f && !translucent)) continue;
                    if (num_records < st.shadow_stride) {
                        const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                        const V3     origin    = offsetPoint(p, geo_n, dir);
                        const V3     light_pos = offsetRay(ws, wn);
                        st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, __fdiv_rn(float(ns) * rs_pdf * sl, c * area) * pick.pdf);
                        st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                        st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                        if (nullptr != st.sh_uv) st.sh_uv[rec] = make_float2(su, sv);
                        num_records += 1;
                    } else {
                        st.counters[3] = 1;
                    }
                }
            } else if (kFinitePicks && ZYG_SHAPE_RECTANGLE == shape) {  // Rectangle.sampleTo, rectangle.zig:305-357
                const uint32_t ns  = lightNumSamples(light, threshold);
                const float    nsf = float(ns);
                SphQuadD       squad;
                squad.init(trafo.scale, trafo.worldToFramePoint(p));
                const float sample_pdf = nsf * squad.pdf(trafo.scale);
                for (uint32_t k = 0; k < ns; ++k) {
                    float u0, u1;
                    sampler.sample2D(u0, u1);

                    const V3 ls  = squad.sample(u0, u1);
                    const V3 ws  = trafo.frameToWorldPoint(ls);
                    const V3 dir = normalize3(sub3(ws, p));

                    V3 wn = trafo.r2;
                    if (0 != light.two_sided && dot3(wn, dir) > 0.f) wn = neg3(wn);

                    if (-dot3(wn, dir) < kDotMin || 0.f == squad.S || (dot3(dir, n) <= 0.f && !translucent)) continue;

                    if (num_records < st.shadow_stride) {
                        const size_t rec       = size_t(slot) * st.shadow_stride + num_records;
                        const V3     origin    = offsetPoint(p, geo_n, dir);
                        const V3     light_pos = offsetRay(ws, wn);
                        st.sh_o[rec]  = make_float4(origin.x, origin.y, origin.z, sample_pdf * pick.pdf);
                        st.sh_p[rec]  = make_float4(light_pos.x, light_pos.y, light_pos.z, __uint_as_float(pick.offset));
                        st.sh_wi[rec] = make_float4(dir.x, dir.y, dir.z, 0.f);
                        num_records += 1;
                    } else {
                        st.counters[3] = 1;
                    }
                }
            }
        }
        if (active && pick_i >= pick_count) {
            if (1 == Phase) st.pick_n[slot] = pick_count | (pick_count << 16);  // no finite pick: phase 2 only queues the records
            st.sh_n[slot] = num_records;
            storeSampler(st, slot, sampler, pool_word);
            if (1 != Phase && nullptr != st.queue_r && 0 != num_records) {
                const uint32_t base = atomicAdd(&st.counters[10], num_records);
                for (uint32_t k = 0; k < num_records; ++k) st.queue_r[base + k] = slot * st.shadow_stride + k;
            }
            active = false;
        }
    }
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

// The rest of PathtracerMIS.li: evaluateLight after the visibility test (pathtracer_mis.zig:252-277), the direct-light
// add (:116-117), mat_sample.sample and the next vertex (:121-166).
template <bool Split, bool Textured>
__global__ void __launch_bounds__(kBlock, Split ? 3 : ZYGPU_SHADE_BLOCKS) shadeBKernel(SceneDevice sc, ZygpuView view, PathState st, PassParams pass, uint32_t round) {
    __shared__ __align__(16) uint32_t sobol_tables[kSobolTableWords];
    const uint32_t count = st.counters[1];
    if (blockIdx.x * blockDim.x >= count) return;  // see shade_a
    loadSobolTables(sobol_tables);
    const LutsD    luts{sc.luts};
    const uint32_t iters = (count + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t i     = it * gridDim.x * blockDim.x + blockIdx.x * blockDim.x + threadIdx.x;
        bool           alive = false;  // the slot got its first vertex of the next generation
        uint32_t       slot  = 0;
        uint32_t       child_id[2] = {0, 0};
        bool           child_on[2] = {false, false};
        if (i < count) {
            slot                 = st.queue_b[i];
            const uint4    smp   = st.smp[slot];
            uint32_t       pool  = smp.w;
            const uint32_t lane  = Split ? poolCurLane(pool, round) : 0;
            const uint32_t vid   = lane * st.capacity + slot;
            LoadedVertex   lv     = loadVertex(st, vid);
            const VertexD& vertex = lv.v;

            const uint32_t total_depth = vertex.probe_depth;

            SamplerD sampler;
            sampler.sobol.tables = sobol_tables;
            loadSampler(st, slot, smp, pass, view.spp_total, total_depth, sampler);
            if (ZYG_SAMPLER_SOBOL != view.sampler) sampler.use_sobol = false;

            FragD frag;
            shapeFragment(sc, lv.prop, vertex.ray, lv.isec, frag);

            MediaD media{0, {0, 0, 0}, {0, 0, 0}};
            if (Split && 0 != lv.num_media) media = unpackMedia(st.med[vid], lv.num_media);

            const bool          caustics   = 0 == (vertex.state & kPrimaryRay) ? 0 != view.caustics_path : true;
            const V3      wo = neg3(vertex.ray.d);
            ZygpuMaterial m  = sc.materials[__ldg(sc.material_ids + sc.props[frag.prop].parts_start + frag.part)];
            float               ior_outside      = 1.f;
            int                 highest_priority = -128;
            if (Split) mediaForSample(sc, media, frag, wo, ior_outside, highest_priority);
            // the same texels shade_a's material sample read
            const MatSampleD mat_sample = texturedMaterialSample<Split, Textured>(sc, m, frag, wo, Textured ? st.stoch[vid] : 0.f, view.regularize_roughness,
                                                                                  lv.reg_alpha, caustics, view.specular_threshold, ior_outside,
                                                                                  highest_priority);

            const uint32_t max_splits = Split ? maxSplits(lv.path_count_log2, 0 != (vertex.state & kPrimaryRay), total_depth) : 1;

            // evaluateLight for the visible records
            V3             next_light = splat3(0.f);
            const uint32_t num        = st.sh_n[slot];
            for (uint32_t k = 0; k < num; ++k) {
                const size_t rec = size_t(slot) * st.shadow_stride + k;
                const float4 wi4 = st.sh_wi[rec];
                if (0.f == wi4.w) continue;
                const float4 o4 = st.sh_o[rec];
                const float4 p4 = st.sh_p[rec];
                const V3     wi = {wi4.x, wi4.y, wi4.z};

                // Light.evaluateTo, light.zig:119-132
                const float         stochastic_r = sampler.sample1D();
                const uint32_t      light_id = __float_as_uint(p4.w) & 0x7FFFFFFFu;
                const ZygpuLight    light    = sc.lights[light_id];
                const TrafoD        ltrafo   = loadTrafo(sc.trafos, light.prop);
                const ZygpuMaterial lm       = sc.materials[__ldg(sc.material_ids + sc.props[light.prop].parts_start + light.part)];
                const float area     = 0.f != lm.emission_normalize ? shapeArea(sc.props[light.prop].shape, ltrafo.scale) : 1.f;
                // an image-mapped (PROP_IMAGE) light left the uvw of its sample in the record's position lanes (infinite lights:
                // the lanes are free) or in sh_uv (finite lights)
                const float2 luv = 0 != (__float_as_uint(p4.w) & 0x80000000u) || nullptr == st.sh_uv ? make_float2(p4.x, p4.y) : st.sh_uv[rec];
                const V3    radiance = ZYGPU_NULL != lm.emission_map
                                           ? emittanceRadianceMapped(lm, wi, ltrafo, area, false,
                                                                     imageTexel(sc.image_samplers[lm.emission_map], luv.x, luv.y, stochastic_r))
                                           : emittanceRadiance(lm, wi, ltrafo, area, false);

                const BxdfResult bxdf_result = mat_sample.template evaluate<Split>(luts, wi, max_splits);

                const float light_pdf = o4.w;
                const float weight    = predividedPowerHeuristic(light_pdf, bxdf_result.pdf);

                next_light = add3(next_light, mul3(scale3(weight, radiance), bxdf_result.reflection));
            }

            const V3 split_throughput = scale3(lv.split_weight, lv.throughput);
            ivalueAdd(st, slot, mul3(split_throughput, next_light), total_depth, 1, false, false);

            BxdfSample     sample_results[Split ? 2 : 1];
            const uint32_t path_count = mat_sample.template sample<Split>(luts, sampler, max_splits, sample_results);

            if (Split) pool = poolFree(pool, lane);  // the parent lives in registers from here on
            if (slot == pass.debug_slot) {
                for (uint32_t c = 0; c < path_count; ++c) {
                    printf("[gpu]   sample %u/%u event %u sw %g pdf %g wi %.9g %.9g %.9g\n", c, path_count, sample_results[c].event,
                           sample_results[c].split_weight, sample_results[c].pdf, sample_results[c].wi.x, sample_results[c].wi.y, sample_results[c].wi.z);
                }
            }

            for (uint32_t c = 0; c < (Split ? path_count : min(path_count, 1u)); ++c) {
                const BxdfSample& sample_result = sample_results[c];

                // Vertex.State.update, vertex.zig:30-43
                uint32_t state = vertex.state;
                if (kScatterSpecular == sample_result.scattering) {
                    state |= kSpecular;
                    state = 0.f == sample_result.reg_alpha ? (state | kSingular) : (state & ~kSingular);
                    if (0 != (state & kPrimaryRay)) state |= kStartedSpecular;
                } else if (kEventStraight != sample_result.event) {
                    state &= ~(kSpecular | kSingular | kPrimaryRay);
                }

                uint32_t vertex_depth = lv.vertex_depth;
                float    bxdf_pdf     = vertex.bxdf_pdf;
                V3       origin       = vertex.origin;
                V3       geo_n        = vertex.geo_n;
                float    reg_alpha    = lv.reg_alpha;
                if (kEventStraight != sample_result.event) {
                    state        = mat_sample.translucent ? (state | kTranslucent) : (state & ~kTranslucent);
                    vertex_depth = vertex.probe_depth;
                    bxdf_pdf     = sample_result.pdf;
                    origin       = frag.p;
                    geo_n        = mat_sample.geo_n;
                    reg_alpha    = sample_result.reg_alpha;
                }

                const V3 throughput = mul3(lv.throughput, divs3(sample_result.reflection, sample_result.pdf));

                const V3 next_o = frag.offsetP(sample_result.wi);  // Fragment.offsetRay, intersection.zig:118-120

                if (!(kEventTransmission == sample_result.event || kEventStraight == sample_result.event)) state &= ~kTransparent;

                uint32_t cvid = vid;
                uint32_t pcl2 = 0, num_media = 0;
                if (Split) {
                    const uint32_t clane = poolAlloc(pool);
                    if (clane > 3u) break;  // cannot happen: path_count keeps the live vertices of a sample at 4 or fewer
                    alive = alive || 0 == poolNextCount(pool);
                    pool  = poolAppendNext(pool, clane);
                    cvid  = clane * st.capacity + slot;

                    pcl2 = lv.path_count_log2 + (path_count > 1 ? 1u : 0u);  // path_count *= number of samples (1 or 2)

                    MediaD next_media = media;
                    if (kEventTransmission == sample_result.event) {  // Vertex.interfaceChange, vertex.zig:95-110
                        if (frag.sameHemisphere(sample_result.wi)) {
                            mediaRemove(next_media, frag.prop, frag.part);
                        } else {
                            mediaPush(next_media, frag.prop, frag.part);
                        }
                    }
                    num_media    = next_media.count;
                    st.med[cvid] = packMedia(next_media);
                    child_id[c]  = cvid;
                    child_on[c]  = true;
                } else {
                    alive = true;
                }

                st.ray_o[cvid]  = make_float4(next_o.x, next_o.y, next_o.z,
                                              __uint_as_float(packFlags(state, vertex.probe_depth + 1, vertex_depth, pcl2, num_media)));
                st.ray_d[cvid]  = make_float4(sample_result.wi.x, sample_result.wi.y, sample_result.wi.z, kRayMaxT);
                st.thr[cvid]    = make_float4(throughput.x, throughput.y, throughput.z, bxdf_pdf);
                st.prev_p[cvid] = make_float4(origin.x, origin.y, origin.z, reg_alpha);
                st.prev_n[cvid] = make_float4(geo_n.x, geo_n.y, geo_n.z, lv.split_weight * sample_result.split_weight);
            }
            sampler.incrementPadding();
            storeSampler(st, slot, sampler, pool);
        }
        queuePush(st.queue_a, &st.counters[4], alive, slot);
        if (Split) {
            queuePush(st.queue_t, &st.counters[7], child_on[0], child_id[0]);
            queuePush(st.queue_t, &st.counters[7], child_on[1], child_id[1]);
        }
    }
}

// Counter bookkeeping between the stages (single thread; the queue lengths never leave the device).
__global__ void beginGenerationKernel(PathState st) {  // after extend: the trace queue is consumed, nothing is queued yet
    st.counters[1]  = 0;
    st.counters[4]  = 0;
    st.counters[7]  = 0;
    st.counters[9]  = 0;
    st.counters[10] = 0;
    st.counters[11] = 0;
}
__global__ void beginRoundKernel(PathState st) {
    st.counters[1]  = 0;
    st.counters[10] = 0;
    st.counters[11] = 0;
}
__global__ void advanceKernel(PathState st) {
    st.counters[0] = st.counters[4];
    st.counters[1] = 0;
    st.counters[4]  = 0;
    st.counters[9]  = 0;
    st.counters[10] = 0;
    st.counters[11] = 0;
}

// Sensor.addSample for every sample of the pass, gathered per film pixel (sensor.zig:168-385, buffer_opaque.zig:39-45).
__device__ __forceinline__ V3 clampColor(V3 color, float mx) {  // sensor.zig:615-624
    const float mc = hmax3(color);
    if (mc > mx) return scale3(__fdiv_rn(mx, mc), color);
    return color;
}

__device__ __forceinline__ float filterEval(const ZygpuView& view, float s) {  // sensor.zig:626-628
    const float    cx     = zmin(fabsf(s), view.filter_range_end);
    const float    o      = cx * view.filter_inverse_interval;
    const uint32_t offset = uint32_t(o);
    const float    t      = o - float(offset);
    return zlerp(view.filter[offset], view.filter[min(offset + 1, 29u)], t);
}

__global__ void __launch_bounds__(kBlock) filmKernel(ZygpuView view, PathState st, PassParams pass, float4* film) {
    const int32_t  w  = view.resolution[0];
    const int32_t  h  = view.resolution[1];
    const int32_t  fr = view.filter_radius_int;
    const uint32_t padded = pass.padded_w * pass.padded_h;

    for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < uint32_t(w * h); pixel += gridDim.x * blockDim.x) {
        const int32_t x = int32_t(pixel % uint32_t(w));
        const int32_t y = int32_t(pixel / uint32_t(w));
        // Sensor.add bounds test: pixels outside the crop receive nothing (sensor.zig:559-562)
        if (x < view.crop[0] || x >= view.crop[2] || y < view.crop[1] || y >= view.crop[3]) continue;

        float4 value = film[pixel];

        for (uint32_t s = 0; s < pass.samples_in_pass; ++s) {
            for (int32_t dy = -fr; dy <= fr; ++dy) {
                for (int32_t dx = -fr; dx <= fr; ++dx) {
                    // the sample's pixel q = (x + dx, y + dy) must have been rendered: crop extended by the filter radius
                    const int32_t qx = x + dx, qy = y + dy;
                    if (qx < view.crop[0] - fr || qx >= view.crop[2] + fr || qy < view.crop[1] - fr || qy >= view.crop[3] + fr) continue;
                    const uint32_t slot = s * padded + uint32_t(qy + fr) * pass.padded_w + uint32_t(qx + fr);

                    const float4 e  = st.acc_e[slot];
                    const float4 d  = st.acc_d[slot];
                    const float4 in = st.acc_i[slot];

                    const V3 emission = clampColor({e.x, e.y, e.z}, view.clamp_emission);
                    const V3 direct   = clampColor({d.x, d.y, d.z}, view.clamp_direct);
                    const V3 indirect = clampColor({in.x, in.y, in.z}, view.clamp_indirect);
                    const V3 composed = add3(add3(emission, direct), indirect);

                    float weight = 1.f;
                    if (fr > 0) {
                        const float ox = e.w - 0.5f;
                        const float oy = d.w - 0.5f;
                        weight         = filterEval(view, ox + float(dx)) * filterEval(view, oy + float(dy));
                    }
                    // Opaque.addPixel
                    value.x += weight * composed.x;
                    value.y += weight * composed.y;
                    value.z += weight * composed.z;
                    value.w += weight;
                }
            }
        }
        film[pixel] = value;
    }
}

// Opaque.resolveTonemap with the Linear tonemapper, buffer_opaque.zig:73-79, tonemapper.zig:36-39, aces.zig:19-27
__global__ void __launch_bounds__(kBlock) resolveKernel(ZygpuView view, const float4* film, float4* rgba, uint32_t num_pixels) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < num_pixels; i += gridDim.x * blockDim.x) {
        const float4 p = film[i];
        const V3     c = {fabsf(__fdiv_rn(p.x, p.w)), fabsf(__fdiv_rn(p.y, p.w)), fabsf(__fdiv_rn(p.z, p.w))};
        const V3     s = scale3(view.exposure_factor, c);
        const V3     srgb = add3(add3(scale3(s.x, {1.70505155f, -0.13025714f, -0.02400328f}), scale3(s.y, {-0.62179068f, 1.14080289f, -0.12896877f})),
                                 scale3(s.z, {-0.08325840f, -0.01054853f, 1.15297171f}));
        rgba[i] = make_float4(srgb.x, srgb.y, srgb.z, 1.f);
    }
}

int numSms() {
    static int sms = 0;
    if (0 == sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

// Blocks per SM of the shade launches (ZYGPU_SHADE_GRID overrides). Measured: exactly the resident 4 for the kernels without
// path splits (Cornell 39.2 -> 37.1 ms: one table prologue per block, no second wave), 16 for the split kernels of glass scenes,
// whose blocks finish unevenly (config 3: 389.8 ms with 4, 378.7 ms with 16).
uint32_t shadeGrid(bool split) {
    static const int v = [] {
        const char* e = getenv("ZYGPU_SHADE_GRID");
        return e ? std::max(1, atoi(e)) : 0;
    }();
    return 0 != v ? uint32_t(v) : (split ? 16u : 4u);
}

// Blocks per SM of the grid-stride walk kernels (top / extend / shadow; ZYGPU_WALK_GRID overrides). They have no per-block
// prologue, so many short blocks even out the uneven walks: 64 measured 1 - 2 % faster than 16 on configs 1, 3 and 4.
uint32_t walkGrid() {
    static const uint32_t v = [] {
        const char* e = getenv("ZYGPU_WALK_GRID");
        return e ? uint32_t(std::max(1, atoi(e))) : 64u;
    }();
    return v;
}

// Grid-stride launches: a multiple of the SM count, never more blocks than there is work.
uint32_t gridFor(uint32_t items, uint32_t blocks_per_sm) {
    const uint32_t needed = (items + kBlock - 1) / kBlock;
    return std::max(1u, std::min(needed, uint32_t(numSms()) * blocks_per_sm));
}

}  // namespace

cudaError_t uploadSobolDirections() {
    // Joe & Kuo direction numbers (new-joe-kuo-6.21201) for dimensions 1-5: degree s, coefficients a, initial m_i.
    static const uint32_t S[5]    = {0, 1, 2, 3, 3};
    static const uint32_t A[5]    = {0, 0, 1, 1, 2};
    static const uint32_t M[5][3] = {{0, 0, 0}, {1, 0, 0}, {1, 3, 0}, {1, 3, 1}, {1, 1, 1}};
    uint32_t              d[5][32];
    for (uint32_t i = 0; i < 32; ++i) d[0][i] = 1u << (31 - i);
    for (uint32_t j = 1; j < 5; ++j) {
        const uint32_t s = S[j];
        for (uint32_t i = 0; i < 32; ++i) {
            if (i < s) {
                d[j][i] = M[j][i] << (31 - i);
            } else {
                uint32_t v = d[j][i - s] ^ (d[j][i - s] >> s);
                for (uint32_t k = 1; k < s; ++k) v ^= ((A[j] >> (s - 1 - k)) & 1u) * d[j][i - k];
                d[j][i] = v;
            }
        }
    }
    static uint32_t tables[kSobolTableWords];
    for (uint32_t byte = 0; byte < 4; ++byte) {
        for (uint32_t dim = 0; dim < 5; ++dim) {
            for (uint32_t value = 0; value < 256; ++value) {
                uint32_t x = 0;
                for (uint32_t j = 0; j < 8; ++j) {
                    if (0 != ((value >> j) & 1u)) x ^= d[dim][8 * byte + j];
                }
                tables[(byte * 5 + dim) * 256 + value] = x;
            }
        }
    }
    return cudaMemcpyToSymbol(d_sobol_tables, tables, sizeof(tables));
}

cudaError_t launchGenerate(const ZygpuView& view, const PathState& st, const PassParams& pass, cudaStream_t stream) {
    generateKernel<<<gridFor(pass.num_paths, 16), kBlock, 0, stream>>>(view, st, pass);
    return cudaGetLastError();
}
namespace {

int envInt(const char* name, int fallback) {
    const char* v = getenv(name);
    return v ? atoi(v) : fallback;
}

struct SceneTraceConfig {
    int              variant;  // 0: one thread per ray (extendKernel / shadowKernel), 1: top kernel + persistent mesh kernel,
                               // 2: fused two-level persistent kernel for scenes with meshes (default)
    SceneTraceTuning tune;
    SceneStepTuning  step;
    int              blocks_per_sm;
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
        return c;
    }();
    return cfg;
}

template <bool AnyHit>
cudaError_t launchSceneTrace(const SceneDevice& scene, const PathState& st, uint32_t max_items, bool has_meshes, cudaStream_t stream) {
    const SceneTraceConfig& cfg = sceneTraceConfig();
    // a prop tree that is a single leaf (a mesh and a few analytic props) gains nothing from the fused walk: the thread-per-ray
    // top kernel deals with it at full lane occupancy (measured on the 1M-triangle sphere scene: 48.2 ms against 51.4 ms fused)
    // (an instrumented pass always takes the fused kernel, the one that counts its fetches; the results are the same)
    if ((2 == cfg.variant && has_meshes && scene.num_solid_nodes > 1) || nullptr != st.tally) {
        static int resident = 0, resident_counted = 0;
        if (0 == resident) {
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sceneTracePersistent<AnyHit, false>, 128, 0);
            if (cfg.blocks_per_sm > 0) per_sm = std::min(per_sm, cfg.blocks_per_sm);
            resident = std::max(per_sm, 1) * numSms();
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sceneTracePersistent<AnyHit, true>, 128, 0);
            resident_counted = std::max(per_sm, 1) * numSms();
        }
        const bool     counted = nullptr != st.tally;
        const uint32_t needed  = (max_items + 127) / 128;
        const uint32_t grid    = std::max(1u, std::min<uint32_t>(uint32_t(counted ? resident_counted : resident), needed));
        cudaError_t    err     = cudaMemsetAsync(st.counters + 8, 0, sizeof(uint32_t), stream);
        if (cudaSuccess != err) return err;
        if (counted) {
            sceneTracePersistent<AnyHit, true><<<grid, 128, 0, stream>>>(scene, st, st.counters + 8, cfg.step, st.tally);
        } else {
            sceneTracePersistent<AnyHit, false><<<grid, 128, 0, stream>>>(scene, st, st.counters + 8, cfg.step, nullptr);
        }
        return cudaGetLastError();
    }
    // counters[2] = mesh queue length, counters[8] = work counter of the persistent kernel
    cudaError_t err = cudaMemsetAsync(st.counters + 2, 0, sizeof(uint32_t), stream);
    if (cudaSuccess != err) return err;
    topKernel<AnyHit><<<gridFor(max_items, walkGrid()), kBlock, 0, stream>>>(scene, st);
    err = cudaGetLastError();
    if (cudaSuccess != err || !has_meshes) return err;

    static int resident = 0;
    if (0 == resident) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, meshTracePersistent<AnyHit>, 128, 0);
        if (cfg.blocks_per_sm > 0) per_sm = std::min(per_sm, cfg.blocks_per_sm);
        resident = std::max(per_sm, 1) * numSms();
    }
    const uint32_t needed = (max_items + 127) / 128;
    const uint32_t grid   = std::max(1u, std::min<uint32_t>(uint32_t(resident), needed));
    err                   = cudaMemsetAsync(st.counters + 8, 0, sizeof(uint32_t), stream);
    if (cudaSuccess != err) return err;
    meshTracePersistent<AnyHit><<<grid, 128, 0, stream>>>(scene, st, st.counters + 8, cfg.tune);
    return cudaGetLastError();
}

}  // namespace

uint32_t sceneTraceLaunches(bool has_meshes, uint32_t num_solid_nodes) {  // kernels per extend / shadow stage
    const int v = sceneTraceConfig().variant;
    if (0 == v || !has_meshes) return 1u;
    return (2 == v && num_solid_nodes > 1) ? 1u : 2u;  // fused kernel, or top kernel + mesh kernel
}

cudaError_t launchExtend(const SceneDevice& scene, const PathState& st, uint32_t max_items, bool has_meshes, cudaStream_t stream) {
    if (0 != sceneTraceConfig().variant) return launchSceneTrace<false>(scene, st, max_items, has_meshes, stream);
    extendKernel<<<gridFor(max_items, walkGrid()), kBlock, 0, stream>>>(scene, st);
    return cudaGetLastError();
}
cudaError_t launchShadeA(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass, uint32_t max_items,
                         uint32_t round, cudaStream_t stream) {
    // triangle-mesh lights are always sampled by the light kernels (capi/zygpu_render.cu): no instance samples them inline
    const uint32_t features = (st.lanes > 1 ? kFeatureSplit : 0u) | (scene.num_mesh_samplers > 0 ? kFeatureMeshLights : 0u) |
                              (scene.num_infinite_props > 0 ? kFeatureInfiniteLights : 0u) | (nullptr != st.queue_l ? kFeatureDeferredLights : 0u) |
                              (nullptr != st.stoch ? kFeatureTextured : 0u);
    if (0 != (features & kFeatureMeshLights) && 0 == (features & kFeatureDeferredLights)) return cudaErrorInvalidValue;
    const uint32_t grid = gridFor(max_items, shadeGrid(st.lanes > 1));
    if (st.lanes <= 1) round = 0;
#define ZYGPU_SHADE_A(F) \
    case F: shadeAKernel<F><<<grid, kBlock, 0, stream>>>(scene, view, st, pass, round); break;
    switch (features) {
        ZYGPU_SHADE_A(0) ZYGPU_SHADE_A(1) ZYGPU_SHADE_A(4) ZYGPU_SHADE_A(5) ZYGPU_SHADE_A(8) ZYGPU_SHADE_A(9) ZYGPU_SHADE_A(10) ZYGPU_SHADE_A(11)
        ZYGPU_SHADE_A(12) ZYGPU_SHADE_A(13) ZYGPU_SHADE_A(14) ZYGPU_SHADE_A(15)
        ZYGPU_SHADE_A(16) ZYGPU_SHADE_A(17) ZYGPU_SHADE_A(20) ZYGPU_SHADE_A(21) ZYGPU_SHADE_A(24) ZYGPU_SHADE_A(25) ZYGPU_SHADE_A(26) ZYGPU_SHADE_A(27)
        ZYGPU_SHADE_A(28) ZYGPU_SHADE_A(29) ZYGPU_SHADE_A(30) ZYGPU_SHADE_A(31)
        default: return cudaErrorInvalidValue;
    }
#undef ZYGPU_SHADE_A
    return cudaGetLastError();
}
cudaError_t launchLightStages(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass, uint32_t max_items,
                              cudaStream_t stream) {
    if (nullptr == st.queue_l) return cudaSuccess;
    cudaError_t err = cudaMemsetAsync(st.counters + 12, 0, 3 * sizeof(uint32_t), stream);
    if (cudaSuccess != err) return err;

    static int resident_select = 0, resident_sample = 0;
    if (0 == resident_select) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lightSelectPersistent, 128, 0);
        resident_select = std::max(per_sm, 1) * numSms();
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lightSamplePersistent<true, true, 0>, 128, 0);
        resident_sample = std::max(per_sm, 1) * numSms();
    }
    const uint32_t needed = (max_items + 127) / 128;
    lightSelectPersistent<<<std::max(1u, std::min<uint32_t>(uint32_t(resident_select), needed)), 128, 0, stream>>>(scene, view, st, st.counters + 12);

    const uint32_t grid = std::max(1u, std::min<uint32_t>(uint32_t(resident_sample), needed));
    const bool     ml = scene.num_mesh_samplers > 0, inf = scene.num_infinite_props > 0;
    static const bool phases = nullptr == getenv("ZYGPU_LIGHT_PHASES") || 0 != atoi(getenv("ZYGPU_LIGHT_PHASES"));
    if (ml && inf && phases) {
        lightSamplePersistent<true, true, 1><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
        lightSamplePersistent<true, true, 2><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 14);
    } else if (ml && inf) {
        lightSamplePersistent<true, true, 0><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
    } else if (ml) {
        lightSamplePersistent<true, false, 0><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
    } else if (inf && phases) {
        lightSamplePersistent<false, true, 1><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
        lightSamplePersistent<false, true, 2><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 14);
    } else if (inf) {
        lightSamplePersistent<false, true, 0><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
    } else {
        lightSamplePersistent<false, false, 0><<<grid, 128, 0, stream>>>(scene, view, st, pass, st.counters + 13);
    }
    return cudaGetLastError();
}
cudaError_t launchBeginGeneration(const PathState& st, cudaStream_t stream) {
    beginGenerationKernel<<<1, 1, 0, stream>>>(st);
    return cudaGetLastError();
}
cudaError_t launchBeginRound(const PathState& st, cudaStream_t stream) {
    beginRoundKernel<<<1, 1, 0, stream>>>(st);
    return cudaGetLastError();
}
cudaError_t launchEndGeneration(const PathState& st, cudaStream_t stream) {
    advanceKernel<<<1, 1, 0, stream>>>(st);
    return cudaGetLastError();
}
cudaError_t launchShadow(const SceneDevice& scene, const PathState& st, uint32_t max_items, bool has_meshes, cudaStream_t stream) {
    if (0 != sceneTraceConfig().variant) return launchSceneTrace<true>(scene, st, max_items * st.shadow_stride, has_meshes, stream);
    shadowKernel<<<gridFor(max_items, walkGrid()), kBlock, 0, stream>>>(scene, st);
    return cudaGetLastError();
}
cudaError_t launchShadeB(const SceneDevice& scene, const ZygpuView& view, const PathState& st, const PassParams& pass, uint32_t max_items,
                         uint32_t round, cudaStream_t stream) {
    if (st.lanes > 1) {
        if (nullptr != st.stoch) {
            shadeBKernel<true, true><<<gridFor(max_items, shadeGrid(true)), kBlock, 0, stream>>>(scene, view, st, pass, round);
        } else {
            shadeBKernel<true, false><<<gridFor(max_items, shadeGrid(true)), kBlock, 0, stream>>>(scene, view, st, pass, round);
        }
    } else {
        if (nullptr != st.stoch) {
            shadeBKernel<false, true><<<gridFor(max_items, shadeGrid(false)), kBlock, 0, stream>>>(scene, view, st, pass, 0);
        } else {
            shadeBKernel<false, false><<<gridFor(max_items, shadeGrid(false)), kBlock, 0, stream>>>(scene, view, st, pass, 0);
        }
        advanceKernel<<<1, 1, 0, stream>>>(st);
    }
    return cudaGetLastError();
}
cudaError_t launchFilm(const ZygpuView& view, const PathState& st, const PassParams& pass, float4* film, cudaStream_t stream) {
    filmKernel<<<gridFor(uint32_t(view.resolution[0] * view.resolution[1]), 16), kBlock, 0, stream>>>(view, st, pass, film);
    return cudaGetLastError();
}
cudaError_t launchResolve(const ZygpuView& view, const float4* film, float4* rgba, uint32_t num_pixels, cudaStream_t stream) {
    resolveKernel<<<gridFor(num_pixels, 16), kBlock, 0, stream>>>(view, film, rgba, num_pixels);
    return cudaGetLastError();
}

}  // namespace zygpu
