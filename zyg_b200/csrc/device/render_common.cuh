#pragma once
// Shared by render.cu (shading stages) and render_trace.cu (traversal stages): path-state packing, the prop tests of the
// thread-per-ray walks, ray loading, grid sizing.
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
        case ZYG_SHAPE_DISK: return kPi * ((0.5f * scale.x) * (0.5f * scale.x));
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
        case ZYG_SHAPE_DISK: return diskIntersect(ray, trafo, isec);
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
        case ZYG_SHAPE_DISK: {
            HitD unused;
            return !diskIntersect(ray, trafo, unused);
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


inline int envInt(const char* name, int fallback) {
    const char* v = getenv(name);
    return v ? atoi(v) : fallback;
}

}  // namespace

}  // namespace zygpu
