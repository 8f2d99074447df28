// Device light-tree build, see light_build.cuh.
#include "light_build.cuh"

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <cub/cub.cuh>
#include <vector>

namespace zygpu {

namespace {

#include "lbvh.cuh"

constexpr float kPi = 3.14159265358979323846f;

__global__ void lightBoundsKernel(LightBuildInput in, Bounds* bounds) {
    const uint32_t l  = blockIdx.x * blockDim.x + threadIdx.x;
    float          v[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (l < in.num_lights) {
        const float4 lo = in.aabb_min[l], hi = in.aabb_max[l];
        v[0] = lo.x, v[1] = lo.y, v[2] = lo.z, v[3] = hi.x, v[4] = hi.y, v[5] = hi.z;
    }
    using Reduce = cub::BlockReduce<float, kThreads>;
    __shared__ typename Reduce::TempStorage tmp;
    for (int k = 0; k < 6; ++k) {
        const float r = k < 3 ? Reduce(tmp).Reduce(v[k], cub::Min()) : Reduce(tmp).Reduce(v[k], cub::Max());
        __syncthreads();
        if (0 == threadIdx.x) {
            if (k < 3) {
                atomicMin(&bounds->lo[k], floatToOrdered(r));
            } else {
                atomicMax(&bounds->hi[k - 3], floatToOrdered(r));
            }
        }
    }
}

__global__ void lightMortonKernel(LightBuildInput in, const Bounds* bounds, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= in.num_lights) return;
    const float4 lo = in.aabb_min[l], hi = in.aabb_max[l];
    const float  c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    uint32_t     q[3];
    for (int a = 0; a < 3; ++a) {
        const float mn = orderedToFloat(bounds->lo[a]), mx = orderedToFloat(bounds->hi[a]);
        const float e  = mx - mn;
        const float f  = e > 0.f ? (c[a] - mn) / e : 0.f;
        q[a]           = uint32_t(fminf(fmaxf(f * 1048576.f, 0.f), 1048575.f));
    }
    keys[l] = (spread20(q[0]) << 2) | (spread20(q[1]) << 1) | spread20(q[2]);
    vals[l] = l;
}

// BuildNode of the reference (light_tree_builder.zig:27-58) plus what the bottom-up merge needs
struct NodeStats {
    float4 lo, hi;    // bounds
    float4 cone;      // scene tree: merged cone, (1, 1, 1, 1) = no powered light below; primitive tree: sum of power * normal (w unused)
    float  power;
    float  sum_p2;    // sum of squared powers of the powered lights
    uint32_t num_powered;
    uint32_t two_sided;
};

struct StatArrays {
    float4* lo;
    float4* hi;
    float4* cone;
    float4* misc;  // power, sum_p2, num_powered (bits), two_sided (bits)
};

__device__ __forceinline__ void storeStats(const StatArrays& a, uint32_t i, const NodeStats& s) {
    a.lo[i]   = s.lo;
    a.hi[i]   = s.hi;
    a.cone[i] = s.cone;
    a.misc[i] = make_float4(s.power, s.sum_p2, __uint_as_float(s.num_powered), __uint_as_float(s.two_sided));
}

__device__ __forceinline__ NodeStats loadStats(const StatArrays& a, uint32_t i) {
    NodeStats    s;
    const float4 m = __ldcg(a.misc + i);
    s.lo           = __ldcg(a.lo + i);
    s.hi           = __ldcg(a.hi + i);
    s.cone         = __ldcg(a.cone + i);
    s.power        = m.x;
    s.sum_p2       = m.y;
    s.num_powered  = __float_as_uint(m.z);
    s.two_sided    = __float_as_uint(m.w);
    return s;
}

__device__ __forceinline__ NodeStats leafStats(const LightBuildInput& in, uint32_t l) {
    NodeStats   s;
    const float p = in.powers[l];
    s.lo          = in.aabb_min[l];
    s.hi          = in.aabb_max[l];
    s.lo.w = s.hi.w = 0.f;
    s.power       = p;
    s.sum_p2      = p * p;
    s.num_powered = p > 0.f ? 1u : 0u;
    const bool ts = in.two_sided ? 0 != in.two_sided[l] : in.all_two_sided;
    s.two_sided   = (p > 0.f && ts) ? 1u : 0u;
    const float4 c = in.cones[l];
    if (in.primitive) {
        s.cone = make_float4(p * c.x, p * c.y, p * c.z, 0.f);
    } else {
        s.cone = p > 0.f ? c : make_float4(1.f, 1.f, 1.f, 1.f);
    }
    return s;
}

__device__ __forceinline__ float clampUnit(float x) { return fminf(fmaxf(x, -1.f), 1.f); }

// math.cone.merge, src/base/math/cone.zig:8-44
__device__ float4 coneMergeD(float4 a, float4 b) {
    const bool a_empty = 1.f == a.x && 1.f == a.y && 1.f == a.z && 1.f == a.w;
    const bool b_empty = 1.f == b.x && 1.f == b.y && 1.f == b.z && 1.f == b.w;
    if (a_empty) return b;
    if (b_empty) return a;
    if (a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w) return a;

    float a_angle = acosf(clampUnit(a.w));
    float b_angle = acosf(clampUnit(b.w));
    if (b_angle > a_angle) {
        const float4 t = a;
        a              = b;
        b              = t;
        const float ta = a_angle;
        a_angle        = b_angle;
        b_angle        = ta;
    }
    const float d_angle = acosf(clampUnit(a.x * b.x + a.y * b.y + a.z * b.z));
    if (fminf(d_angle + b_angle, kPi) <= a_angle) return a;

    const float o_angle = (a_angle + d_angle + b_angle) / 2.f;
    if (o_angle >= kPi) return make_float4(a.x, a.y, a.z, -1.f);

    const float r_angle = o_angle - a_angle;
    float3      v       = make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
    const float vl      = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (!(vl > 0.f)) return make_float4(a.x, a.y, a.z, cosf(o_angle));  // opposite axes: any rotation axis would do
    v = make_float3(v.x / vl, v.y / vl, v.z / vl);

    // Mat3x3.initRotation(v, r_angle).transformVector(a), matrix3x3.zig:50-77, 113-127
    const float c = cosf(r_angle), s = sinf(r_angle), t = 1.f - c;
    const float at0 = v.x * v.y * t, at1 = v.z * s;
    const float bt0 = v.x * v.z * t, bt1 = v.y * s;
    const float ct0 = v.y * v.z * t, ct1 = v.x * s;
    const float3 r0 = make_float3(c + v.x * v.y * t, at0 - at1, bt0 + bt1);
    const float3 r1 = make_float3(at0 + at1, c + v.y * v.y * t, ct0 - ct1);
    const float3 r2 = make_float3(bt0 - bt1, ct0 + ct1, c + v.z * v.z * t);
    float3       r  = make_float3(a.x * r0.x + a.y * r1.x + a.z * r2.x, a.x * r0.y + a.y * r1.y + a.z * r2.y, a.x * r0.z + a.y * r1.z + a.z * r2.z);
    const float  rl = sqrtf(r.x * r.x + r.y * r.y + r.z * r.z);
    return make_float4(r.x / rl, r.y / rl, r.z / rl, cosf(o_angle));
}

__device__ __forceinline__ NodeStats mergeStats(const NodeStats& a, const NodeStats& b, bool primitive) {
    NodeStats s;
    s.lo          = make_float4(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z), 0.f);
    s.hi          = make_float4(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z), 0.f);
    s.power       = a.power + b.power;
    s.sum_p2      = a.sum_p2 + b.sum_p2;
    s.num_powered = a.num_powered + b.num_powered;
    s.two_sided   = a.two_sided | b.two_sided;
    s.cone        = primitive ? make_float4(a.cone.x + b.cone.x, a.cone.y + b.cone.y, a.cone.z + b.cone.z, 0.f) : coneMergeD(a.cone, b.cone);
    return s;
}

__device__ __forceinline__ NodeStats refStats(const LightBuildInput& in, const StatArrays& stats, const uint32_t* __restrict__ vals, uint32_t ref) {
    return 0 != (ref & kLeafBit) ? leafStats(in, vals[ref & ~kLeafBit]) : loadStats(stats, ref);
}

__global__ void aggregateKernel(LightBuildInput in, Hierarchy h, StatArrays stats, const uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= int(in.num_lights)) return;
    int32_t cur = h.leaf_parent[i] & ~kRightChild;
    while (cur >= 0) {
        if (0 == atomicAdd(h.flags + cur, 1u)) return;
        const NodeStats a = refStats(in, stats, vals, h.left[cur]);
        const NodeStats b = refStats(in, stats, vals, h.right[cur]);
        storeStats(stats, uint32_t(cur), mergeStats(a, b, in.primitive));
        __threadfence();
        const int32_t p = h.parent[cur];
        cur             = p < 0 ? -1 : (p & ~kRightChild);
    }
}

__device__ __forceinline__ float3 dominantAxis(float4 weighted_sum, float power) {
    const float3 d = make_float3(weighted_sum.x / power, weighted_sum.y / power, weighted_sum.z / power);
    const float  l = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    if (!(l > 0.f)) return make_float3(0.f, 0.f, 0.f);
    return make_float3(d.x / l, d.y / l, d.z / l);
}

// evaluateSampler, light_tree_builder.zig:193-263: the cone of a primitive-tree node is its power-weighted mean normal and the largest
// deviation of a powered triangle below it. Every triangle reports to all of its ancestors.
__global__ void coneAngleKernel(LightBuildInput in, Hierarchy h, StatArrays stats, const uint32_t* __restrict__ vals, uint32_t* __restrict__ angles) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= int(in.num_lights)) return;
    const uint32_t l = vals[i];
    if (!(in.powers[l] > 0.f)) return;
    const float4 n   = in.cones[l];
    int32_t      cur = h.leaf_parent[i] & ~kRightChild;
    while (cur >= 0) {
        const float3 axis = dominantAxis(stats.cone[cur], stats.misc[cur].x);
        const float  ang  = acosf(clampUnit(axis.x * n.x + axis.y * n.y + axis.z * n.z));
        atomicMax(angles + cur, __float_as_uint(ang));
        const int32_t p = h.parent[cur];
        cur             = p < 0 ? -1 : (p & ~kRightChild);
    }
}

__device__ __forceinline__ uint16_t floatToUnorm16(float x) { return uint16_t(fmaf(x, 65535.f, 0.5f)); }           // encoding.zig
__device__ __forceinline__ uint16_t floatToSnorm16(float x) { return uint16_t((x + 1.f) * (x > 0.f ? 32767.5f : 32768.f)); }

// serialize + Node.compressCenter, light_tree_builder.zig:615-636, light_tree.zig:40-54. Children of node k live at 2k + 1, 2k + 2.
__global__ void emitLightNodesKernel(LightBuildInput in, uint32_t max_leaf, Hierarchy h, StatArrays stats, const uint32_t* __restrict__ vals,
                                     const uint32_t* __restrict__ angles, ZygpuLightNode* __restrict__ nodes, uint32_t* __restrict__ middles) {
    const int n = int(in.num_lights);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n - 1) return;
    const bool     internal = i < n - 1;
    const uint32_t id       = internal ? uint32_t(i) : uint32_t(i - (n - 1));
    const int32_t  pw       = internal ? h.parent[id] : h.leaf_parent[id];
    auto count = [&](uint32_t k) { return h.last[k] - h.first[k] + 1u; };
    uint32_t slot = 0;
    if (pw >= 0) {
        const uint32_t p = uint32_t(pw & ~kRightChild);
        if (count(p) <= max_leaf) return;  // inside a leaf
        slot = 2 * p + 1 + (0 != (pw & kRightChild) ? 1u : 0u);
    }
    const NodeStats s     = internal ? loadStats(stats, id) : leafStats(in, vals[id]);
    const uint32_t  first = internal ? h.first[id] : id;
    const uint32_t  num   = internal ? count(id) : 1u;
    const bool      leaf  = num <= max_leaf;

    float4 cone;
    bool   two_sided;
    if (in.primitive) {
        if (internal) {
            const float3 axis = dominantAxis(s.cone, s.power);
            cone              = make_float4(axis.x, axis.y, axis.z, cosf(__uint_as_float(angles[id])));
            if (0.f == axis.x && 0.f == axis.y && 0.f == axis.z) cone = make_float4(0.f, 0.f, 1.f, -1.f);
        } else {
            const float4 nrm = in.cones[vals[id]];
            cone             = make_float4(nrm.x, nrm.y, nrm.z, 1.f);
        }
        two_sided = in.all_two_sided;
    } else {
        cone      = s.cone;
        two_sided = 0 != s.two_sided;
    }

    float variance = 0.f;
    if (s.num_powered > 0) {
        const float inv = 1.f / float(s.num_powered);
        const float ap  = s.power * inv;
        variance        = fabsf(s.sum_p2 * inv - ap * ap);
    }

    // the tree bounds: root box, radius cached (AABB.cacheRadius, aabb.zig:141-145)
    const NodeStats root = loadStats(stats, 0);
    const float3    te   = make_float3(root.hi.x - root.lo.x, root.hi.y - root.lo.y, root.hi.z - root.lo.z);
    const float     tr   = 0.5f * sqrtf(te.x * te.x + te.y * te.y + te.z * te.z);
    const float3    e    = make_float3(s.hi.x - s.lo.x, s.hi.y - s.lo.y, s.hi.z - s.lo.z);
    const float4    c    = make_float4(0.5f * (s.lo.x + s.hi.x), 0.5f * (s.lo.y + s.hi.y), 0.5f * (s.lo.z + s.hi.z),
                                       0.5f * sqrtf(e.x * e.x + e.y * e.y + e.z * e.z));
    const float q[4] = {(c.x - root.lo.x) / (0.f == te.x ? 1.f : te.x), (c.y - root.lo.y) / (0.f == te.y ? 1.f : te.y),
                        (c.z - root.lo.z) / (0.f == te.z ? 1.f : te.z), 0.f == tr ? 0.f : c.w / tr};
    const float cn[4] = {cone.x, cone.y, cone.z, cone.w};

    ZygpuLightNode out;
    for (int k = 0; k < 4; ++k) {
        out.center[k] = floatToUnorm16(fminf(fmaxf(q[k], 0.f), 1.f));
        out.cone[k]   = floatToSnorm16(clampUnit(cn[k]));
    }
    out.power      = s.power;
    out.variance   = variance;
    const uint32_t children_or_light = leaf ? in.first_order + first : 2 * id + 1;
    out.meta       = (leaf ? 0u : 1u) | (two_sided ? 2u : 0u) | (children_or_light << 2);
    out.num_lights = num;
    nodes[slot]    = out;
    uint32_t middle = 0;
    if (!leaf) {
        const uint32_t r = h.right[id];
        middle           = in.first_order + (0 != (r & kLeafBit) ? (r & ~kLeafBit) : h.first[r]);
    }
    middles[slot] = middle;
}

// Scratch comes from the device's stream-ordered pool (kept across builds: a compile builds one tree per emissive mesh part): two dozen
// cudaMalloc / cudaFree pairs per build cost 20 - 100 ms around well under a millisecond of kernels. The three arrays that outlive the
// build are ordinary allocations.
cudaStream_t g_alloc_stream = nullptr;

template <typename T>
cudaError_t lightAlloc(T*& p, size_t count, std::vector<void*>& scratch, bool keep = false) {
    void*             raw   = nullptr;
    const size_t      bytes = std::max<size_t>(count * sizeof(T), 16);
    const cudaError_t e     = keep ? cudaMalloc(&raw, bytes) : cudaMallocAsync(&raw, bytes, g_alloc_stream);
    if (cudaSuccess != e) return e;
    p = static_cast<T*>(raw);
    if (!keep) scratch.push_back(raw);
    return cudaSuccess;
}

inline uint32_t blocksFor(uint64_t n) { return uint32_t((n + kThreads - 1) / kThreads); }

}  // namespace

void freeLightBuildOutput(LightBuildOutput& out) {
    cudaFree(out.nodes);
    cudaFree(out.node_middles);
    cudaFree(out.order);
    out = LightBuildOutput{};
}

cudaError_t buildLightTreeOnDevice(const LightBuildInput& in, LightBuildOutput& out, cudaStream_t stream) {
    const uint32_t n = in.num_lights;
    if (n < 2) return cudaErrorInvalidValue;

    static bool pool_kept = false;
    if (!pool_kept) {  // keep what the pool has handed out once instead of returning it to the driver at every synchronisation
        int device = 0;
        cudaGetDevice(&device);
        cudaMemPool_t mem_pool = nullptr;
        if (cudaSuccess == cudaDeviceGetDefaultMemPool(&mem_pool, device)) {
            uint64_t threshold = ~0ull;
            cudaMemPoolSetAttribute(mem_pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
        pool_kept = true;
    }
    g_alloc_stream = stream;

    std::vector<void*> scratch;
    struct Cleanup {
        std::vector<void*>& s;
        cudaStream_t        stream;
        ~Cleanup() {
            for (void* p : s) cudaFreeAsync(p, stream);
        }
    } cleanup{scratch, stream};

    cudaEvent_t ev0, ev1;
    BUILD_OK(cudaEventCreate(&ev0));
    BUILD_OK(cudaEventCreate(&ev1));
    BUILD_OK(cudaEventRecord(ev0, stream));

    Bounds*   bounds;
    uint32_t* counters;
    BUILD_OK(lightAlloc(bounds, 1, scratch));
    BUILD_OK(lightAlloc(counters, 8, scratch));
    uint64_t *keys, *keys_sorted;
    uint32_t* vals;
    BUILD_OK(lightAlloc(keys, n, scratch));
    BUILD_OK(lightAlloc(keys_sorted, n, scratch));
    BUILD_OK(lightAlloc(vals, n, scratch));
    BUILD_OK(lightAlloc(out.order, n, scratch, true));

    initBoundsKernel<<<1, 32, 0, stream>>>(bounds, counters, 8);
    lightBoundsKernel<<<blocksFor(n), kThreads, 0, stream>>>(in, bounds);
    lightMortonKernel<<<blocksFor(n), kThreads, 0, stream>>>(in, bounds, keys, vals);
    size_t sort_bytes = 0;
    BUILD_OK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys, keys_sorted, vals, out.order, int(n), 0, 60, stream));
    char* sort_tmp;
    BUILD_OK(lightAlloc(sort_tmp, sort_bytes, scratch));
    BUILD_OK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys, keys_sorted, vals, out.order, int(n), 0, 60, stream));

    Hierarchy h;
    BUILD_OK(lightAlloc(h.left, n, scratch));
    BUILD_OK(lightAlloc(h.right, n, scratch));
    BUILD_OK(lightAlloc(h.first, n, scratch));
    BUILD_OK(lightAlloc(h.last, n, scratch));
    BUILD_OK(lightAlloc(h.parent, n, scratch));
    BUILD_OK(lightAlloc(h.leaf_parent, n, scratch));
    BUILD_OK(lightAlloc(h.flags, n, scratch));
    h.lo = h.hi = nullptr;
    BUILD_OK(cudaMemsetAsync(h.flags, 0, size_t(n) * sizeof(uint32_t), stream));
    hierarchyKernel<<<blocksFor(n), kThreads, 0, stream>>>(int(n), keys_sorted, h);

    StatArrays stats;
    BUILD_OK(lightAlloc(stats.lo, n, scratch));
    BUILD_OK(lightAlloc(stats.hi, n, scratch));
    BUILD_OK(lightAlloc(stats.cone, n, scratch));
    BUILD_OK(lightAlloc(stats.misc, n, scratch));
    aggregateKernel<<<blocksFor(n), kThreads, 0, stream>>>(in, h, stats, out.order);

    uint32_t* angles;
    BUILD_OK(lightAlloc(angles, n, scratch));
    BUILD_OK(cudaMemsetAsync(angles, 0, size_t(n) * sizeof(uint32_t), stream));
    if (in.primitive) coneAngleKernel<<<blocksFor(n), kThreads, 0, stream>>>(in, h, stats, out.order, angles);

    out.num_nodes = 2 * n - 1;
    BUILD_OK(lightAlloc(out.nodes, out.num_nodes, scratch, true));
    BUILD_OK(lightAlloc(out.node_middles, out.num_nodes, scratch, true));
    BUILD_OK(cudaMemsetAsync(out.nodes, 0, size_t(out.num_nodes) * sizeof(ZygpuLightNode), stream));
    BUILD_OK(cudaMemsetAsync(out.node_middles, 0, size_t(out.num_nodes) * sizeof(uint32_t), stream));
    emitLightNodesKernel<<<blocksFor(2 * uint64_t(n) - 1), kThreads, 0, stream>>>(in, in.primitive ? 4u : 1u, h, stats, out.order, angles, out.nodes,
                                                                                  out.node_middles);

    float4 root[3];
    BUILD_OK(cudaMemcpyAsync(&root[0], stats.lo, sizeof(float4), cudaMemcpyDeviceToHost, stream));
    BUILD_OK(cudaMemcpyAsync(&root[1], stats.hi, sizeof(float4), cudaMemcpyDeviceToHost, stream));
    BUILD_OK(cudaMemcpyAsync(&root[2], stats.misc, sizeof(float4), cudaMemcpyDeviceToHost, stream));
    BUILD_OK(cudaEventRecord(ev1, stream));
    BUILD_OK(cudaStreamSynchronize(stream));
    out.bounds_min[0] = root[0].x, out.bounds_min[1] = root[0].y, out.bounds_min[2] = root[0].z;
    out.bounds_max[0] = root[1].x, out.bounds_max[1] = root[1].y, out.bounds_max[2] = root[1].z;
    out.root_power    = root[2].x;
    BUILD_OK(cudaEventElapsedTime(&out.device_ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return cudaGetLastError();
}

}  // namespace zygpu
