// Device BVH build and refit, see build.cuh. HBM-bound integer / fp32 work: every kernel is one thread per
// triangle or per node with coalesced 16-byte accesses; the only library call is CUB's radix sort of the 64-bit Morton keys.
#include "build.cuh"

#include <cfloat>
#include <cub/cub.cuh>
#include <vector>

namespace zygpu {

namespace {

#include "lbvh.cuh"

__device__ __forceinline__ void loadTriangle(const uint32_t* __restrict__ indices, const float* __restrict__ positions, uint32_t t, float3& a,
                                             float3& b, float3& c) {
    const uint32_t i0 = indices[3 * size_t(t) + 0], i1 = indices[3 * size_t(t) + 1], i2 = indices[3 * size_t(t) + 2];
    a = make_float3(positions[3 * size_t(i0)], positions[3 * size_t(i0) + 1], positions[3 * size_t(i0) + 2]);
    b = make_float3(positions[3 * size_t(i1)], positions[3 * size_t(i1) + 1], positions[3 * size_t(i1) + 2]);
    c = make_float3(positions[3 * size_t(i2)], positions[3 * size_t(i2) + 1], positions[3 * size_t(i2) + 2]);
}

// Triangle.aabb (triangle.zig:18-24) per source triangle + the mesh bounds
__global__ void triangleBoundsKernel(MeshBuildInput in, float4* __restrict__ tri_lo, float4* __restrict__ tri_hi, Bounds* bounds) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (t < in.num_triangles) {
        float3 a, b, c;
        loadTriangle(in.indices, in.positions, t, a, b, c);
        lo        = make_float3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
        hi        = make_float3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
        tri_lo[t] = make_float4(lo.x, lo.y, lo.z, 0.f);
        tri_hi[t] = make_float4(hi.x, hi.y, hi.z, 0.f);
    }
    using Reduce = cub::BlockReduce<float, kThreads>;
    __shared__ typename Reduce::TempStorage tmp;
    float v[6] = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z};
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

__global__ void mortonKernel(uint32_t n, const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi, const Bounds* bounds,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float4 lo = tri_lo[t], hi = tri_hi[t];
    const float  c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    uint32_t     q[3];
    for (int a = 0; a < 3; ++a) {
        const float mn = orderedToFloat(bounds->lo[a]), mx = orderedToFloat(bounds->hi[a]);
        const float e  = mx - mn;
        const float f  = e > 0.f ? (c[a] - mn) / e : 0.f;
        q[a]           = uint32_t(fminf(fmaxf(f * 1048576.f, 0.f), 1048575.f));
    }
    keys[t] = (spread20(q[0]) << 2) | (spread20(q[1]) << 1) | spread20(q[2]);
    vals[t] = t;
}

// tree-order arrays (triangle_tree_builder.zig:190-199: triangles are rewritten in leaf order) + boxes in that order
__global__ void gatherKernel(MeshBuildInput in, const uint32_t* __restrict__ vals, const float4* __restrict__ tri_lo,
                             const float4* __restrict__ tri_hi, uint32_t* __restrict__ triangles, uint32_t* __restrict__ original,
                             uint16_t* __restrict__ parts, float4* __restrict__ sorted_lo, float4* __restrict__ sorted_hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= in.num_triangles) return;
    const uint32_t src        = vals[i];
    triangles[3 * size_t(i)]     = in.indices[3 * size_t(src)];
    triangles[3 * size_t(i) + 1] = in.indices[3 * size_t(src) + 1];
    triangles[3 * size_t(i) + 2] = in.indices[3 * size_t(src) + 2];
    original[i]               = src;
    parts[i]                  = in.parts[src];
    sorted_lo[i]              = tri_lo[src];
    sorted_hi[i]              = tri_hi[src];
}

__device__ __forceinline__ void childBox(const Hierarchy& h, const float4* __restrict__ sorted_lo, const float4* __restrict__ sorted_hi, uint32_t ref,
                                         float4& lo, float4& hi) {
    if (0 != (ref & kLeafBit)) {
        lo = sorted_lo[ref & ~kLeafBit];
        hi = sorted_hi[ref & ~kLeafBit];
    } else {
        lo = __ldcg(h.lo + ref);
        hi = __ldcg(h.hi + ref);
    }
}

// bottom-up boxes: the second thread to arrive at a node merges its children and goes on
__global__ void fitKernel(int n, Hierarchy h, const float4* __restrict__ sorted_lo, const float4* __restrict__ sorted_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t cur = h.leaf_parent[i] & ~kRightChild;
    while (cur >= 0) {
        if (0 == atomicAdd(h.flags + cur, 1u)) return;
        float4 alo, ahi, blo, bhi;
        childBox(h, sorted_lo, sorted_hi, h.left[cur], alo, ahi);
        childBox(h, sorted_lo, sorted_hi, h.right[cur], blo, bhi);
        h.lo[cur] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.f);
        h.hi[cur] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.f);
        __threadfence();
        const int32_t p = h.parent[cur];
        cur             = p < 0 ? -1 : (p & ~kRightChild);
    }
}

__device__ __forceinline__ uint32_t refCount(const Hierarchy& h, uint32_t ref) {
    return 0 != (ref & kLeafBit) ? 1u : h.last[ref] - h.first[ref] + 1u;
}
__device__ __forceinline__ bool refIsLeaf(const Hierarchy& h, uint32_t ref) { return refCount(h, ref) <= kMaxLeaf; }

// The 32-byte reference layout (bvh/node.zig:9-71): the children of internal node k live at 2k + 1 and 2k + 2, the root at 0; a
// subtree of at most three triangles is one leaf (they are consecutive in Morton order). Also the depth of the binary tree.
__global__ void emitBinaryKernel(int n, Hierarchy h, const float4* __restrict__ sorted_lo, const float4* __restrict__ sorted_hi,
                                 float4* __restrict__ nodes, uint32_t* __restrict__ max_depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n - 1) return;
    const bool     internal = i < n - 1;
    const uint32_t id       = internal ? uint32_t(i) : uint32_t(i - (n - 1));
    const int32_t  pw       = internal ? h.parent[id] : h.leaf_parent[id];
    uint32_t       slot     = 0;
    if (pw >= 0) {
        const uint32_t p = uint32_t(pw & ~kRightChild);
        if (refIsLeaf(h, p)) return;  // swallowed by a leaf further up
        slot = 2 * p + 1 + (0 != (pw & kRightChild) ? 1u : 0u);
    }
    float4 lo, hi;
    childBox(h, sorted_lo, sorted_hi, internal ? id : (kLeafBit | id), lo, hi);
    const bool     leaf  = !internal || refIsLeaf(h, id);
    const uint32_t first = internal ? h.first[id] : id;
    lo.w                 = __uint_as_float(leaf ? first : 2 * id + 1);
    hi.w                 = __uint_as_float(leaf ? refCount(h, internal ? id : (kLeafBit | id)) : 0u);
    nodes[2 * size_t(slot)]     = lo;
    nodes[2 * size_t(slot) + 1] = hi;
    if (leaf) {  // depth = number of ancestors + 1
        uint32_t depth = 1;
        int32_t  p     = pw;
        while (p >= 0) {
            depth += 1;
            p = h.parent[p & ~kRightChild];
        }
        atomicMax(max_depth, depth);
    }
}

// ---- wide collapse, one level per launch pair ----------------------------------------------------------------------------------

struct LevelItem {
    uint32_t ref;   // internal node that becomes this wide node
    uint32_t wide;  // its index
};

struct Gathered {
    uint32_t child[8];  // child references in gathering order
    uint32_t num;
};

__device__ __forceinline__ float boxArea(float4 lo, float4 hi) {
    const float dx = fmaxf(hi.x - lo.x, 0.f), dy = fmaxf(hi.y - lo.y, 0.f), dz = fmaxf(hi.z - lo.z, 0.f);
    return dx * dy + dx * dz + dy * dz;
}

// host/wide_bvh.cpp collapseWide step 1: up to eight children by repeatedly opening the inner child with the largest area
__global__ void gatherChildrenKernel(uint32_t num_items, const LevelItem* __restrict__ items, Hierarchy h, const float4* __restrict__ sorted_lo,
                                     const float4* __restrict__ sorted_hi, Gathered* __restrict__ gathered, uint64_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_items) return;
    const uint32_t ref = items[i].ref;
    Gathered       g;
    float          area[8];
    g.num        = 2;
    g.child[0]   = h.left[ref];
    g.child[1]   = h.right[ref];
    for (uint32_t k = 0; k < 2; ++k) {
        float4 lo, hi;
        childBox(h, sorted_lo, sorted_hi, g.child[k], lo, hi);
        area[k] = refIsLeaf(h, g.child[k]) ? -1.f : boxArea(lo, hi);
    }
    while (g.num < 8) {
        int   best      = -1;
        float best_area = -1.f;
        for (uint32_t k = 0; k < g.num; ++k) {
            if (area[k] > best_area) {
                best_area = area[k];
                best      = int(k);
            }
        }
        if (best < 0) break;
        const uint32_t open = g.child[best];
        const uint32_t c[2] = {h.left[open], h.right[open]};
        const uint32_t at[2] = {uint32_t(best), g.num};
        for (int k = 0; k < 2; ++k) {
            float4 lo, hi;
            childBox(h, sorted_lo, sorted_hi, c[k], lo, hi);
            g.child[at[k]] = c[k];
            area[at[k]]    = refIsLeaf(h, c[k]) ? -1.f : boxArea(lo, hi);
        }
        g.num += 1;
    }
    uint32_t inner = 0, tris = 0;
    for (uint32_t k = 0; k < g.num; ++k) {
        if (refIsLeaf(h, g.child[k])) {
            tris += refCount(h, g.child[k]);
        } else {
            inner += 1;
        }
    }
    for (uint32_t k = g.num; k < 8; ++k) g.child[k] = 0;
    gathered[i] = g;
    counts[i]   = (uint64_t(tris) << 32) | inner;
}

struct NodeFrame {
    float p[3];
    int   ex[3];
};

// host/wide_bvh.cpp collapseWide step 2: grid origin and power-of-two cell sizes that cover the node box with 255 cells
__device__ __forceinline__ NodeFrame nodeFrame(float4 lo, float4 hi) {
    NodeFrame   f;
    const float l[3] = {lo.x, lo.y, lo.z}, u[3] = {hi.x, hi.y, hi.z};
    for (int a = 0; a < 3; ++a) {
        f.p[a]              = l[a];
        const double extent = double(u[a]) - double(l[a]);
        int          e      = extent > 0.0 ? int(ceil(log2(extent / 255.0))) : -126;
        e                   = max(e, -126);
        while (ldexp(1.0, e) * 255.0 < extent) ++e;
        f.ex[a] = e;
    }
    return f;
}

__device__ __forceinline__ void quantiseBox(const NodeFrame& f, float4 lo, float4 hi, uint8_t qlo[3], uint8_t qhi[3]) {
    const float l[3] = {lo.x, lo.y, lo.z}, u[3] = {hi.x, hi.y, hi.z};
    for (int a = 0; a < 3; ++a) {
        const double cell = ldexp(1.0, f.ex[a]);
        const double p    = double(f.p[a]);
        int          ql   = int(floor((double(l[a]) - p) / cell));
        int          qh   = int(ceil((double(u[a]) - p) / cell));
        ql                = min(max(ql, 0), 255);
        qh                = min(max(qh, 0), 255);
        while (ql > 0 && p + ql * cell > double(l[a])) --ql;
        while (qh < 255 && p + qh * cell < double(u[a])) ++qh;
        qlo[a] = uint8_t(ql);
        qhi[a] = uint8_t(qh);
    }
}

// host/wide_bvh.cpp collapseWide step 3: greedy assignment of children to the slot whose diagonal they lie along
__device__ __forceinline__ void assignSlots(uint32_t num, const float4* clo, const float4* chi, float4 nlo, float4 nhi, int child_in_slot[8]) {
    float cost[8][8];
    for (uint32_t c = 0; c < num; ++c) {
        const float d[3] = {0.5f * (clo[c].x + chi[c].x) - 0.5f * (nlo.x + nhi.x), 0.5f * (clo[c].y + chi[c].y) - 0.5f * (nlo.y + nhi.y),
                            0.5f * (clo[c].z + chi[c].z) - 0.5f * (nlo.z + nhi.z)};
        for (int s = 0; s < 8; ++s) {
            const float sx = (s & 4) ? 1.f : -1.f, sy = (s & 2) ? 1.f : -1.f, sz = (s & 1) ? 1.f : -1.f;
            cost[c][s]     = d[0] * sx + d[1] * sy + d[2] * sz;
        }
    }
    uint32_t slot_used = 0, child_done = 0;
    for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
    for (uint32_t k = 0; k < num; ++k) {
        float best = -FLT_MAX;
        int   bc = -1, bs = -1;
        for (uint32_t c = 0; c < num; ++c) {
            if (0 != (child_done & (1u << c))) continue;
            for (int s = 0; s < 8; ++s) {
                if (0 != (slot_used & (1u << s))) continue;
                if (cost[c][s] > best || bc < 0) {
                    best = cost[c][s];
                    bc   = int(c);
                    bs   = s;
                }
            }
        }
        child_done |= 1u << bc;
        slot_used |= 1u << bs;
        child_in_slot[bs] = bc;
    }
}

__device__ __forceinline__ void writeTriangleRecord(float4* __restrict__ rec, const uint32_t* __restrict__ triangles, const float* __restrict__ positions,
                                                    uint32_t prim, float4 gate_lo, float4 gate_hi) {
    float3 a, b, c;
    loadTriangle(triangles, positions, prim, a, b, c);
    rec[0] = make_float4(a.x, a.y, a.z, __uint_as_float(prim));
    rec[1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, gate_lo.x);
    rec[2] = make_float4(c.x - a.x, c.y - a.y, c.z - a.z, gate_lo.y);
    rec[3] = make_float4(gate_lo.z, gate_hi.x, gate_hi.y, gate_hi.z);
}

struct WideNodeWords {  // host/wide_bvh.hpp WideNode as six 16-byte words
    float4 w[6];
};

__device__ __forceinline__ void packNode(WideNodeWords& out, const NodeFrame& f, uint32_t imask, uint32_t child_base, uint32_t tri_base,
                                         const uint8_t meta[8], const uint8_t qlo[3][8], const uint8_t qhi[3][8]) {
    auto word = [](const uint8_t* b) { return uint32_t(b[0]) | (uint32_t(b[1]) << 8) | (uint32_t(b[2]) << 16) | (uint32_t(b[3]) << 24); };
    const uint32_t e = uint32_t(f.ex[0] + 127) | (uint32_t(f.ex[1] + 127) << 8) | (uint32_t(f.ex[2] + 127) << 16) | (imask << 24);
    out.w[0] = make_float4(f.p[0], f.p[1], f.p[2], __uint_as_float(e));
    out.w[1] = make_float4(__uint_as_float(child_base), __uint_as_float(tri_base), __uint_as_float(word(meta)), __uint_as_float(word(meta + 4)));
    out.w[2] = make_float4(__uint_as_float(word(qlo[0])), __uint_as_float(word(qlo[0] + 4)), __uint_as_float(word(qlo[1])),
                           __uint_as_float(word(qlo[1] + 4)));
    out.w[3] = make_float4(__uint_as_float(word(qlo[2])), __uint_as_float(word(qlo[2] + 4)), __uint_as_float(word(qhi[0])),
                           __uint_as_float(word(qhi[0] + 4)));
    out.w[4] = make_float4(__uint_as_float(word(qhi[1])), __uint_as_float(word(qhi[1] + 4)), __uint_as_float(word(qhi[2])),
                           __uint_as_float(word(qhi[2] + 4)));
    out.w[5] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// host/wide_bvh.cpp collapseWide steps 2-4 for one level: `offsets` is the exclusive scan of the level's (triangles << 32 | inner)
// counts; inner children get consecutive node ids from `node_base`, leaf slots consecutive records from `tri_base`
__global__ void emitWideKernel(uint32_t num_items, const LevelItem* __restrict__ items, const Gathered* __restrict__ gathered,
                               const uint64_t* __restrict__ offsets, uint32_t node_base, uint32_t tri_base, Hierarchy h,
                               const float4* __restrict__ sorted_lo, const float4* __restrict__ sorted_hi, const uint32_t* __restrict__ triangles,
                               const float* __restrict__ positions, float4* __restrict__ wide_nodes, float4* __restrict__ wide_tris,
                               LevelItem* __restrict__ next_items) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_items) return;
    const LevelItem item = items[i];
    const Gathered  g    = gathered[i];
    const uint64_t  off  = offsets[i];
    const uint32_t  first_child = node_base + uint32_t(off & 0xFFFFFFFFu);
    const uint32_t  first_tri   = tri_base + uint32_t(off >> 32);

    float4 clo[8], chi[8];
    for (uint32_t c = 0; c < g.num; ++c) childBox(h, sorted_lo, sorted_hi, g.child[c], clo[c], chi[c]);
    float4 nlo = clo[0], nhi = chi[0];
    for (uint32_t c = 1; c < g.num; ++c) {
        nlo = make_float4(fminf(nlo.x, clo[c].x), fminf(nlo.y, clo[c].y), fminf(nlo.z, clo[c].z), 0.f);
        nhi = make_float4(fmaxf(nhi.x, chi[c].x), fmaxf(nhi.y, chi[c].y), fmaxf(nhi.z, chi[c].z), 0.f);
    }
    const NodeFrame frame = nodeFrame(nlo, nhi);
    int             child_in_slot[8];
    assignSlots(g.num, clo, chi, nlo, nhi, child_in_slot);

    uint8_t  meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint8_t  qlo[3][8], qhi[3][8];
    uint32_t imask = 0, inner_rank = 0, tri_offset = 0;
    for (int s = 0; s < 8; ++s) {
        for (int a = 0; a < 3; ++a) qlo[a][s] = qhi[a][s] = 0;
        const int c = child_in_slot[s];
        if (c < 0) continue;
        uint8_t ql[3], qh[3];
        quantiseBox(frame, clo[c], chi[c], ql, qh);
        for (int a = 0; a < 3; ++a) {
            qlo[a][s] = ql[a];
            qhi[a][s] = qh[a];
        }
        const uint32_t ref = g.child[c];
        if (!refIsLeaf(h, ref)) {
            imask |= 1u << s;
            meta[s]                                    = uint8_t((1u << 5) | (24u + uint32_t(s)));
            next_items[uint32_t(off & 0xFFFFFFFFu) + inner_rank] = LevelItem{ref, first_child + inner_rank};
            inner_rank += 1;
        } else {
            const uint32_t count = refCount(h, ref);
            const uint32_t start = 0 != (ref & kLeafBit) ? (ref & ~kLeafBit) : h.first[ref];
            meta[s]              = uint8_t((((1u << count) - 1u) << 5) | tri_offset);
            for (uint32_t k = 0; k < count; ++k) {
                writeTriangleRecord(wide_tris + 4 * size_t(first_tri + tri_offset + k), triangles, positions, start + k, clo[c], chi[c]);
            }
            tri_offset += count;
        }
    }
    WideNodeWords words;
    packNode(words, frame, imask, first_child, first_tri, meta, qlo, qhi);
    float4* dst = wide_nodes + 6 * size_t(item.wide);
    for (int k = 0; k < 6; ++k) dst[k] = words.w[k];
}

// bounding sphere of the referenced vertices around the centre of the root box (host/wide_bvh.cpp buildWideBvh)
__global__ void boundRadiusKernel(uint32_t num_triangles, const uint32_t* __restrict__ triangles, const float* __restrict__ positions, float3 centre,
                                  uint32_t* __restrict__ radius2_bits) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    float          r2 = 0.f;
    if (t < num_triangles) {
        float3 v[3];
        loadTriangle(triangles, positions, t, v[0], v[1], v[2]);
        for (int k = 0; k < 3; ++k) {
            const float dx = v[k].x - centre.x, dy = v[k].y - centre.y, dz = v[k].z - centre.z;
            r2             = fmaxf(r2, dx * dx + dy * dy + dz * dz);
        }
    }
    using Reduce = cub::BlockReduce<float, kThreads>;
    __shared__ typename Reduce::TempStorage tmp;
    const float m = Reduce(tmp).Reduce(r2, cub::Max());
    if (0 == threadIdx.x) atomicMax(radius2_bits, __float_as_uint(m));  // non-negative floats order like their bits
}

template <typename T>
cudaError_t deviceAlloc(T*& p, size_t count, std::vector<void*>& scratch, bool keep = false) {
    void*             raw = nullptr;
    const cudaError_t e   = cudaMalloc(&raw, std::max<size_t>(count * sizeof(T), 16));
    if (cudaSuccess != e) return e;
    p = static_cast<T*>(raw);
    if (!keep) scratch.push_back(raw);
    return cudaSuccess;
}

inline uint32_t blocksFor(uint64_t n) { return uint32_t((n + kThreads - 1) / kThreads); }

// ---- refit -------------------------------------------------------------------------------------------------------------------------

// one wide node: leaf slots take the boxes of their triangles (records are rewritten from the moved vertices), inner slots the exact
// boxes of the level below; the node is re-quantised around the union
__global__ void refitWideLevelKernel(uint32_t begin, uint32_t end, float4* __restrict__ wide_nodes, float4* __restrict__ wide_tris,
                                     const uint32_t* __restrict__ triangles, const float* __restrict__ positions, float4* __restrict__ exact_lo,
                                     float4* __restrict__ exact_hi) {
    const uint32_t n = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= end) return;
    float4*        node = wide_nodes + 6 * size_t(n);
    const float4   w0 = node[0], w1 = node[1];
    const uint32_t imask      = __float_as_uint(w0.w) >> 24;
    const uint32_t child_base = __float_as_uint(w1.x), tri_base = __float_as_uint(w1.y);
    const uint32_t meta_lo = __float_as_uint(w1.z), meta_hi = __float_as_uint(w1.w);

    float4   clo[8], chi[8];
    bool     used[8];
    float4   nlo = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), nhi = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
    uint32_t inner_rank = 0;
    for (int s = 0; s < 8; ++s) {
        const uint32_t meta = ((s < 4 ? meta_lo : meta_hi) >> (8 * (s & 3))) & 0xffu;
        used[s]             = 0 != meta;
        if (!used[s]) continue;
        if (0 != (imask & (1u << s))) {
            clo[s] = exact_lo[child_base + inner_rank];
            chi[s] = exact_hi[child_base + inner_rank];
            inner_rank += 1;
        } else {
            const uint32_t count = __popc(meta >> 5), offset = meta & 31u;
            float3         lo = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), hi = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
            for (uint32_t k = 0; k < count; ++k) {
                const uint32_t prim = __float_as_uint(wide_tris[4 * size_t(tri_base + offset + k)].w);
                float3         v[3];
                loadTriangle(triangles, positions, prim, v[0], v[1], v[2]);
                for (int j = 0; j < 3; ++j) {
                    lo = make_float3(fminf(lo.x, v[j].x), fminf(lo.y, v[j].y), fminf(lo.z, v[j].z));
                    hi = make_float3(fmaxf(hi.x, v[j].x), fmaxf(hi.y, v[j].y), fmaxf(hi.z, v[j].z));
                }
            }
            clo[s] = make_float4(lo.x, lo.y, lo.z, 0.f);
            chi[s] = make_float4(hi.x, hi.y, hi.z, 0.f);
            for (uint32_t k = 0; k < count; ++k) {
                float4*        rec  = wide_tris + 4 * size_t(tri_base + offset + k);
                const uint32_t prim = __float_as_uint(rec[0].w);
                writeTriangleRecord(rec, triangles, positions, prim, clo[s], chi[s]);
            }
        }
        nlo = make_float4(fminf(nlo.x, clo[s].x), fminf(nlo.y, clo[s].y), fminf(nlo.z, clo[s].z), 0.f);
        nhi = make_float4(fmaxf(nhi.x, chi[s].x), fmaxf(nhi.y, chi[s].y), fmaxf(nhi.z, chi[s].z), 0.f);
    }
    exact_lo[n] = nlo;
    exact_hi[n] = nhi;

    const NodeFrame frame = nodeFrame(nlo, nhi);
    uint8_t         meta[8], qlo[3][8], qhi[3][8];
    for (int s = 0; s < 8; ++s) {
        meta[s] = uint8_t(((s < 4 ? meta_lo : meta_hi) >> (8 * (s & 3))) & 0xffu);
        for (int a = 0; a < 3; ++a) qlo[a][s] = qhi[a][s] = 0;
        if (!used[s]) continue;
        uint8_t ql[3], qh[3];
        quantiseBox(frame, clo[s], chi[s], ql, qh);
        for (int a = 0; a < 3; ++a) {
            qlo[a][s] = ql[a];
            qhi[a][s] = qh[a];
        }
    }
    WideNodeWords words;
    packNode(words, frame, imask, child_base, tri_base, meta, qlo, qhi);
    for (int k = 0; k < 5; ++k) node[k] = words.w[k];
}

// binary tree: leaves from their triangles, then `depth` sweeps in which every inner node takes the union of its children
__global__ void refitBinaryLeavesKernel(uint32_t num_nodes, float4* __restrict__ nodes, const uint32_t* __restrict__ triangles,
                                        const float* __restrict__ positions) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_nodes) return;
    float4         lo = nodes[2 * size_t(n)], hi = nodes[2 * size_t(n) + 1];
    const uint32_t count = __float_as_uint(hi.w), start = __float_as_uint(lo.w);
    if (0 == count) return;
    float3 l = make_float3(FLT_MAX, FLT_MAX, FLT_MAX), u = make_float3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (uint32_t k = 0; k < count; ++k) {
        float3 v[3];
        loadTriangle(triangles, positions, start + k, v[0], v[1], v[2]);
        for (int j = 0; j < 3; ++j) {
            l = make_float3(fminf(l.x, v[j].x), fminf(l.y, v[j].y), fminf(l.z, v[j].z));
            u = make_float3(fmaxf(u.x, v[j].x), fmaxf(u.y, v[j].y), fmaxf(u.z, v[j].z));
        }
    }
    nodes[2 * size_t(n)]     = make_float4(l.x, l.y, l.z, lo.w);
    nodes[2 * size_t(n) + 1] = make_float4(u.x, u.y, u.z, hi.w);
}

__global__ void refitBinaryInnerKernel(uint32_t num_nodes, float4* __restrict__ nodes) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_nodes) return;
    const float4   lo = nodes[2 * size_t(n)], hi = nodes[2 * size_t(n) + 1];
    const uint32_t c  = __float_as_uint(lo.w);
    if (0 != __float_as_uint(hi.w) || 0 == c || c + 1 >= num_nodes) return;  // a leaf, or an unused slot
    const float4 alo = nodes[2 * size_t(c)], ahi = nodes[2 * size_t(c) + 1], blo = nodes[2 * size_t(c) + 2], bhi = nodes[2 * size_t(c) + 3];
    nodes[2 * size_t(n)]     = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), lo.w);
    nodes[2 * size_t(n) + 1] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), hi.w);
}

}  // namespace

void freeMeshBuildOutput(MeshBuildOutput& out) {
    cudaFree(out.wide_nodes);
    cudaFree(out.wide_tris);
    cudaFree(out.binary_nodes);
    cudaFree(out.triangles);
    cudaFree(out.original);
    cudaFree(out.triangle_parts);
    out = MeshBuildOutput{};
}

namespace {
// bump allocator over one cudaMalloc
struct Arena {
    char*  base = nullptr;
    size_t size = 0, used = 0;
    ~Arena() { cudaFree(base); }
    static size_t padded(size_t bytes) { return (std::max<size_t>(bytes, 16) + 255) & ~size_t(255); }
    template <typename T>
    void reserve(size_t count) { size += padded(count * sizeof(T)); }
    template <typename T>
    cudaError_t take(T*& p, size_t count) {
        const size_t bytes = padded(count * sizeof(T));
        if (used + bytes > size) return cudaErrorMemoryAllocation;
        p = reinterpret_cast<T*>(base + used);
        used += bytes;
        return cudaSuccess;
    }
};
}  // namespace

cudaError_t buildMeshOnDevice(const MeshBuildInput& in, MeshBuildOutput& out, cudaStream_t stream) {
    const uint32_t n = in.num_triangles;
    if (n < 4) return cudaErrorInvalidValue;

    // All scratch comes out of one allocation and the six arrays that outlive the build are made before the clock starts: ~35 cudaMalloc
    // calls inside the timed region cost 10 - 50 ms for a million triangles, against 1.6 ms of kernels (profiles/r02_build_launches.md).
    Arena arena;
    std::vector<void*> scratch;  // unused: the outputs belong to the caller (freeMeshBuildOutput), everything else to the arena
    auto deviceAlloc = [&arena](auto*& p, size_t count, std::vector<void*>&, bool keep = false) -> cudaError_t {
        using T = std::remove_reference_t<decltype(*p)>;
        if (!keep) return arena.take(p, count);
        void*             raw = nullptr;
        const cudaError_t e   = cudaMalloc(&raw, std::max<size_t>(count * sizeof(T), 16));
        if (cudaSuccess == e) p = static_cast<T*>(raw);
        return e;
    };

    size_t sort_bytes = 0, scan_bytes = 0;
    BUILD_OK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                             (uint32_t*)nullptr, int(n), 0, 60, stream));
    BUILD_OK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, int(n) + 1, stream));
    for (int k = 0; k < 6; ++k) arena.reserve<float4>(n);           // tri_lo / hi, sorted_lo / hi, h.lo / hi
    arena.reserve<Bounds>(1);
    arena.reserve<uint32_t>(8);
    for (int k = 0; k < 2; ++k) arena.reserve<uint64_t>(n);         // keys
    for (int k = 0; k < 2; ++k) arena.reserve<uint32_t>(n);         // vals
    arena.reserve<char>(sort_bytes);
    for (int k = 0; k < 7; ++k) arena.reserve<uint32_t>(n);         // h.left / right / first / last / parent / leaf_parent / flags
    for (int k = 0; k < 2; ++k) arena.reserve<LevelItem>(n);
    arena.reserve<Gathered>(n);
    for (int k = 0; k < 2; ++k) arena.reserve<uint64_t>(size_t(n) + 1);
    arena.reserve<char>(scan_bytes);
    arena.size += 64 * 256;  // slack for members whose type is wider than assumed above
    BUILD_OK(cudaMalloc(reinterpret_cast<void**>(&arena.base), arena.size));

    BUILD_OK(deviceAlloc(out.triangles, size_t(n) * 3, scratch, true));
    BUILD_OK(deviceAlloc(out.original, n, scratch, true));
    BUILD_OK(deviceAlloc(out.triangle_parts, n, scratch, true));
    out.num_binary_nodes = 2 * n - 1;
    BUILD_OK(deviceAlloc(out.binary_nodes, size_t(out.num_binary_nodes) * 2, scratch, true));
    BUILD_OK(deviceAlloc(out.wide_nodes, size_t(n) * 6, scratch, true));
    BUILD_OK(deviceAlloc(out.wide_tris, size_t(n) * 4, scratch, true));

    cudaEvent_t ev0, ev1;
    BUILD_OK(cudaEventCreate(&ev0));
    BUILD_OK(cudaEventCreate(&ev1));
    BUILD_OK(cudaEventRecord(ev0, stream));

    float4 *tri_lo, *tri_hi, *sorted_lo, *sorted_hi;
    BUILD_OK(deviceAlloc(tri_lo, n, scratch));
    BUILD_OK(deviceAlloc(tri_hi, n, scratch));
    BUILD_OK(deviceAlloc(sorted_lo, n, scratch));
    BUILD_OK(deviceAlloc(sorted_hi, n, scratch));
    Bounds*   bounds;
    uint32_t* counters;  // [0] binary max depth, [1] radius^2 bits
    BUILD_OK(deviceAlloc(bounds, 1, scratch));
    BUILD_OK(deviceAlloc(counters, 8, scratch));
    uint64_t *keys, *keys_sorted;
    uint32_t *vals, *vals_sorted;
    BUILD_OK(deviceAlloc(keys, n, scratch));
    BUILD_OK(deviceAlloc(keys_sorted, n, scratch));
    BUILD_OK(deviceAlloc(vals, n, scratch));
    BUILD_OK(deviceAlloc(vals_sorted, n, scratch));

    // 1. boxes, bounds, Morton keys, sort
    initBoundsKernel<<<1, 32, 0, stream>>>(bounds, counters, 8);
    triangleBoundsKernel<<<blocksFor(n), kThreads, 0, stream>>>(in, tri_lo, tri_hi, bounds);
    mortonKernel<<<blocksFor(n), kThreads, 0, stream>>>(n, tri_lo, tri_hi, bounds, keys, vals);
    char* sort_tmp;
    BUILD_OK(deviceAlloc(sort_tmp, sort_bytes, scratch));
    BUILD_OK(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys, keys_sorted, vals, vals_sorted, int(n), 0, 60, stream));

    // 2. tree-order triangle arrays
    gatherKernel<<<blocksFor(n), kThreads, 0, stream>>>(in, vals_sorted, tri_lo, tri_hi, out.triangles, out.original, out.triangle_parts, sorted_lo,
                                                        sorted_hi);

    // 3. binary radix tree + boxes
    Hierarchy h;
    BUILD_OK(deviceAlloc(h.left, n, scratch));
    BUILD_OK(deviceAlloc(h.right, n, scratch));
    BUILD_OK(deviceAlloc(h.first, n, scratch));
    BUILD_OK(deviceAlloc(h.last, n, scratch));
    BUILD_OK(deviceAlloc(h.parent, n, scratch));
    BUILD_OK(deviceAlloc(h.leaf_parent, n, scratch));
    BUILD_OK(deviceAlloc(h.lo, n, scratch));
    BUILD_OK(deviceAlloc(h.hi, n, scratch));
    BUILD_OK(deviceAlloc(h.flags, n, scratch));
    BUILD_OK(cudaMemsetAsync(h.flags, 0, size_t(n) * sizeof(uint32_t), stream));
    hierarchyKernel<<<blocksFor(n), kThreads, 0, stream>>>(int(n), keys_sorted, h);
    fitKernel<<<blocksFor(n), kThreads, 0, stream>>>(int(n), h, sorted_lo, sorted_hi);

    // 4. reference-layout binary nodes
    BUILD_OK(cudaMemsetAsync(out.binary_nodes, 0, size_t(out.num_binary_nodes) * 2 * sizeof(float4), stream));
    emitBinaryKernel<<<blocksFor(2 * uint64_t(n) - 1), kThreads, 0, stream>>>(int(n), h, sorted_lo, sorted_hi, out.binary_nodes, counters);

    // 5. wide collapse, level by level: every wide node is an internal node of more than three triangles, so n bounds their number
    LevelItem *items, *next_items;
    Gathered*  gathered;
    uint64_t * counts, *offsets;
    BUILD_OK(deviceAlloc(items, n, scratch));
    BUILD_OK(deviceAlloc(next_items, n, scratch));
    BUILD_OK(deviceAlloc(gathered, n, scratch));
    BUILD_OK(deviceAlloc(counts, size_t(n) + 1, scratch));
    BUILD_OK(deviceAlloc(offsets, size_t(n) + 1, scratch));
    char* scan_tmp;
    BUILD_OK(deviceAlloc(scan_tmp, scan_bytes, scratch));

    const LevelItem root{0u, 0u};
    BUILD_OK(cudaMemcpyAsync(items, &root, sizeof(root), cudaMemcpyHostToDevice, stream));
    uint32_t level_items = 1, num_nodes = 1, num_tris = 0, depth = 0;
    while (level_items > 0) {
        depth += 1;
        gatherChildrenKernel<<<blocksFor(level_items), kThreads, 0, stream>>>(level_items, items, h, sorted_lo, sorted_hi, gathered, counts);
        BUILD_OK(cudaMemsetAsync(counts + level_items, 0, sizeof(uint64_t), stream));
        BUILD_OK(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, counts, offsets, int(level_items) + 1, stream));
        emitWideKernel<<<blocksFor(level_items), kThreads, 0, stream>>>(level_items, items, gathered, offsets, num_nodes, num_tris, h, sorted_lo,
                                                                         sorted_hi, out.triangles, in.positions, out.wide_nodes, out.wide_tris,
                                                                         next_items);
        uint64_t total = 0;
        BUILD_OK(cudaMemcpyAsync(&total, offsets + level_items, sizeof(total), cudaMemcpyDeviceToHost, stream));
        BUILD_OK(cudaStreamSynchronize(stream));
        level_items = uint32_t(total & 0xFFFFFFFFu);
        num_nodes += level_items;
        num_tris += uint32_t(total >> 32);
        std::swap(items, next_items);
    }
    if (num_tris != n) return cudaErrorUnknown;
    out.num_wide_nodes = num_nodes;
    out.wide_max_depth = depth;

    // 6. bounds, bounding sphere
    Bounds   hb;
    uint32_t hc[8];
    BUILD_OK(cudaMemcpyAsync(&hb, bounds, sizeof(hb), cudaMemcpyDeviceToHost, stream));
    BUILD_OK(cudaStreamSynchronize(stream));
    auto ordered = [](int i) {
        const int  b = i >= 0 ? i : i ^ 0x7FFFFFFF;
        float      f;
        memcpy(&f, &b, 4);
        return f;
    };
    double centre[3];
    for (int a = 0; a < 3; ++a) {
        out.aabb_min[a]     = ordered(hb.lo[a]);
        out.aabb_max[a]     = ordered(hb.hi[a]);
        centre[a]           = 0.5 * (double(out.aabb_min[a]) + double(out.aabb_max[a]));
        out.bound_center[a] = float(centre[a]);
    }
    boundRadiusKernel<<<blocksFor(n), kThreads, 0, stream>>>(n, out.triangles, in.positions,
                                                             make_float3(out.bound_center[0], out.bound_center[1], out.bound_center[2]), counters + 1);
    BUILD_OK(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, stream));
    BUILD_OK(cudaEventRecord(ev1, stream));
    BUILD_OK(cudaStreamSynchronize(stream));
    out.binary_max_depth = hc[0];
    float r2;
    memcpy(&r2, &hc[1], 4);
    // the centre was rounded to fp32 and the distances were taken in fp32: widen a little more than the host's double-precision path
    out.bound_radius = sqrtf(r2) * (1.f + 1e-5f) + FLT_MIN;
    BUILD_OK(cudaEventElapsedTime(&out.device_ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return cudaGetLastError();
}

cudaError_t refitMeshOnDevice(float4* wide_nodes, float4* wide_tris, float4* binary_nodes, const uint32_t* triangles, const float* positions,
                              const uint32_t* level_offsets, uint32_t num_levels, uint32_t num_binary_nodes, uint32_t binary_max_depth,
                              float bound[4], cudaStream_t stream) {
    if (0 == num_levels) return cudaErrorInvalidValue;
    const uint32_t num_wide = level_offsets[num_levels];
    float4 *       exact_lo = nullptr, *exact_hi = nullptr;
    BUILD_OK(cudaMalloc(&exact_lo, size_t(num_wide) * sizeof(float4)));
    if (cudaSuccess != cudaMalloc(&exact_hi, size_t(num_wide) * sizeof(float4))) {
        cudaFree(exact_lo);
        return cudaErrorMemoryAllocation;
    }
    for (uint32_t level = num_levels; level-- > 0;) {
        const uint32_t begin = level_offsets[level], end = level_offsets[level + 1];
        if (end > begin) {
            refitWideLevelKernel<<<blocksFor(end - begin), kThreads, 0, stream>>>(begin, end, wide_nodes, wide_tris, triangles, positions, exact_lo,
                                                                                  exact_hi);
        }
    }
    refitBinaryLeavesKernel<<<blocksFor(num_binary_nodes), kThreads, 0, stream>>>(num_binary_nodes, binary_nodes, triangles, positions);
    for (uint32_t d = 0; d < binary_max_depth; ++d) {
        refitBinaryInnerKernel<<<blocksFor(num_binary_nodes), kThreads, 0, stream>>>(num_binary_nodes, binary_nodes);
    }
    float4 root[2];
    cudaMemcpyAsync(&root[0], exact_lo, sizeof(float4), cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(&root[1], exact_hi, sizeof(float4), cudaMemcpyDeviceToHost, stream);
    const cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(exact_lo);
    cudaFree(exact_hi);
    if (cudaSuccess != e) return e;
    // a sphere around the centre of the new root box that contains the box (conservative: culling only)
    const float c[3] = {0.5f * (root[0].x + root[1].x), 0.5f * (root[0].y + root[1].y), 0.5f * (root[0].z + root[1].z)};
    const float d[3] = {root[1].x - c[0], root[1].y - c[1], root[1].z - c[2]};
    bound[0] = c[0], bound[1] = c[1], bound[2] = c[2];
    bound[3] = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * (1.f + 1e-5f) + FLT_MIN;
    return cudaGetLastError();
}

}  // namespace zygpu
