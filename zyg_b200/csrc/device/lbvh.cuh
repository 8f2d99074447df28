// Pieces shared by the device builders (build.cu: triangle BVH, light_build.cu: light trees): ordered-int float atomics, 60-bit
// Morton keys, Karras' binary radix tree over sorted keys. Included inside an anonymous namespace of each translation unit.
// (the includer provides <cfloat>, <cstdint> and <cuda_runtime.h> before opening its namespace).
#pragma once

constexpr uint32_t kLeafBit   = 0x80000000u;  // child reference: a single triangle (position in Morton order)
constexpr uint32_t kMaxLeaf   = 3;            // triangles per leaf slot of a wide node (unary count in 3 bits)
constexpr int      kThreads   = 256;

#define BUILD_OK(expr)                          \
    do {                                        \
        const cudaError_t e_ = (expr);          \
        if (cudaSuccess != e_) return e_;       \
    } while (0)

__device__ __forceinline__ int floatToOrdered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float orderedToFloat(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

struct Bounds {
    int lo[3], hi[3];  // ordered-int encoded
};

__global__ void initBoundsKernel(Bounds* b, uint32_t* counters, uint32_t num_counters) {
    if (0 == threadIdx.x) {
        for (int a = 0; a < 3; ++a) {
            b->lo[a] = floatToOrdered(FLT_MAX);
            b->hi[a] = floatToOrdered(-FLT_MAX);
        }
    }
    if (threadIdx.x < num_counters) counters[threadIdx.x] = 0;
}

__device__ __forceinline__ uint64_t spread20(uint32_t v) {  // 20 bits -> every third bit of 60
    uint64_t x = v & 0xFFFFFu;
    x          = (x | (x << 32)) & 0x000F00000000FFFFull;
    x          = (x | (x << 16)) & 0x000F0000FF0000FFull;
    x          = (x | (x << 8)) & 0x000F00F00F00F00Full;
    x          = (x | (x << 4)) & 0x00C30C30C30C30C3ull;
    x          = (x | (x << 2)) & 0x0249249249249249ull;
    return x;
}

// Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees": keys are made unique by the position
__device__ __forceinline__ int keyDelta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(uint32_t(i) ^ uint32_t(j));
    return __clzll((long long)(a ^ b));
}

struct Hierarchy {
    uint32_t* left;         // per internal node: child reference (kLeafBit | position, or internal index)
    uint32_t* right;
    uint32_t* first;        // per internal node: covered range of positions [first, last]
    uint32_t* last;
    int32_t*  parent;       // per internal node: parent internal node (-1 for the root); bit 30 set when it is the right child
    int32_t*  leaf_parent;  // per position
    float4*   lo;           // per internal node: box
    float4*   hi;
    uint32_t* flags;        // per internal node: bottom-up arrival counter
};

constexpr int32_t kRightChild = 0x40000000;

__global__ void hierarchyKernel(int n, const uint64_t* __restrict__ keys, Hierarchy h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d     = keyDelta(keys, n, i, i + 1) - keyDelta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin  = keyDelta(keys, n, i, i - d);
    int       lmax  = 2;
    while (keyDelta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2) {
        if (keyDelta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    }
    const int j     = i + l * d;
    const int dnode = keyDelta(keys, n, i, j);
    int       s     = 0;
    int       t     = l;
    do {
        t = (t + 1) >> 1;
        if (keyDelta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);

    h.first[i] = uint32_t(lo);
    h.last[i]  = uint32_t(hi);
    if (lo == gamma) {
        h.left[i]            = kLeafBit | uint32_t(gamma);
        h.leaf_parent[gamma] = i;
    } else {
        h.left[i]       = uint32_t(gamma);
        h.parent[gamma] = i;
    }
    if (hi == gamma + 1) {
        h.right[i]               = kLeafBit | uint32_t(gamma + 1);
        h.leaf_parent[gamma + 1] = i | kRightChild;
    } else {
        h.right[i]          = uint32_t(gamma + 1);
        h.parent[gamma + 1] = i | kRightChild;
    }
    if (0 == i) h.parent[0] = -1;
}

