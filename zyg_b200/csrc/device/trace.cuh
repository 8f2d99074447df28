// Ray / triangle-BVH traversal kernels for sm_100a.
//
// Two families over the same mesh:
//  * `Wide*`  — product path: 8-wide quantised nodes (host/wide_bvh.hpp), 64-byte triangle records,
//               per-lane traversal with a node-group / triangle-group stack.
//  * `Binary*` — order-exact restatement of TriangleTree.intersect / intersectP
//               (src/core/scene/shape/triangle/triangle_tree.zig:46-109, 197-242) over the uploaded
//               32-byte reference nodes; used to validate the wide path on the device itself.
// Both use the same bit-exact Moeller-Trumbore test (triangle.zig:26-52): this translation unit is
// compiled with -fmad=false and fuses only where the reference writes @mulAdd.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace zygpu {

struct RayIn {  // 32 bytes: include/zygpu.h ZygpuRay
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
};

struct HitOut {  // 16 bytes: include/zygpu.h ZygpuHit
    float    t, u, v;
    uint32_t primitive;
};

struct MeshDevice {
    const float4*   wide_nodes;    // 5 per node
    const float4*   wide_tris;     // 4 per record
    const float4*   binary_nodes;  // 2 per node
    const uint32_t* triangles;     // 3 per BVH-order triangle
    const float*    positions;     // 3 per vertex (+1 pad)
    uint32_t        num_wide_nodes;
    uint32_t        num_tris;
};

struct TraceCounters {  // instrumented builds only
    unsigned long long nodes;      // wide: 80-byte node fetches; binary: 32-byte node fetches
    unsigned long long triangles;  // triangle tests
    unsigned long long rays;
    unsigned long long max_stack;
};

enum TraceMode : int {
    kClosestWide   = 0,
    kAnyWide       = 1,
    kClosestBinary = 2,
    kAnyBinary     = 3,
};

// Launches on `stream`. `out` is HitOut[n] for closest modes, uint32_t[n] (1 = occluded) for any-hit.
// `counters` may be null; when set the instrumented variant runs (slower) and accumulates into it.
// `work_counter` is one device word the persistent wide kernel hands ray blocks out from (reset on
// `stream` before the launch); it must not be shared by launches that can run concurrently.
// `stack_scratch`: device memory for the traversal stacks of the ray-pool kernel, traceStackBytesPerBlock() per resident block
// (it launches at most as many blocks as fit); must not be shared by launches that can run concurrently. Null: lock-step kernel.
cudaError_t launchTrace(const MeshDevice& mesh, int mode, const RayIn* rays, void* out, uint32_t n,
                        TraceCounters* counters, uint32_t* work_counter, void* stack_scratch, size_t stack_scratch_bytes, cudaStream_t stream);
size_t      traceStackBytesPerBlock();

}  // namespace zygpu
