// ORACLE — test infrastructure only (see zmath.hpp). CPU restatement of the reference's
// triangle-BVH queries over plain arrays in the reference's own layout:
//   nodes      32-byte bvh.Node                    src/core/scene/bvh/node.zig:9-20
//   triangles  u32[3] per tree-order triangle      src/core/scene/shape/triangle/triangle.zig:6-10
//   positions  f32[3] per vertex, tightly packed   src/core/scene/shape/triangle/triangle_data.zig:62-64
#include "zyg_oracle.h"
#include "ztree.hpp"

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace zo {

namespace {

template <typename F>
void parallelRanges(uint64_t n, uint32_t threads, F&& fn) {
    if (0 == threads) threads = std::max(1u, std::thread::hardware_concurrency());
    if (threads <= 1 || n < 4096) {
        fn(0, n);
        return;
    }
    const uint64_t           grain = 4096;
    std::atomic<uint64_t>    next{0};
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < threads; ++t) {
        pool.emplace_back([&] {
            for (;;) {
                const uint64_t b = next.fetch_add(grain);
                if (b >= n) return;
                fn(b, std::min(b + grain, n));
            }
        });
    }
    for (auto& t : pool) t.join();
}

inline Ray makeRay(const ZoRay& r) {
    return Ray::init({{r.origin[0], r.origin[1], r.origin[2], 0.f}}, {{r.direction[0], r.direction[1], r.direction[2], 0.f}},
                     r.min_t, r.max_t);
}

}  // namespace

}  // namespace zo

extern "C" {

void zo_trace_closest(const void* nodes, const uint32_t* triangles, const float* positions, const ZoRay* rays,
                      uint64_t n, ZoHit* out, uint32_t threads, uint64_t* visited_nodes, uint64_t* tested_tris) {
    const zo::Mesh        mesh{static_cast<const zo::Node*>(nodes), triangles, positions};
    std::atomic<uint64_t> vn{0}, tt{0};
    const bool            count = visited_nodes || tested_tris;
    zo::parallelRanges(n, threads, [&](uint64_t b, uint64_t e) {
        uint64_t lvn = 0, ltt = 0;
        for (uint64_t i = b; i < e; ++i) mesh.intersect(zo::makeRay(rays[i]), out[i], count ? &lvn : nullptr, count ? &ltt : nullptr);
        vn += lvn;
        tt += ltt;
    });
    if (visited_nodes) *visited_nodes = vn;
    if (tested_tris) *tested_tris = tt;
}

void zo_trace_any(const void* nodes, const uint32_t* triangles, const float* positions, const ZoRay* rays, uint64_t n,
                  uint32_t* out, uint32_t threads) {
    const zo::Mesh mesh{static_cast<const zo::Node*>(nodes), triangles, positions};
    zo::parallelRanges(n, threads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) out[i] = mesh.intersectP(zo::makeRay(rays[i])) ? 1u : 0u;
    });
}

// Every triangle in list order with the reference's accept rule (<= max_t, later equal hit wins):
// the tree-independent ground truth. Also reports how many triangles share the winning t.
void zo_brute_closest(const uint32_t* triangles, uint32_t num_triangles, const float* positions, const ZoRay* rays,
                      uint64_t n, ZoHit* out, uint32_t* num_ties, uint32_t threads) {
    const zo::Mesh mesh{nullptr, triangles, positions};
    zo::parallelRanges(n, threads, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            zo::Ray  ray = zo::makeRay(rays[i]);
            ZoHit    best{ray.max_t, 0.f, 0.f, 0xFFFFFFFFu};
            uint32_t ties = 0;
            for (uint32_t t = 0; t < num_triangles; ++t) {
                zo::Hit hit;
                if (mesh.intersectIndexed(ray, t, hit)) {
                    ties      = (0xFFFFFFFFu != best.primitive && hit.t == best.t) ? ties + 1 : 0;
                    ray.max_t = hit.t;
                    best      = {hit.t, hit.u, hit.v, t};
                }
            }
            out[i] = best;
            if (num_ties) num_ties[i] = ties;
        }
    });
}

}  // extern "C"
