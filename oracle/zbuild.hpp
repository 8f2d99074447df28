// ORACLE — test infrastructure only (see zmath.hpp). Build-side value types shared by the oracle's own restatement of
// the reference's scene-compile builders (builders.cpp): the AABB operations of src/base/math/aabb.zig the builders use
// and the bvh.Node record (src/core/scene/bvh/node.zig:9-72). Written from the Zig sources, independently of
// zyg_b200/csrc/host — the point of this file is that two restatements have to agree byte for byte.
#pragma once

#include "zmath.hpp"

#include <cstdint>
#include <vector>

namespace zo {

struct Box {  // math.AABB, aabb.zig:10-24
    Vec4f lo, hi;

    static Box none() { return {splat(FLT_MAX), splat(-FLT_MAX)}; }  // AABB.empty, :13

    Vec4f position() const { return splat(0.5f) * (lo + hi); }  // :23-25
    Vec4f extent() const { return hi - lo; }                    // :31-33
    float surfaceArea() const {                                 // :35-38
        const Vec4f d = hi - lo;
        return 2.f * (d[0] * d[1] + d[0] * d[2] + d[1] * d[2]);
    }
    void absorb(const Box& o) {  // mergeAssign, :199-202
        lo = min4(lo, o.lo);
        hi = max4(hi, o.hi);
    }
    Box common(const Box& o) const { return {max4(lo, o.lo), min4(hi, o.hi)}; }  // intersection, :192-197
    void clipLo(float d, uint32_t axis) { lo[int(axis)] = zo::max(d, lo[int(axis)]); }  // clipMin, :204-217
    void clipHi(float d, uint32_t axis) { hi[int(axis)] = zo::min(d, hi[int(axis)]); }  // clipMax, :219-223
    bool covers(const Box& o) const {                                                   // :225-232
        return lo[0] <= o.lo[0] && lo[1] <= o.lo[1] && lo[2] <= o.lo[2] && hi[0] >= o.hi[0] && hi[1] >= o.hi[1] && hi[2] >= o.hi[2];
    }
    void cacheRadius() {  // :141-144
        lo[3] = 0.f;
        hi[3] = 0.5f * length3(extent());
    }
};

struct BNode {  // bvh.Node: min xyz | children or first index, max xyz | count (0 = inner)
    float    mn[3];
    uint32_t a;
    float    mx[3];
    uint32_t n;

    void setBox(const Box& b) {
        for (int i = 0; i < 3; ++i) {
            mn[i] = b.lo[i];
            mx[i] = b.hi[i];
        }
    }
};
static_assert(sizeof(BNode) == 32, "bvh.Node is 32 bytes (size_test.zig:44)");

struct BRef {  // split_candidate.zig:9-76
    float    mn[3];
    uint32_t prim;
    float    mx[3];
    uint32_t pad;

    Box box() const { return {{{mn[0], mn[1], mn[2], 0.f}}, {{mx[0], mx[1], mx[2], 0.f}}}; }
};

struct BinaryBuild {
    std::vector<BNode>    nodes;  // builder order
    std::vector<uint32_t> ids;    // Kernel.reference_ids
    uint32_t              unsplittable = 0;     // "Cannot split node further" events (builder_base.zig:158-160)
    uint32_t              task_root_leaves = 0; // tasks whose root stayed a leaf (workOnTasks `continue`, :368-370)
};

// Base.split + workOnTasks, builder_base.zig:323-390.
void binarySplit(std::vector<BRef>&& refs, const Box& bounds, uint32_t num_slices, uint32_t sweep_threshold, uint32_t max_primitives,
                 uint32_t threads, BinaryBuild& out);

}  // namespace zo
