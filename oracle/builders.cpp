// ORACLE — test infrastructure only. The oracle's own restatement of the reference's scene-compile builders, written from the
// Zig sources and independently of zyg_b200/csrc/host: tests/test_builders_host.py demands that what this file builds and what
// the product's host builds are identical byte for byte, so an error in either restatement shows ("host compile is
// in the product and nowhere else"), and bench.py --impl reference builds its tree here so the CPU arm never maps
// libzyg_b200.so.
//
//   SAH / spatial-split binary BVH     src/core/scene/bvh/builder_base.zig:65-390, split_candidate.zig:80-197
//   triangle tree (BLAS)               src/core/scene/shape/triangle/triangle_tree_builder.zig:33-65, 112-135, 166-207,
//                                      triangle_data.zig:40-60, vertex_buffer.zig:215-245, shape_provider.zig:863-924
//   prop tree (TLAS)                   src/core/scene/prop/prop_tree_builder.zig:24-96
//   light tree (scene + per part)      src/core/scene/light/light_tree_builder.zig:23-821, base/math/cone.zig:8-44
//   mesh-light sampling tables         src/core/scene/shape/triangle/triangle_mesh.zig:57-149, 160-230, 705-746,
//                                      shape_sampler.zig:149-226, base/math/distribution_1d.zig:87-124
//
// PARITY UNPINNED: the reference ships no tree fixtures and cannot be built here (Zig); the pins are structural (the O(N)
// brute-force intersection equals the traversal of these trees, leaves partition the references) plus the agreement of the
// two restatements.
#include "zbuild.hpp"
#include "zscene.hpp"
#include "zyg_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>

namespace zo {

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// builder_base.zig / split_candidate.zig
// ---------------------------------------------------------------------------------------------------------------------

constexpr uint32_t kParallelizeThreshold = 1024;  // builder_base.zig:16

struct Plane {  // SplitCandidate, split_candidate.zig:80-92
    Box      sides[2];
    uint32_t counts[2];
    float    d, cost;
    uint8_t  axis;
    bool     spatial;
};

inline Plane plane(uint8_t axis, Vec4f p, bool spatial) {  // SplitCandidate.init
    Plane c;
    c.d       = p[int(axis)];
    c.axis    = axis;
    c.spatial = spatial;
    c.cost    = 0.f;
    return c;
}

// SplitCandidate.evaluate, split_candidate.zig:97-160
void evaluatePlane(Plane& c, const std::vector<BRef>& refs, float area) {
    uint32_t cnt[2] = {0, 0};
    Box      bx[2]  = {Box::none(), Box::none()};
    const int ax    = int(c.axis);

    if (c.spatial) {
        bool straddled = false;
        for (const BRef& r : refs) {
            const Box b = r.box();
            if (b.hi[ax] < c.d) {  // behind(bounds[1])
                cnt[0] += 1;
                bx[0].absorb(b);
            } else if (!(b.lo[ax] < c.d)) {
                cnt[1] += 1;
                bx[1].absorb(b);
            } else {
                cnt[0] += 1;
                cnt[1] += 1;
                bx[0].absorb(b);
                bx[1].absorb(b);
                straddled = true;
            }
        }
        if (straddled) {
            bx[0].clipHi(c.d, c.axis);
            bx[1].clipLo(c.d, c.axis);
        } else {
            c.spatial = false;
        }
    } else {
        for (const BRef& r : refs) {
            const Box b = r.box();
            if (b.hi[ax] < c.d) {
                cnt[0] += 1;
                bx[0].absorb(b);
            } else {
                cnt[1] += 1;
                bx[1].absorb(b);
            }
        }
    }

    const size_t n = refs.size();
    if (0 == cnt[0] || 0 == cnt[1]) {
        c.cost = 2.f + float(n);
    } else {
        const float w0      = float(cnt[0]) * bx[0].surfaceArea();
        const float w1      = float(cnt[1]) * bx[1].surfaceArea();
        const float penalty = 0.125f * float(size_t(cnt[0]) + cnt[1] - n);
        c.cost              = 2.f + (w0 + w1) / area + penalty;
    }
    c.counts[0] = cnt[0];
    c.counts[1] = cnt[1];
    c.sides[0]  = bx[0];
    c.sides[1]  = bx[1];
}

// SplitCandidate.distribute, split_candidate.zig:162-192
void distribute(const Plane& c, const std::vector<BRef>& refs, std::vector<BRef>& left, std::vector<BRef>& right) {
    left.reserve(c.counts[0]);
    right.reserve(c.counts[1]);
    const int ax = int(c.axis);
    for (const BRef& r : refs) {
        if (r.mx[ax] < c.d) {
            left.push_back(r);
        } else if (!c.spatial || !(r.mn[ax] < c.d)) {
            right.push_back(r);
        } else {
            BRef a = r, b = r;
            a.mx[ax] = zo::min(c.d, r.mx[ax]);  // clippedMax, :67-73
            b.mn[ax] = zo::max(c.d, r.mn[ax]);  // clippedMin, :59-65
            left.push_back(a);
            right.push_back(b);
        }
    }
}

struct Settings {
    uint32_t num_slices, sweep_threshold, max_primitives, spatial_split_threshold, parallel_build_depth;
};

struct PendingTask {  // Task, builder_base.zig:18-27
    uint32_t          root, depth;
    Box               box;
    std::vector<BRef> refs;
};

struct Builder {  // Kernel, builder_base.zig:31-318
    Settings              s;
    std::vector<BNode>    nodes;
    std::vector<uint32_t> ids;
    uint32_t              unsplittable = 0;

    void begin(uint32_t num_primitives, const Settings& settings) {  // reserve, :305-317
        s = settings;
        nodes.clear();
        nodes.reserve(std::max<size_t>(size_t(3) * num_primitives / s.max_primitives, 1));
        nodes.push_back(BNode{});
        ids.clear();
        ids.reserve(size_t(num_primitives) * 12 / 10);
    }

    void leaf(uint32_t node, const std::vector<BRef>& refs) {  // assign, :295-303
        nodes[node].a = uint32_t(ids.size());
        nodes[node].n = uint32_t(refs.size());
        for (const BRef& r : refs) ids.push_back(r.prim);
    }

    // splittingPlane, :165-283. Returns false for "no plane" (null).
    bool choose(const std::vector<BRef>& refs, const Box& box, uint32_t depth, Plane& best) const {
        const float area = box.surfaceArea();
        if (0.f == area) return false;

        std::vector<Plane> cands;
        const Vec4f        centre = box.position();
        cands.push_back(plane(0, centre, true));
        cands.push_back(plane(1, centre, true));
        cands.push_back(plane(2, centre, true));

        if (refs.size() <= s.sweep_threshold) {
            for (const BRef& r : refs) {
                const Vec4f top = {{r.mx[0], r.mx[1], r.mx[2], 0.f}};
                for (uint8_t a = 0; a < 3; ++a) cands.push_back(plane(a, top, false));
            }
        } else {
            const Vec4f    ext     = box.extent();
            const uint32_t longest = indexMaxComponent3(ext);
            const float    step    = ext[int(longest)] / float(s.num_slices);
            for (uint8_t a = 0; a < 3; ++a) {
                const float    ea    = ext[a];
                const uint32_t steps = std::max(1u, uint32_t(std::ceil(ea / step)));
                const float    sa    = ea / float(steps);
                for (uint32_t i = 1; i < steps; ++i) {
                    Vec4f p = centre;
                    p[a]    = box.lo[a] + float(i) * sa;
                    cands.push_back(plane(a, p, false));
                    if (depth < s.spatial_split_threshold) cands.push_back(plane(a, p, true));
                }
            }
        }

        for (Plane& c : cands) evaluatePlane(c, refs, area);

        size_t pick = 0;
        float  low  = cands[0].cost;
        for (size_t i = 1; i < cands.size(); ++i) {
            if (cands[i].cost < low) {
                pick = i;
                low  = cands[i].cost;
            }
        }
        const Plane&   p = cands[pick];
        const uint32_t n = uint32_t(refs.size());
        if ((p.sides[0].covers(box) && n == p.counts[0]) || (p.sides[1].covers(box) && n == p.counts[1])) return false;
        best = p;
        return true;
    }

    // split, :65-163. `tasks` non-null = main-thread phase (threads.running_parallel == false) with tasks.capacity > 0.
    void grow(uint32_t node, std::vector<BRef>&& refs, const Box& box, uint32_t depth, std::vector<PendingTask>* tasks) {
        nodes[node].setBox(box);
        const uint32_t n = uint32_t(refs.size());
        if (n <= s.max_primitives) {
            leaf(node, refs);
            return;
        }
        if (tasks && (n < kParallelizeThreshold || depth == s.parallel_build_depth)) {
            tasks->push_back({node, depth, box, std::move(refs)});
            return;
        }
        Plane p;
        if (!choose(refs, box, depth, p)) {
            if (n <= 0x2FF) {
                leaf(node, refs);
            } else {
                unsplittable += 1;  // the reference logs an error and leaves the node as it is (:158-160)
                leaf(node, refs);
            }
            return;
        }
        if (n <= 0xFF && float(n) <= p.cost) {
            leaf(node, refs);
            return;
        }
        std::vector<BRef> left, right;
        distribute(p, refs, left, right);
        if (n <= 0x2FF && (left.empty() || right.empty())) {
            leaf(node, refs);
            return;
        }
        const uint32_t child = uint32_t(nodes.size());
        nodes[node].a        = child;  // setSplitNode
        nodes[node].n        = 0;
        nodes.push_back(BNode{});
        nodes.push_back(BNode{});
        std::vector<BRef>().swap(refs);
        grow(child, std::move(left), p.sides[0].common(box), depth + 1, tasks);
        grow(child + 1, std::move(right), p.sides[1].common(box), depth + 1, tasks);
    }
};

}  // namespace

void binarySplit(std::vector<BRef>&& refs, const Box& bounds, uint32_t num_slices, uint32_t sweep_threshold, uint32_t max_primitives,
                 uint32_t threads, BinaryBuild& out) {
    // Base.split, builder_base.zig:323-352
    Settings s{num_slices, sweep_threshold, max_primitives, 0, 0};
    const uint32_t count        = uint32_t(refs.size());
    s.spatial_split_threshold   = uint32_t(std::round(std::log2(float(count)) / 2.f));
    s.parallel_build_depth      = std::min(s.spatial_split_threshold, 6u);
    const uint32_t num_tasks    = std::min(1u << s.parallel_build_depth, count / kParallelizeThreshold);

    Builder main;
    main.begin(count, s);
    std::vector<PendingTask> tasks;
    main.grow(0, std::move(refs), bounds, 0, num_tasks > 0 ? &tasks : nullptr);

    // workOnTasks, :354-390: every task is an independent kernel, appended in task order
    std::vector<Builder>  subs(tasks.size());
    std::atomic<uint32_t> next{0};
    auto                  worker = [&] {
        for (;;) {
            const uint32_t i = next.fetch_add(1);
            if (i >= tasks.size()) return;
            subs[i].begin(uint32_t(tasks[i].refs.size()), s);
            subs[i].grow(0, std::move(tasks[i].refs), tasks[i].box, tasks[i].depth, nullptr);
        }
    };
    const uint32_t nt = std::max(1u, std::min<uint32_t>(threads, uint32_t(tasks.size())));
    if (nt <= 1) {
        worker();
    } else {
        std::vector<std::thread> pool;
        for (uint32_t t = 0; t < nt; ++t) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }

    for (size_t i = 0; i < tasks.size(); ++i) {
        const std::vector<BNode>& sub = subs[i].nodes;
        main.unsplittable += subs[i].unsplittable;
        main.nodes[tasks[i].root] = sub[0];
        if (1 == sub.size()) {
            // :368-370 `continue`s here: the leaf keeps an offset into the task's own id list, which is never appended. Counted;
            // the ids are appended so the leaf stays meaningful (zero occurrences on every mesh of the test suite).
            out.task_root_leaves += 1;
            const uint32_t ref_offset = uint32_t(main.ids.size());
            main.ids.insert(main.ids.end(), subs[i].ids.begin(), subs[i].ids.end());
            main.nodes[tasks[i].root].a += ref_offset;
            continue;
        }
        const uint32_t node_offset = uint32_t(main.nodes.size() - 1);
        const uint32_t ref_offset  = uint32_t(main.ids.size());
        main.ids.insert(main.ids.end(), subs[i].ids.begin(), subs[i].ids.end());
        main.nodes[tasks[i].root].a += node_offset;  // parent.offset(node_offset)
        for (size_t c = 1; c < sub.size(); ++c) {
            BNode nd = sub[c];
            nd.a += 0 == nd.n ? node_offset : ref_offset;  // Node.initFrom
            main.nodes.push_back(nd);
        }
    }

    out.nodes        = std::move(main.nodes);
    out.ids          = std::move(main.ids);
    out.unsplittable = main.unsplittable;
}

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// triangle tree
// ---------------------------------------------------------------------------------------------------------------------

struct MeshBuild {
    std::vector<BNode>    nodes;      // serialised: children adjacent, first child's subtree first
    std::vector<uint32_t> triangles;  // 3 per tree-order triangle
    std::vector<uint16_t> parts;
    std::vector<uint32_t> original;
    std::vector<float>    positions;  // 3 per vertex + 1
    std::vector<uint16_t> normals;    // 2 per vertex
    std::vector<float>    uvs;        // 2 per vertex
    uint32_t              leaf_offset_mismatches = 0;
    uint32_t              unsplittable           = 0;
    uint32_t              task_root_leaves       = 0;
};

struct SrcTriangle {  // Builder.IndexTriangle, triangle_tree_builder.zig:18-21
    uint32_t i[3], part;
};

// enc.compressNormal = floatToSnorm16(octEncode(n)), encoding.zig:60-68, 81-86, 100-103
void compressNormal(Vec4f n, uint16_t out[2]) {
    const float inorm = 1.f / (std::fabs(n[0]) + std::fabs(n[1]) + std::fabs(n[2]));
    const float t     = zo::max(n[2], 0.f);
    for (int k = 0; k < 2; ++k) {
        const float o = (n[k] + (n[k] > 0.f ? t : -t)) * inorm;
        out[k]        = uint16_t((o + 1.f) * (o > 0.f ? 32767.5f : 32768.f));
    }
}

// Builder.serialize, triangle_tree_builder.zig:166-207: recursion replaced by an explicit stack, same visiting order.
void serializeTriangles(const BinaryBuild& b, const std::vector<SrcTriangle>& src, MeshBuild& m) {
    m.nodes.assign(b.nodes.size(), BNode{});
    m.triangles.assign(b.ids.size() * 3, 0);
    m.parts.assign(b.ids.size(), 0);
    m.original.assign(b.ids.size(), 0);

    uint32_t next_node = 1;  // super.newNode() before the first call (:63)
    uint32_t next_tri  = 0;
    std::vector<std::pair<uint32_t, uint32_t>> todo{{0u, 0u}};
    while (!todo.empty()) {
        const auto [from, to] = todo.back();
        todo.pop_back();
        BNode nd = b.nodes[from];
        if (0 == nd.n) {
            const uint32_t kids = nd.a;
            nd.a                = next_node;
            m.nodes[to]         = nd;
            todo.push_back({kids + 1, next_node + 1});
            todo.push_back({kids, next_node});
            next_node += 2;
        } else {
            // the reference stores the builder's id offset in the leaf and writes the triangles at the running counter; the two
            // are the same number whenever leaves were emitted in depth-first order. The running counter is stored.
            if (nd.a != next_tri) m.leaf_offset_mismatches += 1;
            const uint32_t first = nd.a;
            nd.a                 = next_tri;
            m.nodes[to]          = nd;
            for (uint32_t p = first; p < first + nd.n; ++p, ++next_tri) {
                const SrcTriangle& t        = src[b.ids[p]];
                m.triangles[next_tri * 3]     = t.i[0];
                m.triangles[next_tri * 3 + 1] = t.i[1];
                m.triangles[next_tri * 3 + 2] = t.i[2];
                m.parts[next_tri]             = uint16_t(t.part);  // @truncate, triangle_data.zig:59
                m.original[next_tri]          = b.ids[p];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// prop tree
// ---------------------------------------------------------------------------------------------------------------------

struct PropTreeBuild {
    std::vector<BNode>    nodes;
    std::vector<uint32_t> indices;
};

// Builder.serialize, prop_tree_builder.zig:63-96
void serializeProps(const BinaryBuild& b, PropTreeBuild& out) {
    out.nodes.assign(b.nodes.size(), BNode{});
    out.indices.assign(b.ids.size(), 0);
    uint32_t next_node = 1, next_prop = 0;
    std::vector<std::pair<uint32_t, uint32_t>> todo{{0u, 0u}};
    while (!todo.empty()) {
        const auto [from, to] = todo.back();
        todo.pop_back();
        BNode nd = b.nodes[from];
        if (0 == nd.n) {
            const uint32_t kids = nd.a;
            nd.a                = next_node;
            out.nodes[to]       = nd;
            todo.push_back({kids + 1, next_node + 1});
            todo.push_back({kids, next_node});
            next_node += 2;
        } else {
            const uint32_t first = nd.a;
            nd.a                 = next_prop;  // setLeafNode(i, num)
            out.nodes[to]        = nd;
            std::copy(b.ids.begin() + first, b.ids.begin() + first + nd.n, out.indices.begin() + next_prop);
            next_prop += nd.n;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// light trees
// ---------------------------------------------------------------------------------------------------------------------

constexpr uint32_t kSceneSweep = 128, kPartSweep = 32, kLightSlices = 16;  // light_tree_builder.zig:23-25
constexpr uint32_t kMaxSplitDepth = 10, kMaxLights = 64;                   // Tree.MaxSplitDepth / MaxLights, light_tree.zig:248-249
constexpr float    kPiF = 3.14159265358979323846f;

// math.cone.merge, cone.zig:8-44
Vec4f mergeCones(Vec4f a, Vec4f b) {
    if (1.f == a[0] && 1.f == a[1] && 1.f == a[2] && 1.f == a[3]) return b;
    if (a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3]) return a;

    float a_angle = std::acos(a[3]);
    float b_angle = std::acos(b[3]);
    if (b_angle > a_angle) {
        std::swap(a, b);
        std::swap(a_angle, b_angle);
    }
    const float d_angle = std::acos(clamp(dot3(a, b), -1.f, 1.f));
    if (zo::min(d_angle + b_angle, kPiF) <= a_angle) return a;

    const float o_angle = (a_angle + d_angle + b_angle) / 2.f;
    if (o_angle >= kPiF) return {{a[0], a[1], a[2], -1.f}};

    const float r_angle = o_angle - a_angle;
    // Mat3x3.initRotation(normalize3(cross3(a, b)), r_angle), matrix3x3.zig:50-77
    const Vec4f v = normalize3(cross3(a, b));
    const float c = std::cos(r_angle), s = std::sin(r_angle), t = 1.f - c;
    const float at0 = v[0] * v[1] * t, at1 = v[2] * s;
    const float bt0 = v[0] * v[2] * t, bt1 = v[1] * s;
    const float ct0 = v[1] * v[2] * t, ct1 = v[0] * s;
    const Vec4f rows[3] = {{{c + v[0] * v[1] * t, at0 - at1, bt0 + bt1, 0.f}},
                           {{at0 + at1, c + v[1] * v[1] * t, ct0 - ct1, 0.f}},
                           {{bt0 - bt1, ct0 + ct1, c + v[2] * v[2] * t, 0.f}}};
    // rot.transformVector(a), :113-127
    Vec4f r = splat(a[0]) * rows[0];
    r       = mulAdd(splat(a[1]), rows[1], r);
    r       = mulAdd(splat(a[2]), rows[2], r);
    const Vec4f axis = normalize3(r);
    return {{axis[0], axis[1], axis[2], std::cos(o_angle)}};
}

float coneCost(float cos_a, bool two_sided) {  // light_tree_builder.zig:811-821
    const float o   = two_sided ? kPiF : std::acos(cos_a);
    const float w   = zo::min(o + (kPiF / 2.f), kPiF);
    const float sin = std::sin(o);
    const float b   = (kPiF / 2.f) * (2.f * w * sin - std::cos(o - 2.f * w) - 2.f * o * sin + cos_a);
    return (2.f * kPiF) * (1.f - cos_a) + b;
}

// The "set" a tree is built over: scene lights (Scene.lightAabb / lightCone / lightPower / lightTwoSided, scene.zig:650-664)
// or the emitting triangles of one mesh part (MeshImpl.lightAabb / lightCone / lightPower, shape_sampler.zig:183-199).
struct LightSetView {
    const Box*     boxes;
    const Vec4f*   cones;
    const float*   powers;
    const uint8_t* two_sided;  // scene lights only
    bool           part;       // evaluateSampler instead of evaluateScene
    bool           part_two_sided;
};

struct LightCut {  // SplitCandidate, light_tree_builder.zig:58-262
    enum Kind { Axis, Angle, Partition } kind;
    float    d;
    uint32_t axis;
    Vec4f    normal;
    uint32_t num_left, left[2];

    Box   boxes[2];
    Vec4f cones[2];
    float powers[2];
    float cost;
    bool  two_sided[2];
    bool  exhausted;

    bool onLeft(uint32_t l, const LightSetView& set) const {  // leftSide, :100-113
        switch (kind) {
            case Axis: return set.boxes[l].hi[int(axis)] < d;
            case Angle: return dot3(normal, set.cones[l]) < 0.f;
            default:
                for (uint32_t i = 0; i < num_left; ++i) {
                    if (l == left[i]) return true;
                }
                return false;
        }
    }
    float regularized(Vec4f extent) const {  // :115-122
        const float maxe = hmax3(extent);
        return Axis == kind ? maxe / extent[int(axis)] : maxe / hmin3(extent);
    }

    void finish(uint32_t num_lights, const uint32_t sides[2], const Box& bounds, float cone_weight) {
        const Vec4f extent = bounds.extent();
        if (0 == sides[0] || 0 == sides[1]) {
            const float reg = hmax3(extent) / hmin3(extent);
            cost            = float(num_lights) * reg * (powers[0] + powers[1]);
            exhausted       = true;
        } else {
            const float area = bounds.surfaceArea();
            const float reg  = regularized(extent);
            const float wa   = coneCost(cones[0][3], two_sided[0]);
            const float wb   = coneCost(cones[1][3], two_sided[1]);
            const float aa   = boxes[0].surfaceArea();
            const float ab   = boxes[1].surfaceArea();
            cost             = reg * (((powers[0] * wa * aa) + (powers[1] * wb * ab)) / (area * cone_weight));
            exhausted        = false;
        }
    }

    void evaluate(const uint32_t* lights, uint32_t n, const Box& bounds, float cone_weight, const LightSetView& set) {
        uint32_t sides[2] = {0, 0};
        Box      bx[2]    = {Box::none(), Box::none()};
        float    pw[2]    = {0.f, 0.f};
        if (!set.part) {  // evaluateScene, :131-185
            Vec4f cn[2] = {splat(1.f), splat(1.f)};
            bool  ts[2] = {false, false};
            for (uint32_t k = 0; k < n; ++k) {
                const uint32_t l = lights[k];
                const float    p = set.powers[l];
                if (0.f == p) continue;
                const uint32_t side = onLeft(l, set) ? 0 : 1;
                sides[side] += 1;
                bx[side].absorb(set.boxes[l]);
                cn[side] = mergeCones(cn[side], set.cones[l]);
                ts[side] = ts[side] || 0 != set.two_sided[l];
                pw[side] += p;
            }
            for (int i = 0; i < 2; ++i) {
                boxes[i] = bx[i], cones[i] = cn[i], powers[i] = pw[i], two_sided[i] = ts[i];
            }
        } else {  // evaluateSampler, :187-262
            Vec4f dominant[2] = {splat(0.f), splat(0.f)};
            for (uint32_t k = 0; k < n; ++k) {
                const uint32_t l = lights[k];
                const float    p = set.powers[l];
                if (0.f == p) continue;
                const uint32_t side = onLeft(l, set) ? 0 : 1;
                sides[side] += 1;
                bx[side].absorb(set.boxes[l]);
                dominant[side] = dominant[side] + splat(p) * set.cones[l];
                pw[side] += p;
            }
            dominant[0] = normalize3(dominant[0] / splat(pw[0]));
            dominant[1] = normalize3(dominant[1] / splat(pw[1]));
            float angles[2] = {0.f, 0.f};
            for (uint32_t k = 0; k < n; ++k) {
                const uint32_t l = lights[k];
                if (0.f == set.powers[l]) continue;
                const uint32_t side = onLeft(l, set) ? 0 : 1;
                const float    c    = clamp(dot3(dominant[side], set.cones[l]), -1.f, 1.f);
                angles[side]        = zo::max(angles[side], std::acos(c));
            }
            for (int i = 0; i < 2; ++i) {
                boxes[i]     = bx[i];
                cones[i]     = {{dominant[i][0], dominant[i][1], dominant[i][2], std::cos(angles[i])}};
                powers[i]    = pw[i];
                two_sided[i] = set.part_two_sided;
            }
        }
        finish(n, sides, bounds, cone_weight);
    }
};

struct LightBuildNode {  // BuildNode, light_tree_builder.zig:27-56
    Box      bounds;
    Vec4f    cone;
    float    power, variance;
    uint32_t middle, children_or_light, num_lights;
    bool     two_sided;
};

struct LightTreeBuild {
    std::vector<ZygpuLightNode> nodes;
    std::vector<uint32_t>       middles, orders, mapping;
    std::vector<float>          infinite_cdf;
    Box                         bounds{splat(0.f), splat(0.f)};
    float                       infinite_weight = 0.f, infinite_guard = 0.f;
    uint32_t                    infinite_end = 0, max_split_depth = kMaxSplitDepth, num_infinite = 0;
};

struct LightTreeBuilder {
    const LightSetView&         set;
    std::vector<LightBuildNode> nodes;
    std::vector<uint32_t>&      mapping;
    std::vector<uint32_t>&      orders;
    uint32_t                    next_node = 1, next_order = 0;

    float variance(const uint32_t* lights, uint32_t n) const {  // :646-662
        float    ap = 0.f, aps = 0.f;
        uint32_t k = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const float p = set.powers[lights[i]];
            if (p > 0.f) {
                k += 1;
                const float in = 1.f / float(k);
                ap += (p - ap) * in;
                aps += (p * p - aps) * in;
            }
        }
        return std::fabs(aps - ap * ap);
    }

    // evaluateSplits, :664-788
    LightCut best(const uint32_t* lights, uint32_t n, const Box& bounds, Vec4f cone, bool two_sided, uint32_t sweep) const {
        std::vector<LightCut> cuts;
        auto partition = [&](std::initializer_list<uint32_t> left) {
            LightCut c{};
            c.kind     = LightCut::Partition;
            c.num_left = uint32_t(left.size());
            uint32_t i = 0;
            for (uint32_t l : left) c.left[i++] = l;
            cuts.push_back(c);
        };
        auto axis = [&](Vec4f p, uint32_t a) {
            LightCut c{};
            c.kind = LightCut::Axis;
            c.d    = p[int(a)];
            c.axis = a;
            cuts.push_back(c);
        };
        if (2 == n) {
            partition({lights[0]});
        } else if (3 == n) {
            partition({lights[0]});
            partition({lights[1]});
            partition({lights[2]});
        } else if (4 == n) {
            partition({lights[0]});
            partition({lights[1]});
            partition({lights[2]});
            partition({lights[3]});
            partition({lights[0], lights[1]});
            partition({lights[0], lights[2]});
            partition({lights[0], lights[3]});
        } else {
            if (n <= sweep) {
                for (uint32_t k = 0; k < n; ++k) {
                    const Vec4f top = set.boxes[lights[k]].hi;
                    axis(top, 0);
                    axis(top, 1);
                    axis(top, 2);
                }
            } else {
                const Vec4f    centre  = bounds.position();
                const Vec4f    ext     = bounds.extent();
                const uint32_t longest = indexMaxComponent3(ext);
                const float    step    = ext[int(longest)] / float(kLightSlices);
                for (uint32_t a = 0; a < 3; ++a) {
                    const float    ea    = ext[int(a)];
                    const uint32_t steps = uint32_t(std::ceil(ea / step));
                    const float    sa    = ea / float(steps);
                    for (uint32_t i = 1; i < steps; ++i) {
                        Vec4f p   = centre;
                        p[int(a)] = bounds.lo[int(a)] + float(i) * sa;
                        axis(p, a);
                    }
                }
            }
            // :741-746 configures the same slot three times (two tangents, then the cone axis): the cone axis is what stays
            LightCut c{};
            c.kind   = LightCut::Angle;
            c.normal = cone;
            cuts.push_back(c);
        }
        const float weight = coneCost(cone[3], two_sided);
        for (LightCut& c : cuts) c.evaluate(lights, n, bounds, weight, set);
        size_t pick = 0;
        float  low  = cuts[0].cost;
        for (size_t i = 1; i < cuts.size(); ++i) {
            if (cuts[i].cost < low) {
                pick = i;
                low  = cuts[i].cost;
            }
        }
        return cuts[pick];
    }

    uint32_t close(uint32_t node, uint32_t begin, uint32_t end, const Box& bounds, Vec4f cone, float power) {  // assign(Primitive), :553-617
        bool any_two_sided = false;
        for (uint32_t i = begin; i < end; ++i) {
            const uint32_t l = mapping[i];
            orders[l]        = next_order++;
            if (!set.part) any_two_sided = any_two_sided || 0 != set.two_sided[l];
        }
        LightBuildNode& nd   = nodes[node];
        nd.bounds            = bounds;
        nd.cone              = cone;
        nd.power             = power;
        nd.variance          = variance(mapping.data() + begin, end - begin);
        nd.middle            = 0;
        nd.children_or_light = begin;
        nd.num_lights        = end - begin;
        nd.two_sided         = set.part ? set.part_two_sided : any_two_sided;
        return end;
    }

    // split / splitPrimitive, :430-538
    uint32_t grow(uint32_t node, uint32_t begin, uint32_t end, const Box& bounds, Vec4f cone, bool two_sided, float power, uint32_t depth) {
        const uint32_t len = end - begin;
        const bool     stop = set.part ? len <= 4 : (1 == len || (2 == len && depth > kMaxSplitDepth));
        if (stop) return close(node, begin, end, bounds, cone, power);

        const uint32_t child = next_node;
        const LightCut cut   = best(mapping.data() + begin, len, bounds, cone, two_sided, set.part ? kPartSweep : kSceneSweep);
        if (cut.exhausted) return close(node, begin, end, bounds, cone, power);

        // base.memory.partition, memory/partition.zig:3-29
        uint32_t* data  = mapping.data() + begin;
        uint32_t  first = len;
        for (uint32_t i = 0; i < len; ++i) {
            if (!cut.onLeft(data[i], set)) {
                first = i;
                break;
            }
        }
        if (first != len) {
            for (uint32_t i = first + 1; i < len; ++i) {
                if (cut.onLeft(data[i], set)) {
                    std::swap(data[i], data[first]);
                    first += 1;
                }
            }
        }
        const uint32_t middle = begin + first;

        next_node += 2;
        const uint32_t c0_end = grow(child, begin, middle, cut.boxes[0], cut.cones[0], cut.two_sided[0], cut.powers[0], depth + 1);
        const uint32_t c1_end = grow(child + 1, middle, end, cut.boxes[1], cut.cones[1], cut.two_sided[1], cut.powers[1], depth + 1);

        LightBuildNode& nd   = nodes[node];
        nd.bounds            = bounds;
        nd.cone              = cone;
        nd.power             = power;
        nd.variance          = variance(mapping.data() + begin, len);
        nd.middle            = c0_end;
        nd.children_or_light = child;
        nd.num_lights        = len;
        nd.two_sided         = two_sided;
        return c1_end;
    }

    void serialize(LightTreeBuild& out) {  // :619-644
        nodes[0].bounds.cacheRadius();
        const Box total = nodes[0].bounds;
        out.nodes.assign(next_node, ZygpuLightNode{});
        out.middles.assign(next_node, 0);
        for (uint32_t i = 0; i < next_node; ++i) {
            const LightBuildNode& src = nodes[i];
            ZygpuLightNode&       dst = out.nodes[i];
            const Vec4f           p   = src.bounds.position();
            const Vec4f centre = {{p[0], p[1], p[2], 0.5f * length3(src.bounds.extent())}};
            // Node.compressCenter, light_tree.zig:43-56
            const Vec4f d   = centre - total.lo;
            const Vec4f e   = total.extent();
            const Vec4f div = {{0.f == e[0] ? 1.f : e[0], 0.f == e[1] ? 1.f : e[1], 0.f == e[2] ? 1.f : e[2], total.hi[3]}};
            const Vec4f q   = d / div;
            for (int k = 0; k < 4; ++k) {
                dst.center[k] = uint16_t(std::fmaf(q[k], 65535.f, 0.5f));                                        // floatToUnorm16
                dst.cone[k]   = uint16_t((src.cone[k] + 1.f) * (src.cone[k] > 0.f ? 32767.5f : 32768.f));       // floatToSnorm16
            }
            dst.power      = src.power;
            dst.variance   = src.variance;
            dst.meta       = (src.middle > 0 ? 1u : 0u) | (src.two_sided ? 2u : 0u) | (src.children_or_light << 2);
            dst.num_lights = src.num_lights;
            out.middles[i] = src.middle;
        }
        out.bounds = total;
    }
};

// BuildNode.countPotentialLights, light_tree_builder.zig:42-55
void countPotential(const std::vector<LightBuildNode>& nodes, uint32_t node, uint32_t depth, uint32_t counts[][2]) {
    const LightBuildNode& nd = nodes[node];
    if (0 == nd.middle) {
        counts[depth][0] += 1;
    } else {
        counts[depth][1] += 2;
        if (depth + 1 < kMaxSplitDepth) {
            countPotential(nodes, nd.children_or_light, depth + 1, counts);
            countPotential(nodes, nd.children_or_light + 1, depth + 1, counts);
        }
    }
}

// Distribution1D.precomputePdfCdf, distribution_1d.zig:87-124. Returns the integral.
float precomputeCdf(const float* data, size_t n, std::vector<float>& cdf) {
    float integral = 0.f;
    for (size_t i = 0; i < n; ++i) integral += data[i];
    if (0.f == integral) {
        cdf = {1.f, 1.f};
        return 0.f;
    }
    cdf.assign(n + 1, 0.f);
    const float ii = 1.f / integral;
    float       p  = 0.f;
    for (size_t i = 0; i + 1 < n; ++i) {
        const float c = std::fmaf(data[i], ii, p);
        cdf[i + 1]    = c;
        p             = c;
    }
    cdf[n] = 1.f;
    return integral;
}

struct MeshSamplerBuild {
    std::vector<uint32_t> triangle_mapping, primitive_mapping;
    std::vector<float>    pdfs, part_areas;
    LightTreeBuild        tree;
    Box                   box;
    Vec4f                 cone;
    float                 power;
};

struct Handle {
    MeshBuild        mesh;
    PropTreeBuild    props;
    LightTreeBuild   lights;
    MeshSamplerBuild sampler;
};

Vec4f vertexPosition(const float* positions, uint32_t i) { return {{positions[3 * size_t(i)], positions[3 * size_t(i) + 1], positions[3 * size_t(i) + 2], 0.f}}; }

}  // namespace
}  // namespace zo

using namespace zo;

extern "C" {

void zo_build_free(void* handle) { delete static_cast<Handle*>(handle); }

void* zo_mesh_build(uint32_t num_parts, const uint32_t* parts, uint32_t num_triangles, const uint32_t* indices, uint32_t num_vertices,
                    const float* positions, uint32_t positions_stride, const float* normals, uint32_t normals_stride, const float* uvs,
                    uint32_t uvs_stride, uint32_t threads) {
    if (0 == threads) threads = std::max(1u, std::thread::hardware_concurrency());
    // Provider.buildDescAsync, shape_provider.zig:863-898: triangles are filled part by part
    std::vector<SrcTriangle> src(num_triangles, SrcTriangle{{0, 0, 0}, 0});
    const uint32_t           whole[3] = {0, num_triangles * 3, 0};
    const uint32_t           np       = num_parts > 0 && parts ? num_parts : 1;
    const uint32_t*          ps       = num_parts > 0 && parts ? parts : whole;
    for (uint32_t p = 0; p < np; ++p) {
        const uint32_t begin = ps[3 * p] / 3;
        const uint32_t end   = std::min((ps[3 * p] + ps[3 * p + 1]) / 3, num_triangles);
        for (uint32_t i = begin; i < end; ++i) {
            for (uint32_t k = 0; k < 3; ++k) src[i].i[k] = indices ? indices[3 * i + k] : 3 * i + k;
            src[i].part = p;
        }
    }

    // ReferencesContext.run, triangle_tree_builder.zig:112-135
    std::vector<BRef> refs(num_triangles);
    Box               bounds = Box::none();
    for (uint32_t r = 0; r < num_triangles; ++r) {
        auto at = [&](uint32_t v) -> Vec4f {
            const size_t id = size_t(v) * positions_stride;
            return {{positions[id], positions[id + 1], positions[id + 2], 0.f}};
        };
        const Vec4f a = at(src[r].i[0]), b = at(src[r].i[1]), c = at(src[r].i[2]);
        const Vec4f lo = min4(a, min4(b, c)), hi = max4(a, max4(b, c));  // triangle.zig:18-24
        refs[r] = {{lo[0], lo[1], lo[2]}, r, {hi[0], hi[1], hi[2]}, 0};
        bounds.lo = min4(bounds.lo, lo);
        bounds.hi = max4(bounds.hi, hi);
    }

    BinaryBuild build;
    binarySplit(std::move(refs), bounds, 16, 64, 4, threads, build);  // shape_provider.zig:922

    Handle*    h = new Handle;
    MeshBuild& m = h->mesh;
    serializeTriangles(build, src, m);
    m.unsplittable     = build.unsplittable;
    m.task_root_leaves = build.task_root_leaves;

    // Data.allocateTriangles + CAPI.copy, triangle_data.zig:40-55, vertex_buffer.zig:215-245
    m.positions.assign(size_t(num_vertices) * 3 + 1, 0.f);
    m.normals.assign(size_t(num_vertices) * 2, 0);
    m.uvs.assign(size_t(num_vertices) * 2, 0.f);
    for (uint32_t i = 0; i < num_vertices; ++i) {
        for (int k = 0; k < 3; ++k) m.positions[3 * size_t(i) + k] = positions[size_t(i) * positions_stride + k];
        const Vec4f n = normals ? Vec4f{{normals[size_t(i) * normals_stride], normals[size_t(i) * normals_stride + 1],
                                         normals[size_t(i) * normals_stride + 2], 0.f}}
                                : Vec4f{{0.f, 0.f, 1.f, 0.f}};
        compressNormal(n, &m.normals[2 * size_t(i)]);
        if (uvs) {
            m.uvs[2 * size_t(i)]     = uvs[size_t(i) * uvs_stride];
            m.uvs[2 * size_t(i) + 1] = uvs[size_t(i) * uvs_stride + 1];
        }
    }
    return h;
}

/* which: 0 nodes, 1 triangles, 2 original, 3 positions, 4 normals, 5 uvs, 6 parts (the ZYG_MESH_* numbering of include/zygpu.h);
 * 100 diagnostics {leaf offset mismatches, unsplittable nodes, task roots that stayed leaves}. */
const void* zo_mesh_data(const void* handle, int which, uint64_t* num_bytes) {
    const MeshBuild& m = static_cast<const Handle*>(handle)->mesh;
    static thread_local uint32_t diag[3];
    const void* p = nullptr;
    uint64_t    n = 0;
    switch (which) {
        case 0: p = m.nodes.data(), n = m.nodes.size() * sizeof(BNode); break;
        case 1: p = m.triangles.data(), n = m.triangles.size() * 4; break;
        case 2: p = m.original.data(), n = m.original.size() * 4; break;
        case 3: p = m.positions.data(), n = m.positions.size() * 4; break;
        case 4: p = m.normals.data(), n = m.normals.size() * 2; break;
        case 5: p = m.uvs.data(), n = m.uvs.size() * 4; break;
        case 6: p = m.parts.data(), n = m.parts.size() * 2; break;
        case 100:
            diag[0] = m.leaf_offset_mismatches, diag[1] = m.unsplittable, diag[2] = m.task_root_leaves;
            p = diag, n = sizeof(diag);
            break;
        default: break;
    }
    if (num_bytes) *num_bytes = n;
    return p;
}

/* PropBvhBuilder.build over `indices` (prop ids in the order Scene hands them over) and the world boxes of all props. */
void* zo_prop_tree_build(const uint32_t* indices, uint32_t num_indices, const ZygpuAabb* aabbs, uint32_t threads) {
    Handle* h = new Handle;
    if (0 == num_indices) return h;
    if (0 == threads) threads = std::max(1u, std::thread::hardware_concurrency());
    std::vector<BRef> refs(num_indices);
    Box               bounds = Box::none();
    for (uint32_t i = 0; i < num_indices; ++i) {
        const ZygpuAabb& b = aabbs[indices[i]];
        refs[i]            = {{b.min[0], b.min[1], b.min[2]}, indices[i], {b.max[0], b.max[1], b.max[2]}, 0};
        bounds.absorb({load4(b.min), load4(b.max)});
    }
    BinaryBuild build;
    binarySplit(std::move(refs), bounds, 16, 64, 4, threads, build);  // prop_tree_builder.zig:17
    serializeProps(build, h->props);
    return h;
}

/* which: 0 nodes, 1 indices */
const void* zo_prop_tree_data(const void* handle, int which, uint64_t* num_bytes) {
    const PropTreeBuild& t = static_cast<const Handle*>(handle)->props;
    const void*          p = 0 == which ? static_cast<const void*>(t.nodes.data()) : static_cast<const void*>(t.indices.data());
    if (num_bytes) *num_bytes = 0 == which ? t.nodes.size() * sizeof(BNode) : t.indices.size() * 4;
    return p;
}

/* Builder.build, light_tree_builder.zig:281-376, over the scene's lights: boxes with the power in min[3] and the cached radius in
 * max[3] (scene.zig:496), cones, two-sidedness, and whether the light's shape is finite. */
void* zo_light_tree_build(uint32_t num_lights, const ZygpuAabb* aabbs, const float* cones, const uint8_t* two_sided, const uint8_t* finite) {
    Handle*         h = new Handle;
    LightTreeBuild& t = h->lights;

    std::vector<Box>   boxes(num_lights);
    std::vector<Vec4f> cns(num_lights);
    std::vector<float> powers(num_lights);
    for (uint32_t l = 0; l < num_lights; ++l) {
        boxes[l]  = {load4(aabbs[l].min), load4(aabbs[l].max)};
        cns[l]    = load4(cones + 4 * size_t(l));
        powers[l] = aabbs[l].min[3];
    }
    const LightSetView set{boxes.data(), cns.data(), powers.data(), two_sided, false, false};

    t.mapping.clear();
    t.orders.assign(num_lights, 0);
    for (uint32_t l = 0; l < num_lights; ++l) {
        if (0 == finite[l]) t.mapping.push_back(l);
    }
    const uint32_t num_infinite = uint32_t(t.mapping.size());
    for (uint32_t l = 0; l < num_lights; ++l) {
        if (0 != finite[l]) t.mapping.push_back(l);
    }
    t.num_infinite = num_infinite;

    uint32_t           order = 0;
    float              infinite_power = 0.f;
    std::vector<float> infinite_powers(num_infinite);
    for (uint32_t i = 0; i < num_infinite; ++i) {
        const uint32_t l   = t.mapping[i];
        infinite_powers[i] = powers[l];
        t.orders[l]        = order++;
        infinite_power += powers[l];
    }
    t.infinite_end = order;
    if (num_infinite > 0) precomputeCdf(infinite_powers.data(), num_infinite, t.infinite_cdf);

    const uint32_t num_finite = num_lights - num_infinite;
    float          root_power = 0.f;
    if (num_finite > 0) {
        LightTreeBuilder b{set, std::vector<LightBuildNode>(2 * size_t(num_finite) - 1), t.mapping, t.orders};
        b.next_order = order;

        Box   bounds = Box::none();
        Vec4f cone   = splat(1.f);
        bool  ts     = false;
        float total  = 0.f;
        for (uint32_t i = num_infinite; i < num_lights; ++i) {
            const uint32_t l = t.mapping[i];
            bounds.absorb(boxes[l]);
            cone = mergeCones(cone, cns[l]);
            ts   = ts || 0 != two_sided[l];
            total += powers[l];
        }
        b.grow(0, num_infinite, num_lights, bounds, cone, ts, total, 0);
        b.serialize(t);
        root_power = b.nodes[0].power;

        uint32_t split_lights[kMaxSplitDepth][2] = {};
        countPotential(b.nodes, 0, 0, split_lights);
        uint32_t num_split = 0;
        for (uint32_t i = 0; i < kMaxSplitDepth; ++i) {
            num_split += split_lights[i][0];
            if (num_split + split_lights[i][1] > kMaxLights - num_infinite || 0 == split_lights[i][1]) {
                t.max_split_depth = i;
                break;
            }
        }
    }
    const float pt    = infinite_power + (0 == num_finite ? 0.f : root_power);
    t.infinite_weight = (0 == num_lights || 0.f == pt) ? 0.f : infinite_power / pt;
    t.infinite_guard  = 0 == num_finite ? (0 == num_infinite ? 0.f : 1.1f) : t.infinite_weight;
    return h;
}

/* which: 0 nodes, 1 middles, 2 orders, 3 mapping, 4 infinite cdf, 5 {bounds min[4], max[4], infinite_weight, infinite_guard},
 * 6 {infinite_end, max_split_depth, num_infinite}; `sampler` != 0 reads the primitive tree of a zo_mesh_sampler_build handle. */
const void* zo_light_tree_data(const void* handle, int sampler, int which, uint64_t* num_bytes) {
    const Handle*         h = static_cast<const Handle*>(handle);
    const LightTreeBuild& t = sampler ? h->sampler.tree : h->lights;
    static thread_local float    f[10];
    static thread_local uint32_t u[3];
    const void* p = nullptr;
    uint64_t    n = 0;
    switch (which) {
        case 0: p = t.nodes.data(), n = t.nodes.size() * sizeof(ZygpuLightNode); break;
        case 1: p = t.middles.data(), n = t.middles.size() * 4; break;
        case 2: p = t.orders.data(), n = t.orders.size() * 4; break;
        case 3: p = t.mapping.data(), n = t.mapping.size() * 4; break;
        case 4: p = t.infinite_cdf.data(), n = t.infinite_cdf.size() * 4; break;
        case 5:
            for (int i = 0; i < 4; ++i) f[i] = t.bounds.lo[i], f[4 + i] = t.bounds.hi[i];
            f[8] = t.infinite_weight, f[9] = t.infinite_guard;
            p = f, n = sizeof(f);
            break;
        case 6:
            u[0] = t.infinite_end, u[1] = t.max_split_depth, u[2] = t.num_infinite;
            p = u, n = sizeof(u);
            break;
        default: break;
    }
    if (num_bytes) *num_bytes = n;
    return p;
}

/* Mesh.prepareSampling + Part.configure (uniform emission: every triangle of the part emits) + Builder.buildPrimitive over the
 * reference-layout arrays of one compiled mesh: triangle_mesh.zig:57-149, 160-230, 705-746; light_tree_builder.zig:378-428. */
void* zo_mesh_sampler_build(const ZoMesh* mesh, uint32_t num_tree_triangles, uint32_t num_parts, uint32_t part, int two_sided) {
    Handle*           h = new Handle;
    MeshSamplerBuild& s = h->sampler;

    // prepareSampling: per-part running index of every tree triangle; calculateAreas
    s.primitive_mapping.assign(num_tree_triangles, 0);
    s.part_areas.assign(num_parts, 0.f);
    std::vector<uint32_t> counts(num_parts, 0);
    auto corner = [&](uint32_t t, uint32_t k) { return vertexPosition(mesh->positions, mesh->triangles[3 * size_t(t) + k]); };
    for (uint32_t t = 0; t < num_tree_triangles; ++t) {
        const uint32_t p       = mesh->parts[t];
        s.primitive_mapping[t] = counts[p]++;
        const Vec4f a = corner(t, 0), b = corner(t, 1), c = corner(t, 2);
        s.part_areas[p] += 0.5f * length3(cross3(b - a, c - a));  // triangle.area, triangle.zig:151-153
    }
    for (uint32_t t = 0; t < num_tree_triangles; ++t) {
        if (mesh->parts[t] == part) s.triangle_mapping.push_back(t);
    }
    const uint32_t num = uint32_t(s.triangle_mapping.size());

    // EvalContext.run over the whole range as one task (the reference sums per-thread partial results, so its last bits depend
    // on the thread count; one task is the single-threaded result)
    std::vector<float> powers(num);
    std::vector<Box>   boxes(num);
    std::vector<Vec4f> cones(num);
    Box   bb       = Box::none();
    Vec4f dominant = splat(0.f);
    float total    = 0.f;
    for (uint32_t i = 0; i < num; ++i) {
        const uint32_t t = s.triangle_mapping[i];
        const Vec4f a = corner(t, 0), b = corner(t, 1), c = corner(t, 2);
        const float area = 0.5f * length3(cross3(b - a, c - a));
        const Vec4f n    = normalize3(cross3(b - a, c - a));  // Data.normal, triangle_data.zig:140-149
        powers[i]        = area;
        boxes[i]         = {min4(a, min4(b, c)), max4(a, max4(b, c))};
        cones[i]         = {{n[0], n[1], n[2], 1.f}};
        if (area > 0.f) {
            dominant = dominant + splat(area) * n;
            bb.absorb(boxes[i]);
            total += area;
        }
    }
    if (dominant[0] == dominant[1] && dominant[1] == dominant[2]) {
        s.cone = {{0.f, 0.f, 1.f, -1.f}};
    } else {
        const Vec4f da    = normalize3(dominant / splat(total));
        float       angle = 0.f;
        for (uint32_t i = 0; i < num; ++i) angle = zo::max(angle, std::acos(dot3(da, cones[i])));
        s.cone = {{da[0], da[1], da[2], std::cos(angle)}};
    }
    s.box = bb;

    // distribution.configure(powers): lightPower(l) = pdfI(l) = cdf[l + 1] - cdf[l]
    std::vector<float> cdf;
    s.power = precomputeCdf(powers.data(), num, cdf);
    s.pdfs.assign(num, 0.f);
    for (uint32_t i = 0; i < num && cdf.size() == size_t(num) + 1; ++i) s.pdfs[i] = cdf[i + 1] - cdf[i];

    // Builder.buildPrimitive
    LightTreeBuild& t = s.tree;
    t.mapping.resize(num);
    t.orders.assign(num, 0);
    for (uint32_t i = 0; i < num; ++i) t.mapping[i] = i;
    const LightSetView set{boxes.data(), cones.data(), s.pdfs.data(), nullptr, true, 0 != two_sided};
    if (num > 0) {
        LightTreeBuilder b{set, std::vector<LightBuildNode>(2 * size_t(num) - 1), t.mapping, t.orders};
        b.grow(0, 0, num, s.box, s.cone, 0 != two_sided, s.power, 0);
        b.serialize(t);
    }
    return h;
}

/* which: 0 triangle_mapping, 1 triangle pdfs, 2 primitive_mapping, 3 part areas, 4 {box min[4], max[4], cone[4], power} */
const void* zo_mesh_sampler_data(const void* handle, int which, uint64_t* num_bytes) {
    const MeshSamplerBuild& s = static_cast<const Handle*>(handle)->sampler;
    static thread_local float f[13];
    const void* p = nullptr;
    uint64_t    n = 0;
    switch (which) {
        case 0: p = s.triangle_mapping.data(), n = s.triangle_mapping.size() * 4; break;
        case 1: p = s.pdfs.data(), n = s.pdfs.size() * 4; break;
        case 2: p = s.primitive_mapping.data(), n = s.primitive_mapping.size() * 4; break;
        case 3: p = s.part_areas.data(), n = s.part_areas.size() * 4; break;
        case 4:
            for (int i = 0; i < 4; ++i) f[i] = s.box.lo[i], f[4 + i] = s.box.hi[i], f[8 + i] = s.cone[i];
            f[12] = s.power;
            p = f, n = sizeof(f);
            break;
        default: break;
    }
    if (num_bytes) *num_bytes = n;
    return p;
}

}  // extern "C"
