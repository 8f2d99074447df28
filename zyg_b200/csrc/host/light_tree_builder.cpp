#include "light_tree_builder.hpp"

#include <algorithm>
#include <cmath>

namespace zyg {

namespace {

constexpr uint32_t kSceneSweepThreshold = 128;  // light_tree_builder.zig:23
constexpr uint32_t kPartSweepThreshold  = 32;   // :24
constexpr uint32_t kNumSlices           = 16;   // :25
constexpr float    kPi                  = 3.14159265358979323846f;

bool equal4(Vec4f a, Vec4f b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3]; }

float hmax3(Vec4f v) { return fmax_(v[0], fmax_(v[1], v[2])); }
float hmin3(Vec4f v) { return fmin_(v[0], fmin_(v[1], v[2])); }
float clampf(float x, float mi, float ma) { return fmin_(fmax_(x, mi), ma); }

// coneCost, light_tree_builder.zig:812-820
float coneCost(float cos, bool two_sided) {
    const float o   = two_sided ? kPi : std::acos(cos);
    const float w   = fmin_(o + (kPi / 2.f), kPi);
    const float sin = std::sin(o);
    const float b   = (kPi / 2.f) * (2.f * w * sin - std::cos(o - 2.f * w) - 2.f * o * sin + cos);
    return (2.f * kPi) * (1.f - cos) + b;
}

struct BuildNode {  // :27-58
    AABB     bounds;
    Vec4f    cone;
    float    power, variance;
    uint32_t middle, children_or_light, num_lights;
    bool     two_sided;

    bool hasChildren() const { return middle > 0; }
};

struct SplitCandidate {  // :60-263
    enum Kind { Axis, Angle, Partition } kind;
    float    d;
    uint32_t axis;
    Vec4f    n;
    uint32_t part_num;
    uint32_t part_left[2];

    AABB  aabbs[2];
    Vec4f cones[2];
    float powers[2];
    float cost;
    bool  two_sided[2];
    bool  exhausted;

    void configure(Vec4f p, uint32_t a) {
        kind = Axis;
        d    = p[int(a)];
        axis = a;
    }
    void configureAngle(Vec4f normal) {
        kind = Angle;
        n    = normal;
    }
    void configurePartition(const uint32_t* left, uint32_t num) {
        kind     = Partition;
        part_num = num;
        for (uint32_t i = 0; i < num; ++i) part_left[i] = left[i];
    }

    bool leftSide(uint32_t l, const LightSet& set) const {  // :103-117
        switch (kind) {
            case Axis: return set.aabbs[l].b[1][int(axis)] < d;
            case Angle: return dot3(n, set.cones[l]) < 0.f;
            default:
                for (uint32_t i = 0; i < part_num; ++i) {
                    if (l == part_left[i]) return true;
                }
                return false;
        }
    }

    float regularized(Vec4f extent) const {  // :119-126
        const float maxe = hmax3(extent);
        return Axis == kind ? maxe / extent[int(axis)] : maxe / hmin3(extent);
    }

    void finish(const uint32_t num_sides[2], uint32_t num_lights, const AABB& bounds, float cone_weight) {
        const Vec4f extent = bounds.extent();
        if (0 == num_sides[0] || 0 == num_sides[1]) {
            const float reg = hmax3(extent) / hmin3(extent);
            cost            = float(num_lights) * reg * (powers[0] + powers[1]);
            exhausted       = true;
        } else {
            const float surface_area   = bounds.surfaceArea();
            const float reg            = regularized(extent);
            const float cone_weight_a  = coneCost(cones[0][3], two_sided[0]);
            const float cone_weight_b  = coneCost(cones[1][3], two_sided[1]);
            const float surface_area_a = aabbs[0].surfaceArea();
            const float surface_area_b = aabbs[1].surfaceArea();
            cost = reg * (((powers[0] * cone_weight_a * surface_area_a) + (powers[1] * cone_weight_b * surface_area_b)) /
                          (surface_area * cone_weight));
            exhausted = false;
        }
    }

    // evaluateScene, :137-191
    void evaluateScene(const uint32_t* lights, uint32_t num, const AABB& bounds, float cone_weight, const LightSet& set) {
        uint32_t num_sides[2] = {0, 0};
        aabbs[0] = aabbs[1] = AABB::empty();
        cones[0] = cones[1] = splat(1.f);
        two_sided[0] = two_sided[1] = false;
        powers[0] = powers[1] = 0.f;

        for (uint32_t i = 0; i < num; ++i) {
            const uint32_t l     = lights[i];
            const float    power = set.powers[l];
            if (0.f == power) continue;

            const uint32_t side = leftSide(l, set) ? 0 : 1;
            num_sides[side] += 1;
            aabbs[side].mergeAssign(set.aabbs[l]);
            cones[side]     = coneMerge(cones[side], set.cones[l]);
            two_sided[side] = two_sided[side] || set.twoSided(l);
            powers[side] += power;
        }
        finish(num_sides, num, bounds, cone_weight);
    }

    // evaluateSampler, :193-263
    void evaluateSampler(const uint32_t* lights, uint32_t num, const AABB& bounds, float cone_weight, const LightSet& set) {
        uint32_t num_sides[2] = {0, 0};
        aabbs[0] = aabbs[1] = AABB::empty();
        Vec4f dominant_axis[2] = {splat(0.f), splat(0.f)};
        powers[0] = powers[1] = 0.f;

        for (uint32_t i = 0; i < num; ++i) {
            const uint32_t l     = lights[i];
            const float    power = set.powers[l];
            if (0.f == power) continue;

            const uint32_t side = leftSide(l, set) ? 0 : 1;
            num_sides[side] += 1;
            aabbs[side].mergeAssign(set.aabbs[l]);
            dominant_axis[side] = dominant_axis[side] + splat(power) * set.cones[l];
            powers[side] += power;
        }

        dominant_axis[0] = normalize3(dominant_axis[0] / splat(powers[0]));
        dominant_axis[1] = normalize3(dominant_axis[1] / splat(powers[1]));

        float angles[2] = {0.f, 0.f};
        for (uint32_t i = 0; i < num; ++i) {
            const uint32_t l = lights[i];
            if (0.f == set.powers[l]) continue;
            const uint32_t side = leftSide(l, set) ? 0 : 1;
            const float    c    = clampf(dot3(dominant_axis[side], set.cones[l]), -1.f, 1.f);
            angles[side]        = fmax_(angles[side], std::acos(c));
        }

        cones[0]     = {{dominant_axis[0][0], dominant_axis[0][1], dominant_axis[0][2], std::cos(angles[0])}};
        cones[1]     = {{dominant_axis[1][0], dominant_axis[1][1], dominant_axis[1][2], std::cos(angles[1])}};
        two_sided[0] = two_sided[1] = set.all_two_sided;
        finish(num_sides, num, bounds, cone_weight);
    }

    void evaluate(const uint32_t* lights, uint32_t num, const AABB& bounds, float cone_weight, const LightSet& set) {
        if (set.primitive) {
            evaluateSampler(lights, num, bounds, cone_weight, set);
        } else {
            evaluateScene(lights, num, bounds, cone_weight, set);
        }
    }
};

void orthonormalBasis(Vec4f n, Vec4f& t, Vec4f& b) {  // vector4.zig:98-110
    const float sign = std::copysign(1.f, n[2]);
    const float c    = -1.f / (sign + n[2]);
    const float d    = n[0] * n[1] * c;
    t                = {{1.f + sign * n[0] * n[0] * c, sign * d, -sign * n[0], 0.f}};
    b                = {{d, sign + n[1] * n[1] * c, -n[1], 0.f}};
}

struct Builder {
    const LightSet&             set;
    std::vector<uint32_t>&      mapping;
    std::vector<uint32_t>&      light_orders;
    std::vector<BuildNode>      build_nodes;
    std::vector<SplitCandidate> candidates;
    uint32_t                    current_node = 1;
    uint32_t                    light_order  = 0;

    Builder(const LightSet& s, std::vector<uint32_t>& m, std::vector<uint32_t>& o) : set(s), mapping(m), light_orders(o) {}

    void allocate(uint32_t num_lights, uint32_t sweep_threshold) {  // :430-444
        build_nodes.assign(size_t(2) * num_lights - 1, BuildNode{});
        const uint32_t num_slices = std::min(num_lights, sweep_threshold);
        candidates.assign(num_slices >= 2 ? size_t(num_slices) * 3 + 3 : 0, SplitCandidate{});
    }

    float variance(const uint32_t* lights, uint32_t num) const {  // :638-655
        float    ap = 0.f, aps = 0.f;
        uint32_t n = 0;
        for (uint32_t i = 0; i < num; ++i) {
            const float p = set.powers[lights[i]];
            if (p > 0.f) {
                n += 1;
                const float in = 1.f / float(n);
                ap += (p - ap) * in;
                aps += (p * p - aps) * in;
            }
        }
        return std::fabs(aps - ap * ap);
    }

    // evaluateSplits, :657-789 (candidates are evaluated serially: the result does not depend on the thread count)
    SplitCandidate evaluateSplits(uint32_t* lights, uint32_t len, const AABB& bounds, Vec4f cone, bool two_sided, uint32_t sweep_threshold) {
        uint32_t num_candidates = 0;

        if (2 == len) {
            candidates[num_candidates++].configurePartition(&lights[0], 1);
        } else if (3 == len) {
            candidates[num_candidates++].configurePartition(&lights[0], 1);
            candidates[num_candidates++].configurePartition(&lights[1], 1);
            candidates[num_candidates++].configurePartition(&lights[2], 1);
        } else if (4 == len) {
            candidates[num_candidates++].configurePartition(&lights[0], 1);
            candidates[num_candidates++].configurePartition(&lights[1], 1);
            candidates[num_candidates++].configurePartition(&lights[2], 1);
            candidates[num_candidates++].configurePartition(&lights[3], 1);
            const uint32_t p01[2] = {lights[0], lights[1]}, p02[2] = {lights[0], lights[2]}, p03[2] = {lights[0], lights[3]};
            candidates[num_candidates++].configurePartition(p01, 2);
            candidates[num_candidates++].configurePartition(p02, 2);
            candidates[num_candidates++].configurePartition(p03, 2);
        } else {
            if (len <= sweep_threshold) {
                for (uint32_t i = 0; i < len; ++i) {
                    const Vec4f max = set.aabbs[lights[i]].b[1];
                    candidates[num_candidates++].configure(max, 0);
                    candidates[num_candidates++].configure(max, 1);
                    candidates[num_candidates++].configure(max, 2);
                }
            } else {
                const Vec4f position = bounds.position();
                const Vec4f extent   = bounds.extent();
                const Vec4f min      = bounds.b[0];

                const uint32_t la   = indexMaxComponent3(extent);
                const float    step = extent[int(la)] / float(kNumSlices);

                for (int a = 0; a < 3; ++a) {
                    const float    extent_a  = extent[a];
                    const uint32_t num_steps = uint32_t(std::ceil(extent_a / step));
                    const float    step_a    = extent_a / float(num_steps);
                    for (uint32_t i = 1; i < num_steps; ++i) {
                        Vec4f slice = position;
                        slice[a]    = min[a] + float(i) * step_a;
                        candidates[num_candidates++].configure(slice, uint32_t(a));
                    }
                }
            }
            // :745-751: the three angle conditions are written to the same candidate; the last one (the cone axis) stays
            Vec4f t, b;
            orthonormalBasis(cone, t, b);
            candidates[num_candidates].configureAngle(t);
            candidates[num_candidates].configureAngle(b);
            candidates[num_candidates].configureAngle(cone);
            num_candidates += 1;
        }

        const float cone_weight = coneCost(cone[3], two_sided);
        for (uint32_t c = 0; c < num_candidates; ++c) candidates[c].evaluate(lights, len, bounds, cone_weight, set);

        float    min_cost = candidates[0].cost;
        uint32_t sc       = 0;
        for (uint32_t c = 1; c < num_candidates; ++c) {
            if (candidates[c].cost < min_cost) {
                sc       = c;
                min_cost = candidates[c].cost;
            }
        }
        return candidates[sc];
    }

    uint32_t assign(BuildNode& node, uint32_t begin, uint32_t end, const AABB& bounds, Vec4f cone, float total_power) {  // :558-613
        bool node_two_sided = false;
        for (uint32_t i = begin; i < end; ++i) {
            const uint32_t l = mapping[i];
            light_orders[l]  = light_order++;
            node_two_sided   = node_two_sided || set.twoSided(l);
        }
        node.bounds            = bounds;
        node.cone              = cone;
        node.power             = total_power;
        node.variance          = variance(mapping.data() + begin, end - begin);
        node.middle            = 0;
        node.children_or_light = begin;
        node.num_lights        = end - begin;
        node.two_sided         = set.primitive ? set.all_two_sided : node_two_sided;
        return end;
    }

    // base.memory.partition, src/base/memory/partition.zig:3-30
    uint32_t partition(uint32_t* data, uint32_t len, const SplitCandidate& sc) const {
        uint32_t first = len;
        for (uint32_t i = 0; i < len; ++i) {
            if (!sc.leftSide(data[i], set)) {
                first = i;
                break;
            }
        }
        if (first == len) return first;
        for (uint32_t i = first + 1; i < len; ++i) {
            if (sc.leftSide(data[i], set)) {
                std::swap(data[i], data[first]);
                first += 1;
            }
        }
        return first;
    }

    // split / splitPrimitive, :446-540
    uint32_t split(uint32_t node_id, uint32_t begin, uint32_t end, const AABB& bounds, Vec4f cone, bool two_sided, float total_power,
                   uint32_t depth) {
        uint32_t*      lights = mapping.data() + begin;
        const uint32_t len    = end - begin;

        const bool leaf = set.primitive ? len <= 4 : (1 == len || (2 == len && depth > kLightTreeMaxSplitDepth));
        if (leaf) return assign(build_nodes[node_id], begin, end, bounds, cone, total_power);

        const uint32_t child0 = current_node;

        const SplitCandidate sc = evaluateSplits(lights, len, bounds, cone, two_sided, set.primitive ? kPartSweepThreshold : kSceneSweepThreshold);
        if (sc.exhausted) return assign(build_nodes[node_id], begin, end, bounds, cone, total_power);

        const uint32_t split_node = begin + partition(lights, len, sc);

        current_node += 2;
        const uint32_t c0_end = split(child0, begin, split_node, sc.aabbs[0], sc.cones[0], sc.two_sided[0], sc.powers[0], depth + 1);
        const uint32_t c1_end = split(child0 + 1, split_node, end, sc.aabbs[1], sc.cones[1], sc.two_sided[1], sc.powers[1], depth + 1);

        BuildNode& node        = build_nodes[node_id];
        node.bounds            = bounds;
        node.cone              = cone;
        node.power             = total_power;
        node.variance          = variance(lights, len);
        node.middle            = c0_end;
        node.children_or_light = child0;
        node.num_lights        = len;
        node.two_sided         = two_sided;
        return c1_end;
    }

    // serialize, :615-636 + Node.compressCenter, light_tree.zig:40-54
    void serialize(LightTreeResult& out) {
        build_nodes[0].bounds.cacheRadius();
        const AABB total = build_nodes[0].bounds;
        out.nodes.assign(current_node, ZygpuLightNode{});
        out.node_middles.assign(current_node, 0);
        for (uint32_t i = 0; i < current_node; ++i) {
            const BuildNode& source = build_nodes[i];
            ZygpuLightNode&  dest   = out.nodes[i];

            const Vec4f p      = source.bounds.position();
            const Vec4f center = {{p[0], p[1], p[2], 0.5f * length3(source.bounds.extent())}};
            const Vec4f d      = center - total.b[0];
            const Vec4f e      = total.extent();
            const Vec4f div    = {{0.f == e[0] ? 1.f : e[0], 0.f == e[1] ? 1.f : e[1], 0.f == e[2] ? 1.f : e[2], total.b[1][3]}};
            const Vec4f q      = d / div;
            for (int k = 0; k < 4; ++k) {
                dest.center[k] = uint16_t(std::fmaf(q[k], 65535.f, 0.5f));                                          // enc.floatToUnorm16
                dest.cone[k]   = uint16_t((source.cone[k] + 1.f) * (source.cone[k] > 0.f ? 32767.5f : 32768.f));  // enc.floatToSnorm16
            }
            dest.power      = source.power;
            dest.variance   = source.variance;
            dest.meta       = (source.hasChildren() ? 1u : 0u) | (source.two_sided ? 2u : 0u) | (source.children_or_light << 2);
            dest.num_lights = source.num_lights;
            out.node_middles[i] = source.middle;
        }
        out.bounds     = total;
        out.root_power = build_nodes[0].power;
    }

    // BuildNode.countPotentialLights, :44-57
    void countPotentialLights(uint32_t node, uint32_t depth, uint32_t num_lights[][2]) const {
        const BuildNode& n = build_nodes[node];
        if (!n.hasChildren()) {
            num_lights[depth][0] += 1;
        } else {
            num_lights[depth][1] += 2;
            const uint32_t next_depth = depth + 1;
            if (next_depth < kLightTreeMaxSplitDepth) {
                countPotentialLights(n.children_or_light, next_depth, num_lights);
                countPotentialLights(n.children_or_light + 1, next_depth, num_lights);
            }
        }
    }
};

}  // namespace

// math.cone.merge, src/base/math/cone.zig:8-44 (Mat3x3.initRotation restated as written, matrix3x3.zig:50-77)
Vec4f coneMerge(Vec4f a, Vec4f b) {
    if (equal4(splat(1.f), a)) return b;
    if (equal4(a, b)) return a;

    float a_angle = std::acos(a[3]);
    float b_angle = std::acos(b[3]);
    if (b_angle > a_angle) {
        std::swap(a, b);
        std::swap(a_angle, b_angle);
    }

    const float d_angle = std::acos(clampf(dot3(a, b), -1.f, 1.f));
    if (fmin_(d_angle + b_angle, kPi) <= a_angle) return a;

    const float o_angle = (a_angle + d_angle + b_angle) / 2.f;
    if (o_angle >= kPi) return {{a[0], a[1], a[2], -1.f}};

    const float r_angle = o_angle - a_angle;
    const Vec4f v       = normalize3(cross3(a, b));

    const float c = std::cos(r_angle), s = std::sin(r_angle), t = 1.f - c;
    const float at0 = v[0] * v[1] * t, at1 = v[2] * s;
    const float bt0 = v[0] * v[2] * t, bt1 = v[1] * s;
    const float ct0 = v[1] * v[2] * t, ct1 = v[0] * s;
    const Vec4f r0 = {{c + v[0] * v[1] * t, at0 - at1, bt0 + bt1, 0.f}};
    const Vec4f r1 = {{at0 + at1, c + v[1] * v[1] * t, ct0 - ct1, 0.f}};
    const Vec4f r2 = {{bt0 - bt1, ct0 + ct1, c + v[2] * v[2] * t, 0.f}};

    // Mat3x3.transformVector, matrix3x3.zig:113-127
    Vec4f result = splat(a[0]) * r0;
    result       = mulAdd(splat(a[1]), r1, result);
    result       = mulAdd(splat(a[2]), r2, result);

    const Vec4f axis = normalize3(result);
    return {{axis[0], axis[1], axis[2], std::cos(o_angle)}};
}

namespace {

DeviceLightTreeFn g_device_builder     = nullptr;
uint32_t          g_device_min_lights = 0;

// BuildNode.countPotentialLights (:44-57) + the depth search of Builder.build (:350-361) on serialised nodes
uint32_t maxSplitDepth(const std::vector<ZygpuLightNode>& nodes, uint32_t num_infinite) {
    uint32_t split_lights[kLightTreeMaxSplitDepth][2] = {};
    struct Item {
        uint32_t node, depth;
    };
    std::vector<Item> stack{{0u, 0u}};
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const ZygpuLightNode& n = nodes[it.node];
        if (0 == (n.meta & 1u)) {
            split_lights[it.depth][0] += 1;
        } else {
            split_lights[it.depth][1] += 2;
            if (it.depth + 1 < kLightTreeMaxSplitDepth) {
                stack.push_back({(n.meta >> 2) + 1, it.depth + 1});
                stack.push_back({n.meta >> 2, it.depth + 1});
            }
        }
    }
    uint32_t num_split_lights = 0;
    for (uint32_t i = 0; i < kLightTreeMaxSplitDepth; ++i) {
        num_split_lights += split_lights[i][0];
        if ((num_split_lights + split_lights[i][1]) > (kLightTreeMaxLights - num_infinite) || 0 == split_lights[i][1]) return i;
    }
    return kLightTreeMaxSplitDepth;
}

}  // namespace

void setDeviceLightTreeBuilder(DeviceLightTreeFn fn, uint32_t min_lights) {
    g_device_builder    = fn;
    g_device_min_lights = min_lights;
}

void buildLightTree(const LightSet& set, std::vector<uint32_t>& mapping, uint32_t num_infinite, uint32_t first_order,
                    std::vector<uint32_t>& light_orders, LightTreeResult& out) {
    const uint32_t num_lights = uint32_t(mapping.size());
    const uint32_t num_finite = num_lights - num_infinite;

    out                 = LightTreeResult{};
    out.max_split_depth = kLightTreeMaxSplitDepth;
    out.root_power      = 0.f;
    out.bounds          = AABB::empty();
    if (0 == num_finite) return;

    if (g_device_builder && num_finite >= std::max(g_device_min_lights, 2u)) {
        std::vector<uint32_t> order;
        const std::vector<uint32_t> finite(mapping.begin() + num_infinite, mapping.end());
        if (g_device_builder(set, finite.data(), num_finite, first_order, out, order)) {
            for (uint32_t i = 0; i < num_finite; ++i) {
                const uint32_t l          = finite[order[i]];
                mapping[num_infinite + i] = l;
                light_orders[l]           = first_order + i;
            }
            out.bounds.cacheRadius();
            out.max_split_depth = maxSplitDepth(out.nodes, num_infinite);
            return;
        }
        out = LightTreeResult{};
        out.max_split_depth = kLightTreeMaxSplitDepth;
        out.root_power      = 0.f;
        out.bounds          = AABB::empty();
    }

    Builder builder(set, mapping, light_orders);
    builder.light_order = first_order;
    builder.allocate(num_finite, kSceneSweepThreshold);

    AABB  bounds      = AABB::empty();
    Vec4f cone        = splat(1.f);
    bool  two_sided   = false;
    float total_power = 0.f;
    for (uint32_t i = num_infinite; i < num_lights; ++i) {
        const uint32_t l = mapping[i];
        bounds.mergeAssign(set.aabbs[l]);
        cone      = coneMerge(cone, set.cones[l]);
        two_sided = two_sided || set.twoSided(l);
        total_power += set.powers[l];
    }

    builder.split(0, num_infinite, num_lights, bounds, cone, two_sided, total_power, 0);
    builder.serialize(out);

    // :350-361
    uint32_t split_lights[kLightTreeMaxSplitDepth][2] = {};
    builder.countPotentialLights(0, 0, split_lights);
    uint32_t num_split_lights = 0;
    for (uint32_t i = 0; i < kLightTreeMaxSplitDepth; ++i) {
        num_split_lights += split_lights[i][0];
        if ((num_split_lights + split_lights[i][1]) > (kLightTreeMaxLights - num_infinite) || 0 == split_lights[i][1]) {
            out.max_split_depth = i;
            break;
        }
    }
}

void buildPrimitiveLightTree(const LightSet& set, uint32_t num_triangles, const AABB& bounds, Vec4f cone, float total_power,
                             LightTreeResult& out) {
    out = LightTreeResult{};
    out.light_mapping.resize(num_triangles);
    out.light_orders.assign(num_triangles, 0);
    for (uint32_t l = 0; l < num_triangles; ++l) out.light_mapping[l] = l;
    out.max_split_depth = 0;
    if (0 == num_triangles) return;

    if (g_device_builder && num_triangles >= std::max(g_device_min_lights, 8u)) {
        std::vector<uint32_t> order;
        LightTreeResult       built;
        if (g_device_builder(set, out.light_mapping.data(), num_triangles, 0, built, order)) {
            built.light_mapping.resize(num_triangles);
            built.light_orders.assign(num_triangles, 0);
            for (uint32_t i = 0; i < num_triangles; ++i) {
                built.light_mapping[i]      = order[i];
                built.light_orders[order[i]] = i;
            }
            built.bounds.cacheRadius();
            built.max_split_depth = 0;
            out                   = std::move(built);
            return;
        }
    }

    Builder builder(set, out.light_mapping, out.light_orders);
    builder.allocate(num_triangles, kPartSweepThreshold);
    builder.split(0, 0, num_triangles, bounds, cone, set.all_two_sided, total_power, 0);
    builder.serialize(out);
}

}  // namespace zyg
