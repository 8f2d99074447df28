#include "bvh_builder.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>

namespace zyg {

namespace {

constexpr uint32_t ParallelizeThreshold = 1024;  // builder_base.zig:16

template <typename F>
void parallelFor(uint32_t count, uint32_t num_threads, F&& fn) {
    num_threads = std::min(num_threads, count);
    if (num_threads <= 1) {
        for (uint32_t i = 0; i < count; ++i) fn(i);
        return;
    }
    std::atomic<uint32_t>    next{0};
    std::vector<std::thread> pool;
    pool.reserve(num_threads);
    for (uint32_t t = 0; t < num_threads; ++t) {
        pool.emplace_back([&] {
            for (;;) {
                const uint32_t i = next.fetch_add(1, std::memory_order_relaxed);
                if (i >= count) return;
                fn(i);
            }
        });
    }
    for (auto& t : pool) t.join();
}

// split_candidate.zig:80-197
struct SplitCandidate {
    AABB     aabbs[2];
    uint32_t num_sides[2];
    float    d;
    float    cost;
    uint8_t  axis;
    bool     spatial;

    SplitCandidate(uint8_t split_axis, Vec4f p, bool sp) : d(p[split_axis]), cost(0.f), axis(split_axis), spatial(sp) {}

    bool behind(const float* point) const { return point[axis] < d; }  // :194-196

    void evaluate(const std::vector<Reference>& references, float aabb_surface_area) {  // :97-160
        uint32_t ns[2] = {0, 0};
        AABB     bs[2] = {AABB::empty(), AABB::empty()};

        if (spatial) {
            bool used_spatial = false;
            for (const Reference& r : references) {
                const AABB b = r.aabb();
                if (behind(r.max)) {
                    ns[0] += 1;
                    bs[0].mergeAssign(b);
                } else if (!behind(r.min)) {
                    ns[1] += 1;
                    bs[1].mergeAssign(b);
                } else {
                    ns[0] += 1;
                    ns[1] += 1;
                    bs[0].mergeAssign(b);
                    bs[1].mergeAssign(b);
                    used_spatial = true;
                }
            }
            if (used_spatial) {
                bs[0].clipMax(d, axis);
                bs[1].clipMin(d, axis);
            } else {
                spatial = false;
            }
        } else {
            for (const Reference& r : references) {
                const AABB b = r.aabb();
                if (behind(r.max)) {
                    ns[0] += 1;
                    bs[0].mergeAssign(b);
                } else {
                    ns[1] += 1;
                    bs[1].mergeAssign(b);
                }
            }
        }

        const uint32_t n          = uint32_t(references.size());
        const bool     empty_side = 0 == ns[0] || 0 == ns[1];
        if (empty_side) {
            cost = 2.f + float(n);
        } else {
            const float weight0             = float(ns[0]) * bs[0].surfaceArea();
            const float weight1             = float(ns[1]) * bs[1].surfaceArea();
            const float duplication_penalty = 0.125f * float(ns[0] + ns[1] - n);
            cost                            = 2.f + (weight0 + weight1) / aabb_surface_area + duplication_penalty;
        }

        num_sides[0] = ns[0];
        num_sides[1] = ns[1];
        aabbs[0]     = bs[0];
        aabbs[1]     = bs[1];
    }

    void distribute(const std::vector<Reference>& references, std::vector<Reference>& r0,
                    std::vector<Reference>& r1) const {  // :162-192
        r0.reserve(num_sides[0]);
        r1.reserve(num_sides[1]);
        if (spatial) {
            for (const Reference& r : references) {
                if (behind(r.max)) {
                    r0.push_back(r);
                } else if (!behind(r.min)) {
                    r1.push_back(r);
                } else {
                    Reference a = r, b = r;
                    a.max[axis] = fmin_(d, r.max[axis]);  // clippedMax :67-73
                    b.min[axis] = fmax_(d, r.min[axis]);  // clippedMin :59-65
                    r0.push_back(a);
                    r1.push_back(b);
                }
            }
        } else {
            for (const Reference& r : references) {
                if (behind(r.max)) {
                    r0.push_back(r);
                } else {
                    r1.push_back(r);
                }
            }
        }
    }
};

struct Kernel;

struct Task {  // builder_base.zig:18-27
    Kernel*                kernel;
    uint32_t               root;
    uint32_t               depth;
    AABB                   aabb;
    std::vector<Reference> references;
};

struct Kernel {  // builder_base.zig:31-318
    BuildSettings         settings;
    std::vector<BvhNode>  build_nodes;
    std::vector<uint32_t> reference_ids;
    uint32_t              num_degenerate = 0;

    void reserve(uint32_t num_primitives, const BuildSettings& s) {  // :305-317
        settings = s;
        build_nodes.clear();
        build_nodes.reserve(std::max((3 * num_primitives) / s.max_primitives, 1u));
        build_nodes.push_back(BvhNode{});
        reference_ids.clear();
        reference_ids.reserve((size_t(num_primitives) * 12) / 10);
    }

    void assign(uint32_t node_id, const std::vector<Reference>& references) {  // :295-303
        build_nodes[node_id].setLeafNode(uint32_t(reference_ids.size()), uint32_t(references.size()));
        for (const Reference& r : references) reference_ids.push_back(r.primitive());
    }

    // :165-283. `eval_threads` > 1 only on the calling thread outside of the task phase
    // (the reference's `threads.running_parallel == false` case).
    bool splittingPlane(const std::vector<Reference>& references, const AABB& aabb, uint32_t depth,
                        uint32_t eval_threads, std::vector<SplitCandidate>& cands, SplitCandidate& out) {
        cands.clear();

        const float aabb_surface_area = aabb.surfaceArea();
        if (0.f == aabb_surface_area) return false;

        const uint32_t num_references = uint32_t(references.size());
        const Vec4f    position       = aabb.position();

        cands.emplace_back(0, position, true);
        cands.emplace_back(1, position, true);
        cands.emplace_back(2, position, true);

        if (num_references <= settings.sweep_threshold) {
            for (const Reference& r : references) {
                const Vec4f max = {{r.max[0], r.max[1], r.max[2], 0.f}};
                cands.emplace_back(0, max, false);
                cands.emplace_back(1, max, false);
                cands.emplace_back(2, max, false);
            }
        } else {
            const Vec4f    extent = aabb.extent();
            const Vec4f    min    = aabb.b[0];
            const uint32_t la     = indexMaxComponent3(extent);
            const float    step   = extent[la] / float(settings.num_slices);

            for (uint8_t a = 0; a < 3; ++a) {
                const float    extent_a  = extent[a];
                const uint32_t num_steps = std::max(1u, uint32_t(std::ceil(extent_a / step)));
                const float    step_a    = extent_a / float(num_steps);

                for (uint32_t i = 1; i < num_steps; ++i) {
                    const float fi    = float(i);
                    Vec4f       slice = position;
                    slice[a]          = min[a] + fi * step_a;
                    cands.emplace_back(a, slice, false);
                    if (depth < settings.spatial_split_threshold) cands.emplace_back(a, slice, true);
                }
            }
        }

        if (eval_threads <= 1 || references.size() < ParallelizeThreshold) {
            for (SplitCandidate& sc : cands) sc.evaluate(references, aabb_surface_area);
        } else {
            parallelFor(uint32_t(cands.size()), eval_threads,
                        [&](uint32_t i) { cands[i].evaluate(references, aabb_surface_area); });
        }

        size_t sc       = 0;
        float  min_cost = cands[0].cost;
        for (size_t i = 1; i < cands.size(); ++i) {
            const float cost = cands[i].cost;
            if (cost < min_cost) {
                sc       = i;
                min_cost = cost;
            }
        }

        const SplitCandidate& sp = cands[sc];
        if ((sp.aabbs[0].covers(aabb) && num_references == sp.num_sides[0]) ||
            (sp.aabbs[1].covers(aabb) && num_references == sp.num_sides[1])) {
            return false;
        }

        out = sp;
        return true;
    }

    // :65-163. `tasks` == nullptr corresponds to `threads.running_parallel` (inside a task).
    void split(uint32_t node_id, std::vector<Reference>&& references, const AABB& aabb, uint32_t depth,
               uint32_t eval_threads, std::vector<Task>* tasks, bool tasks_enabled) {
        build_nodes[node_id].setAABB(aabb);

        const uint32_t num_primitives = uint32_t(references.size());
        if (num_primitives <= settings.max_primitives) {
            assign(node_id, references);
            return;
        }

        if (tasks && tasks_enabled && (num_primitives < ParallelizeThreshold || depth == settings.parallel_build_depth)) {
            tasks->push_back(Task{nullptr, node_id, depth, aabb, std::move(references)});
            return;
        }

        std::vector<SplitCandidate> cands;
        SplitCandidate              sp(0, splat(0.f), false);
        if (splittingPlane(references, aabb, depth, tasks ? eval_threads : 1, cands, sp)) {
            if (num_primitives <= 0xFF && float(num_primitives) <= sp.cost) {
                assign(node_id, references);
            } else {
                std::vector<Reference> references0, references1;
                sp.distribute(references, references0, references1);

                if (num_primitives <= 0x2FF && (references0.empty() || references1.empty())) {
                    // Every primitive ended up (partially) on the same side of the plane.
                    assign(node_id, references);
                } else {
                    const uint32_t child0 = uint32_t(build_nodes.size());
                    build_nodes[node_id].setSplitNode(child0);
                    build_nodes.push_back(BvhNode{});
                    build_nodes.push_back(BvhNode{});

                    std::vector<Reference>().swap(references);
                    cands.clear();
                    cands.shrink_to_fit();

                    const uint32_t next_depth = depth + 1;
                    split(child0, std::move(references0), sp.aabbs[0].intersection(aabb), next_depth, eval_threads, tasks,
                          tasks_enabled);
                    split(child0 + 1, std::move(references1), sp.aabbs[1].intersection(aabb), next_depth, eval_threads,
                          tasks, tasks_enabled);
                }
            }
        } else {
            // The reference logs "Cannot split node further" above 0x2FF primitives and leaves the
            // node undefined (:158-160); we keep the primitives in a leaf and count the event.
            if (num_primitives > 0x2FF) num_degenerate += 1;
            assign(node_id, references);
        }
    }
};

}  // namespace

void buildBinaryBvh(std::vector<Reference>&& references, const AABB& bounds, uint32_t num_slices,
                    uint32_t sweep_threshold, uint32_t max_primitives, uint32_t num_threads, BuildResult& out) {
    BuildSettings settings{num_slices, sweep_threshold, max_primitives};

    const uint32_t num_references = uint32_t(references.size());

    // builder_base.zig:330-332
    const float log2_num_references  = std::log2(float(num_references));
    settings.spatial_split_threshold = uint32_t(std::round(log2_num_references / 2.f));
    settings.parallel_build_depth    = std::min(settings.spatial_split_threshold, 6u);

    Kernel main;
    main.reserve(num_references, settings);

    // :337-340 — a task list with zero capacity disables the decomposition entirely
    const uint32_t num_tasks = std::min(1u << settings.parallel_build_depth, num_references / ParallelizeThreshold);

    std::vector<Task> tasks;
    main.split(0, std::move(references), bounds, 0, num_threads, &tasks, num_tasks > 0);

    // :354-390
    std::vector<Kernel> kernels(tasks.size());
    parallelFor(uint32_t(tasks.size()), num_threads, [&](uint32_t i) {
        Task&   t = tasks[i];
        Kernel& k = kernels[i];
        k.reserve(uint32_t(t.references.size()), settings);
        k.split(0, std::move(t.references), t.aabb, t.depth, 1, nullptr, false);
    });

    for (size_t i = 0; i < tasks.size(); ++i) {
        const Task&                 t        = tasks[i];
        const std::vector<BvhNode>& children = kernels[i].build_nodes;

        main.build_nodes[t.root] = children[0];
        main.num_degenerate += kernels[i].num_degenerate;

        if (1 == children.size()) {
            // NOTE: the reference `continue`s here without appending the task's reference ids
            // (:368-370), which leaves a leaf pointing at foreign ids. We append them and point the
            // leaf at its own ids; a task root can only become a leaf through the cost test.
            const uint32_t reference_offset = uint32_t(main.reference_ids.size());
            main.reference_ids.insert(main.reference_ids.end(), kernels[i].reference_ids.begin(),
                                      kernels[i].reference_ids.end());
            main.build_nodes[t.root].min_data += reference_offset;
            continue;
        }

        const uint32_t node_offset      = uint32_t(main.build_nodes.size() - 1);
        const uint32_t reference_offset = uint32_t(main.reference_ids.size());

        main.reference_ids.insert(main.reference_ids.end(), kernels[i].reference_ids.begin(),
                                  kernels[i].reference_ids.end());

        main.build_nodes[t.root].min_data += node_offset;

        for (size_t c = 1; c < children.size(); ++c) {
            BvhNode sn = children[c];
            sn.min_data += (0 == sn.numIndices()) ? node_offset : reference_offset;  // Node.initFrom
            main.build_nodes.push_back(sn);
        }
    }

    out.build_nodes           = std::move(main.build_nodes);
    out.reference_ids         = std::move(main.reference_ids);
    out.num_degenerate_leaves = main.num_degenerate;
}

}  // namespace zyg
