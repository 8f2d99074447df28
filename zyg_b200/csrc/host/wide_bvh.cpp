#include "wide_bvh.hpp"

#include <algorithm>
#include <cmath>
#include <deque>

namespace zyg {

namespace {

struct Box {
    float lo[3], hi[3];

    static Box empty() { return {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; }
    void       grow(const Box& o) {
        for (int i = 0; i < 3; ++i) {
            lo[i] = std::min(lo[i], o.lo[i]);
            hi[i] = std::max(hi[i], o.hi[i]);
        }
    }
    float area() const {
        const float d[3] = {std::max(hi[0] - lo[0], 0.f), std::max(hi[1] - lo[1], 0.f), std::max(hi[2] - lo[2], 0.f)};
        return d[0] * d[1] + d[0] * d[2] + d[1] * d[2];
    }
};

// Binary tree with leaves of at most three triangles ("augmented" tree): reference leaves that
// hold more are bisected by centroid into virtual nodes with boxes clipped to the reference leaf.
struct ANode {
    Box      box;
    int32_t  left  = -1;  // -1 => leaf
    int32_t  right = -1;
    uint32_t first = 0;  // into `leaf_prims`
    uint32_t count = 0;
    Box      leaf_box;  // leaves: exact box of the reference leaf (the gate), see wide_bvh.hpp
};

// `ItemBox(item)` = the unclipped box of leaf item `item` (a BVH-order triangle, or a position in a prop tree's index list).
template <typename ItemBox>
struct Augment {
    const BvhNode*        binary_nodes;
    ItemBox               itemBox;
    std::vector<ANode>    nodes;
    std::vector<uint32_t> leaf_prims;  // leaf items in leaf order

    Box triBox(uint32_t prim, const Box& clip) const {
        Box b = itemBox(prim);
        for (int i = 0; i < 3; ++i) {
            b.lo[i] = std::max(b.lo[i], clip.lo[i]);
            b.hi[i] = std::min(b.hi[i], clip.hi[i]);
            if (b.lo[i] > b.hi[i]) {  // touching only; keep a valid (degenerate) box inside the clip region
                b.lo[i] = b.hi[i] = std::min(std::max(b.lo[i], clip.lo[i]), clip.hi[i]);
            }
        }
        return b;
    }

    int32_t makeLeafTree(std::vector<uint32_t>& prims, size_t begin, size_t end, const Box& clip) {
        const int32_t id = int32_t(nodes.size());
        nodes.emplace_back();
        Box box = Box::empty();
        for (size_t i = begin; i < end; ++i) box.grow(triBox(prims[i], clip));
        nodes[id].box = box;

        const size_t n = end - begin;
        if (n <= 3) {
            nodes[id].first    = uint32_t(leaf_prims.size());
            nodes[id].count    = uint32_t(n);
            nodes[id].leaf_box = clip;
            for (size_t i = begin; i < end; ++i) leaf_prims.push_back(prims[i]);
            return id;
        }

        int axis = 0;
        {
            const float d[3] = {box.hi[0] - box.lo[0], box.hi[1] - box.lo[1], box.hi[2] - box.lo[2]};
            if (d[1] > d[axis]) axis = 1;
            if (d[2] > d[axis]) axis = 2;
        }
        std::stable_sort(prims.begin() + begin, prims.begin() + end, [&](uint32_t x, uint32_t y) {
            const Box bx = triBox(x, clip), by = triBox(y, clip);
            return bx.lo[axis] + bx.hi[axis] < by.lo[axis] + by.hi[axis];
        });
        const size_t  mid = begin + (n + 1) / 2;
        const int32_t l   = makeLeafTree(prims, begin, mid, clip);
        const int32_t r   = makeLeafTree(prims, mid, end, clip);
        nodes[id].left    = l;
        nodes[id].right   = r;
        return id;
    }

    void setGate(int32_t id, const Box& gate) {
        if (nodes[id].left < 0) {
            nodes[id].leaf_box = gate;
        } else {
            setGate(nodes[id].left, gate);
            setGate(nodes[id].right, gate);
        }
    }

    int32_t convert(uint32_t n) {
        // iterative post-order would be nicer; depth is bounded by the binary tree depth (< 128, the
        // reference's own traversal stack limit node_stack.zig:2), so recursion is safe here.
        const BvhNode& node = binary_nodes[n];
        Box            box;
        for (int i = 0; i < 3; ++i) {
            box.lo[i] = node.min[i];
            box.hi[i] = node.max[i];
        }
        if (0 != node.numIndices()) {
            if (0 == n) {
                // A tree that is a single leaf: the reference tests its triangles without any box test
                // (triangle_tree.zig:57-72), so the gate must always pass.
                for (int i = 0; i < 3; ++i) {
                    box.lo[i] = -INFINITY;
                    box.hi[i] = INFINITY;
                }
                Box tight;
                for (int i = 0; i < 3; ++i) {
                    tight.lo[i] = node.min[i];
                    tight.hi[i] = node.max[i];
                }
                std::vector<uint32_t> prims(node.numIndices());
                for (uint32_t i = 0; i < node.numIndices(); ++i) prims[i] = node.indicesStart() + i;
                const int32_t id = makeLeafTree(prims, 0, prims.size(), tight);
                setGate(id, box);
                return id;
            }
            std::vector<uint32_t> prims(node.numIndices());
            for (uint32_t i = 0; i < node.numIndices(); ++i) prims[i] = node.indicesStart() + i;
            return makeLeafTree(prims, 0, prims.size(), box);
        }
        const int32_t id = int32_t(nodes.size());
        nodes.emplace_back();
        nodes[id].box   = box;
        const int32_t l = convert(node.children());
        const int32_t r = convert(node.children() + 1);
        nodes[id].left  = l;
        nodes[id].right = r;
        return id;
    }
};

inline float exp2i(int e) { return std::ldexp(1.f, e); }

// Collapses the augmented binary tree below `root` into 8-wide quantised nodes, breadth first. `emit(item, gate)` appends the
// record of one leaf item (in slot order: records of a node's leaf slots are consecutive from `tri_base`); returns max depth.
template <typename Aug, typename Emit>
uint32_t collapseWide(const Aug& aug, int32_t root, std::vector<WideNode>& out_nodes, uint32_t& num_records, Emit&& emit) {
    uint32_t max_depth = 0;
    struct Work {
        int32_t  anode;  // augmented node that becomes this wide node
        uint32_t wide;
        uint32_t depth;
    };
    std::deque<Work> queue;
    out_nodes.emplace_back();
    queue.push_back({root, 0, 1});

    while (!queue.empty()) {
        const Work w = queue.front();
        queue.pop_front();
        max_depth = std::max(max_depth, w.depth);

        // 1. gather up to eight children by repeatedly opening the inner child with the largest area
        int32_t  children[8];
        uint32_t num = 0;
        {
            const ANode& a = aug.nodes[w.anode];
            if (a.left < 0) {
                children[num++] = w.anode;  // whole tree is a single small leaf
            } else {
                children[num++] = a.left;
                children[num++] = a.right;
            }
        }
        while (num < 8) {
            int   best      = -1;
            float best_area = -1.f;
            for (uint32_t i = 0; i < num; ++i) {
                const ANode& c = aug.nodes[children[i]];
                if (c.left >= 0) {
                    const float ar = c.box.area();
                    if (ar > best_area) {
                        best_area = ar;
                        best      = int(i);
                    }
                }
            }
            if (best < 0) break;
            const ANode& c   = aug.nodes[children[best]];
            children[best]   = c.left;
            children[num++]  = c.right;
        }

        // 2. node frame
        Box nb = Box::empty();
        for (uint32_t i = 0; i < num; ++i) nb.grow(aug.nodes[children[i]].box);

        WideNode node;
        std::memset(&node, 0, sizeof(node));
        int ex[3];
        for (int a = 0; a < 3; ++a) {
            node.p[a]          = nb.lo[a];
            const double extent = double(nb.hi[a]) - double(nb.lo[a]);
            int          e      = extent > 0.0 ? int(std::ceil(std::log2(extent / 255.0))) : -126;
            e                   = std::max(e, -126);
            while (double(exp2i(e)) * 255.0 < extent) ++e;
            ex[a] = e;
        }

        // 3. slot assignment: greedy on dot(child centre - node centre, slot diagonal)
        int   slot_of[8];
        bool  slot_used[8]  = {false, false, false, false, false, false, false, false};
        bool  child_done[8] = {false, false, false, false, false, false, false, false};
        float cost[8][8];
        for (uint32_t c = 0; c < num; ++c) {
            const Box& cb = aug.nodes[children[c]].box;
            float      d[3];
            for (int a = 0; a < 3; ++a) d[a] = 0.5f * (cb.lo[a] + cb.hi[a]) - 0.5f * (nb.lo[a] + nb.hi[a]);
            for (int s = 0; s < 8; ++s) {
                const float sx = (s & 4) ? 1.f : -1.f, sy = (s & 2) ? 1.f : -1.f, sz = (s & 1) ? 1.f : -1.f;
                cost[c][s]     = d[0] * sx + d[1] * sy + d[2] * sz;
            }
        }
        for (uint32_t k = 0; k < num; ++k) {
            float best = -FLT_MAX;
            int   bc = -1, bs = -1;
            for (uint32_t c = 0; c < num; ++c) {
                if (child_done[c]) continue;
                for (int s = 0; s < 8; ++s) {
                    if (slot_used[s]) continue;
                    if (cost[c][s] > best) {
                        best = cost[c][s];
                        bc   = int(c);
                        bs   = s;
                    }
                }
            }
            child_done[bc] = true;
            slot_used[bs]  = true;
            slot_of[bc]    = bs;
        }
        int child_in_slot[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
        for (uint32_t c = 0; c < num; ++c) child_in_slot[slot_of[c]] = int(c);

        // 4. fill slots in slot order (inner children and triangles are addressed by rank)
        node.child_base = uint32_t(out_nodes.size());
        node.tri_base   = num_records;
        uint32_t tri_offset = 0;
        for (int s = 0; s < 8; ++s) {
            const int c = child_in_slot[s];
            if (c < 0) continue;
            const ANode& ch = aug.nodes[children[c]];

            for (int a = 0; a < 3; ++a) {
                const double cell = double(exp2i(ex[a]));
                const double p    = double(node.p[a]);
                int          ql   = int(std::floor((double(ch.box.lo[a]) - p) / cell));
                int          qh   = int(std::ceil((double(ch.box.hi[a]) - p) / cell));
                ql                = std::min(std::max(ql, 0), 255);
                qh                = std::min(std::max(qh, 0), 255);
                while (ql > 0 && p + ql * cell > double(ch.box.lo[a])) --ql;
                while (qh < 255 && p + qh * cell < double(ch.box.hi[a])) ++qh;
                node.qlo[a][s] = uint8_t(ql);
                node.qhi[a][s] = uint8_t(qh);
            }

            if (ch.left >= 0) {
                node.imask |= uint8_t(1u << s);
                node.meta[s] = uint8_t((1u << 5) | (24u + uint32_t(s)));
                const uint32_t wi = uint32_t(out_nodes.size());
                out_nodes.emplace_back();
                queue.push_back({children[c], wi, w.depth + 1});
            } else {
                const uint32_t unary = (1u << ch.count) - 1u;  // 1 -> 0b001, 2 -> 0b011, 3 -> 0b111
                node.meta[s]         = uint8_t((unary << 5) | tri_offset);
                for (uint32_t i = 0; i < ch.count; ++i) emit(aug.leaf_prims[ch.first + i], ch.leaf_box);
                num_records += ch.count;
                tri_offset += ch.count;
            }
        }
        for (int a = 0; a < 3; ++a) node.e[a] = uint8_t(ex[a] + 127);

        out_nodes[w.wide] = node;
    }
    return max_depth;
}

}  // namespace

void buildWideBvh(const TriangleTree& tree, WideBvh& out) {
    auto item_box = [&tree](uint32_t prim) {
        Box b = Box::empty();
        for (int k = 0; k < 3; ++k) {
            const float* p = &tree.positions[size_t(tree.triangles[size_t(prim) * 3 + k]) * 3];
            for (int i = 0; i < 3; ++i) {
                b.lo[i] = std::min(b.lo[i], p[i]);
                b.hi[i] = std::max(b.hi[i], p[i]);
            }
        }
        return b;
    };
    Augment<decltype(item_box)> aug{tree.nodes.data(), item_box, {}, {}};
    aug.nodes.reserve(tree.nodes.size() * 2);
    aug.leaf_prims.reserve(tree.numTriangles());
    const int32_t root = aug.convert(0);

    out.nodes.clear();
    out.triangles.clear();
    out.nodes.reserve(tree.nodes.size() / 3 + 1);
    out.triangles.reserve(tree.numTriangles());

    auto emitTriangle = [&](uint32_t prim, const Box& gate) {
        TriRecord    r;
        const float* a = &tree.positions[size_t(tree.triangles[size_t(prim) * 3 + 0]) * 3];
        const float* b = &tree.positions[size_t(tree.triangles[size_t(prim) * 3 + 1]) * 3];
        const float* c = &tree.positions[size_t(tree.triangles[size_t(prim) * 3 + 2]) * 3];
        for (int i = 0; i < 3; ++i) {
            r.a[i]  = a[i];
            r.e1[i] = b[i] - a[i];
            r.e2[i] = c[i] - a[i];
        }
        r.primitive  = prim;
        r.leaf_min_x = gate.lo[0];
        r.leaf_min_y = gate.lo[1];
        r.leaf_min_z = gate.lo[2];
        for (int i = 0; i < 3; ++i) r.leaf_max[i] = gate.hi[i];
        out.triangles.push_back(r);
    };
    uint32_t num_records = 0;
    out.max_depth        = collapseWide(aug, root, out.nodes, num_records, emitTriangle);

    const BvhNode& top = tree.nodes[0];
    double         c[3], r2 = 0.0;
    for (int i = 0; i < 3; ++i) c[i] = 0.5 * (double(top.min[i]) + double(top.max[i]));
    for (uint32_t index : tree.triangles) {
        const float* p = &tree.positions[size_t(index) * 3];
        double       d2 = 0.0;
        for (int i = 0; i < 3; ++i) d2 += (double(p[i]) - c[i]) * (double(p[i]) - c[i]);
        r2 = std::max(r2, d2);
    }
    for (int i = 0; i < 3; ++i) out.bound_center[i] = float(c[i]);
    out.bound_radius = float(std::sqrt(r2) * (1.0 + 1e-6)) + FLT_MIN;
}

void buildWidePropBvh(const ZygpuBvhNode* nodes, uint32_t num_nodes, const uint32_t* indices, const ZygpuAabb* aabbs, const float* spheres,
                      WidePropBvh& out) {
    out.nodes.clear();
    out.records.clear();
    out.max_depth = 0;
    if (0 == num_nodes) return;
    static_assert(sizeof(ZygpuBvhNode) == sizeof(BvhNode), "the prop tree uses the bvh.Node layout");
    auto item_box = [&](uint32_t position) {
        const ZygpuAabb& b = aabbs[indices[position]];
        return Box{{b.min[0], b.min[1], b.min[2]}, {b.max[0], b.max[1], b.max[2]}};
    };
    Augment<decltype(item_box)> aug{reinterpret_cast<const BvhNode*>(nodes), item_box, {}, {}};
    aug.nodes.reserve(size_t(num_nodes) * 2);
    const int32_t root = aug.convert(0);

    auto emit = [&](uint32_t position, const Box& gate) {
        PropRecord r;
        for (int i = 0; i < 3; ++i) {
            r.leaf_min[i] = gate.lo[i];
            r.leaf_max[i] = gate.hi[i];
        }
        r.prop = indices[position];
        r.pad  = 0;
        for (int i = 0; i < 4; ++i) r.sphere[i] = spheres ? spheres[size_t(r.prop) * 4 + i] : (3 == i ? FLT_MAX : 0.f);
        for (int i = 0; i < 4; ++i) r.pad2[i] = 0.f;
        out.records.push_back(r);
    };
    uint32_t num_records = 0;
    out.max_depth        = collapseWide(aug, root, out.nodes, num_records, emit);
}

}  // namespace zyg
