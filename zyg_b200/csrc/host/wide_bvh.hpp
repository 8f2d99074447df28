// Device layout of a triangle BVH: 8-wide nodes with 8-bit quantised child boxes + 64-byte
// triangle records. Derived ("flattened on upload") from the reference-order binary tree, so the
// set of triangles and their clip regions is the reference's; only the visiting order changes.
//
// Node, 80 bytes + 16 bytes of padding = three 32-byte loads (LDG.E.256):
//   +0   float  p[3]        origin (min corner) of the quantisation grid
//   +12  u8     e[3]        biased exponents: cell size = 2^(e-127) per axis
//   +15  u8     imask       bit s set: slot s holds an inner node
//   +16  u32    child_base  index of the first inner child; inner children are consecutive in slot order
//   +20  u32    tri_base    index of the first triangle record referenced by leaf slots
//   +24  u8     meta[8]     inner: 0b001_11sss (s = slot); leaf: unary count (1..3) << 5 | first triangle offset;
//                           empty: 0
//   +32  u8     qlo[3][8]   quantised lower bounds, x then y then z, one byte per slot
//   +56  u8     qhi[3][8]   quantised upper bounds
// Slots are assigned so that `slot ^ octant(ray)` visits children roughly front to back.
//
// Triangle record, 64 bytes = four 16-byte loads (two whole 32-byte sectors): a, e1 = b - a,
// e2 = c - a (the same fp32 subtractions the reference does per test, triangle.zig:27-28),
// `primitive` = index into the reference-order triangle list, and the exact fp32 box of the
// reference leaf the triangle sits in. The kernels gate every triangle test with the reference's own
// slab test on that box (node.zig:73-87): the reference only ever tests a triangle after its leaf box
// passed, and the box test is not watertight against the triangle test, so without the gate the wide
// path would report hits the reference misses (rays through shared edges of axis-aligned geometry).
#pragma once

#include "../../../include/zygpu_scene.h"
#include "triangle_tree.hpp"

namespace zyg {

struct WideNode {
    float    p[3];
    uint8_t  e[3];
    uint8_t  imask;
    uint32_t child_base;
    uint32_t tri_base;
    uint8_t  meta[8];
    uint8_t  qlo[3][8];
    uint8_t  qhi[3][8];
    uint8_t  pad[16];  // to 96 bytes: a node is three aligned 32-byte (256-bit) loads on the device
};
static_assert(sizeof(WideNode) == 96, "wide node must be three 32-byte words");

struct TriRecord {
    float    a[3];
    uint32_t primitive;
    float    e1[3];
    float    leaf_min_x;
    float    e2[3];
    float    leaf_min_y;
    float    leaf_min_z;
    float    leaf_max[3];
};
static_assert(sizeof(TriRecord) == 64, "triangle record must be four 16-byte words");

struct WideBvh {
    std::vector<WideNode>  nodes;
    std::vector<TriRecord> triangles;
    uint32_t               max_depth = 0;  // in wide nodes, root = 1
    // bounding sphere of the referenced vertices around the centre of the root box (object space): an instance of the mesh is
    // only entered by rays that pass its sphere (culling only: no triangle lies outside)
    float bound_center[3] = {0.f, 0.f, 0.f};
    float bound_radius    = 0.f;
};

void buildWideBvh(const TriangleTree& tree, WideBvh& out);

// The same node format over a prop tree (PropBvh.Tree, prop_tree.zig:31-36): the "two-level layout for prop instances". Leaf
// slots reference prop records instead of triangle records: the prop id and the exact box of the reference leaf the prop sits in
// (a prop duplicated by a spatial split has one record per leaf). The kernels gate a prop with the reference's slab test on that
// box, then test the prop's own world box like Prop.intersect does (prop.zig:163-197). 64 bytes = two 32-byte loads.
struct PropRecord {
    float    leaf_min[3];
    uint32_t prop;
    float    leaf_max[3];
    uint32_t pad;
    float    sphere[4];  // world-space bounding sphere of a mesh prop (centre, radius); radius FLT_MAX = no sphere test
    float    pad2[4];
};
static_assert(sizeof(PropRecord) == 64, "prop record must be two 32-byte words");

struct WidePropBvh {
    std::vector<WideNode>   nodes;
    std::vector<PropRecord> records;
    uint32_t                max_depth = 0;
};

// `spheres`: 4 floats per prop (world-space centre, radius; FLT_MAX radius for props without one), may be null.
void buildWidePropBvh(const ZygpuBvhNode* nodes, uint32_t num_nodes, const uint32_t* indices, const ZygpuAabb* aabbs, const float* spheres,
                      WidePropBvh& out);

}  // namespace zyg
