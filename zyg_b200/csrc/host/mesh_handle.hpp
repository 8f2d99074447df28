// The object behind the opaque `zyg_mesh` handle of include/zygpu.h.
#pragma once

#include "wide_bvh.hpp"

struct zyg_mesh {
    zyg::TriangleTree tree;
    zyg::WideBvh      wide;
};
