// The object behind the opaque `zyg_mesh` handle of include/zygpu.h.
#pragma once

#include "wide_bvh.hpp"

#include <atomic>
#include <cstdint>

struct zyg_mesh {
    // process-unique: a device recognises an already uploaded mesh by it (an address can be reused after zyg_mesh_free)
    uint64_t serial = nextSerial();

    static uint64_t nextSerial() {
        static std::atomic<uint64_t> counter{1};
        return counter.fetch_add(1);
    }

    zyg::TriangleTree tree;
    zyg::WideBvh      wide;
};
