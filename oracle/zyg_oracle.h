/* ORACLE — test infrastructure only. CPU restatement of the reference (Opioid/zyg) hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. The product (zyg_b200/) never does.
 *
 * PARITY UNPINNED for geometry and shading: the reference has no automated tests, golden vectors or
 * fixtures and cannot be built here (Zig, no toolchain). Pins that do exist:
 * the published PCG32 known-answer vector (pcg32.cpp) and structural self-checks in tests/.
 */
#ifndef ZYG_ORACLE_H
#define ZYG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ZoRay {
    float origin[3];
    float min_t;
    float direction[3];
    float max_t;
} ZoRay;

typedef struct ZoHit {
    float    t, u, v;
    uint32_t primitive;
} ZoHit;

/* TriangleTree.intersect (triangle_tree.zig:46-109), identity transformation. threads = 0: all cores. */
void zo_trace_closest(const void* nodes, const uint32_t* triangles, const float* positions, const ZoRay* rays,
                      uint64_t n, ZoHit* out, uint32_t threads, uint64_t* visited_nodes, uint64_t* tested_tris);

/* TriangleTree.intersectP (triangle_tree.zig:197-242). out[i] = 1 if occluded. */
void zo_trace_any(const void* nodes, const uint32_t* triangles, const float* positions, const ZoRay* rays, uint64_t n,
                  uint32_t* out, uint32_t threads);

/* O(N) closest hit over all triangles with the reference's accept rule; no tree involved. */
void zo_brute_closest(const uint32_t* triangles, uint32_t num_triangles, const float* positions, const ZoRay* rays,
                      uint64_t n, ZoHit* out, uint32_t* num_ties, uint32_t threads);

/* rnd.Generator (src/base/random/generator.zig:1-47): start(state, sequence) then n draws. */
void zo_pcg32_uints(uint64_t state, uint64_t sequence, uint32_t n, uint32_t* out);
void zo_pcg32_floats(uint64_t state, uint64_t sequence, uint32_t n, float* out);

#ifdef __cplusplus
}
#endif

#endif
