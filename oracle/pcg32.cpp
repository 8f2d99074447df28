// ORACLE — test infrastructure only. rnd.Generator, src/base/random/generator.zig:1-47 (PCG32,
// XSH-RR, O'Neill 2014). Pinned by the published pcg32-demo vector (seed 42, stream 54).
#include "zyg_oracle.h"
#include "zsampler.hpp"




extern "C" {

void zo_pcg32_uints(uint64_t state, uint64_t sequence, uint32_t n, uint32_t* out) {
    zo::Generator g;
    g.start(state, sequence);
    for (uint32_t i = 0; i < n; ++i) out[i] = g.randomUint();
}

void zo_pcg32_floats(uint64_t state, uint64_t sequence, uint32_t n, float* out) {
    zo::Generator g;
    g.start(state, sequence);
    for (uint32_t i = 0; i < n; ++i) out[i] = g.randomFloat();
}

}
