// ORACLE — test infrastructure only. rnd.Generator, src/base/random/generator.zig:1-47 (PCG32,
// XSH-RR, O'Neill 2014). Pinned by the published pcg32-demo vector (seed 42, stream 54).
#include "zyg_oracle.h"
#include "zmath.hpp"


namespace zo {

struct Generator {
    uint64_t state, inc;

    void start(uint64_t s, uint64_t sequence) {  // :13-20
        state = 0;
        inc   = (sequence << 1) | 1;
        randomUint();
        state += s;
        randomUint();
    }
    uint32_t randomUint() {  // :35-46
        const uint64_t old = state;
        state              = old * 6364136223846793005ull + inc;
        const uint32_t xrs = uint32_t(((old >> 18) ^ old) >> 27);
        const uint32_t rot = uint32_t(old >> 59);
        return (xrs >> rot) | (xrs << ((0u - rot) & 31));
    }
    float randomFloat() {  // :26-33
        uint32_t bits = randomUint();
        bits &= 0x007FFFFFu;
        bits |= 0x3F800000u;
        float f;
        std::memcpy(&f, &bits, 4);
        return f - 1.f;
    }
};

}  // namespace zo

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
