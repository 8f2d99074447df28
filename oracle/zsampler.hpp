// ORACLE — test infrastructure only (see zmath.hpp). Samplers of the reference:
//   Sobol (Owen-scrambled, 5-D padded)   src/core/sampler/sobol.zig:8-190
//   rnd.Generator PCG32                  src/base/random/generator.zig:1-47
//   Sampler union                        src/core/sampler/sampler.zig:17-74
// The direction numbers are not copied from sobol.zig:194-245: they are regenerated from the
// Joe-Kuo recurrence (new-joe-kuo-6.21201, dimensions 1-5) and pinned against the reference's table
// by tests/test_oracle_pins.py (tests/golden/sobol_directions.npy).
#pragma once

#include "zmath.hpp"

namespace zo {

struct Generator {  // generator.zig
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

struct SobolDirections {
    uint32_t d[5][32];

    SobolDirections() {
        // Joe & Kuo: degree s, polynomial coefficients a, initial m_i.
        static const uint32_t S[5]    = {0, 1, 2, 3, 3};
        static const uint32_t A[5]    = {0, 0, 1, 1, 2};
        static const uint32_t M[5][3] = {{0, 0, 0}, {1, 0, 0}, {1, 3, 0}, {1, 3, 1}, {1, 1, 1}};
        for (uint32_t i = 0; i < 32; ++i) d[0][i] = 1u << (31 - i);
        for (uint32_t j = 1; j < 5; ++j) {
            const uint32_t s = S[j];
            for (uint32_t i = 0; i < 32; ++i) {
                if (i < s) {
                    d[j][i] = M[j][i] << (31 - i);
                } else {
                    uint32_t v = d[j][i - s] ^ (d[j][i - s] >> s);
                    for (uint32_t k = 1; k < s; ++k) {
                        v ^= ((A[j] >> (s - 1 - k)) & 1u) * d[j][i - k];
                    }
                    d[j][i] = v;
                }
            }
        }
    }
};

inline const SobolDirections& sobolDirections() {
    static const SobolDirections dirs;
    return dirs;
}

inline uint32_t bitReverse(uint32_t x) {
    x = (x << 16) | (x >> 16);
    x = ((x & 0x00FF00FFu) << 8) | ((x & 0xFF00FF00u) >> 8);
    x = ((x & 0x0F0F0F0Fu) << 4) | ((x & 0xF0F0F0F0u) >> 4);
    x = ((x & 0x33333333u) << 2) | ((x & 0xCCCCCCCCu) >> 2);
    x = ((x & 0x55555555u) << 1) | ((x & 0xAAAAAAAAu) >> 1);
    return x;
}

inline uint32_t sobolHash(uint32_t i) {  // sobol.zig:107-124
    uint32_t x = i ^ (i >> 16);
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

inline uint32_t hashCombine(uint32_t seed, uint32_t v) {  // :126-128 (and the 4-lane form :130-134)
    return seed ^ (v + (seed << 6) + (seed >> 2));
}

inline uint32_t laineKarrasPermutation(uint32_t i, uint32_t seed) {  // :142-174
    uint32_t x = i ^ (i * 0x3d20adeau);
    x += seed;
    x *= (seed >> 16) | 1u;
    x ^= x * 0x05526c56u;
    x ^= x * 0x53a22864u;
    return x;
}

inline uint32_t nestedUniformScrambleBase2(uint32_t x, uint32_t seed) {  // :136-140
    uint32_t o = bitReverse(x);
    o          = laineKarrasPermutation(o, seed);
    return bitReverse(o);
}

inline void sobol5(uint32_t index, uint32_t out[5]) {  // :176-192
    const SobolDirections& dirs = sobolDirections();
    out[0] = out[1] = out[2] = out[3] = out[4] = 0;
    for (uint32_t bit = 0; bit < 32; ++bit) {
        const uint32_t mask = (index >> bit) & 1u;
        for (uint32_t k = 0; k < 5; ++k) out[k] ^= mask * dirs.d[k][bit];
    }
}

struct Sobol {  // sobol.zig:8-105
    float    buffer[5];
    uint32_t sample, dimension, start_seed, run_seed;

    void startPixel(uint32_t s, uint32_t seed) {
        sample                = s;
        dimension             = 5;
        const uint32_t hashed = sobolHash(seed);
        start_seed            = hashed;
        run_seed              = hashed;
    }
    void incrementSample() {
        sample += 1;
        dimension = 5;
        run_seed  = start_seed;
    }
    void incrementPadding() { dimension = 5; }

    void incrementSeed() {  // :36-60
        const float    S = 1.f / 4294967296.f;
        const uint32_t s = run_seed;
        const uint32_t i = nestedUniformScrambleBase2(sample, s);
        uint32_t       sob[5];
        sobol5(i, sob);
        for (uint32_t k = 0; k < 5; ++k) {
            const uint32_t hc  = hashCombine(s, k);
            const uint32_t nus = nestedUniformScrambleBase2(sob[k], hc);
            buffer[k]          = float(nus) * S;
        }
        run_seed  = sobolHash(s + 1);
        dimension = 0;
    }
    float sample1D() {
        if (dimension >= 5) incrementSeed();
        return buffer[dimension++];
    }
    void sample2D(float out[2]) {
        if (dimension >= 4) incrementSeed();
        const uint32_t d = dimension;
        dimension        = d + 2;
        out[0]           = buffer[d];
        out[1]           = buffer[d + 1];
    }
    Vec4f sample3D() {
        if (dimension >= 3) incrementSeed();
        const uint32_t d = dimension;
        dimension        = d + 3;
        return {{buffer[d], buffer[d + 1], buffer[d + 2], 0.f}};
    }
    Vec4f sample4D() {
        if (dimension >= 2) incrementSeed();
        const uint32_t d = dimension;
        dimension        = d + 4;
        return {{buffer[d], buffer[d + 1], buffer[d + 2], buffer[d + 3]}};
    }
};

struct Vec2f {
    float v[2];
    float operator[](int i) const { return v[i]; }
};

// Sampler union, sampler.zig:17-74. Random draws come from the worker's generator.
struct Sampler {
    bool       is_sobol;
    Sobol      sobol;
    Generator* rng;

    void startPixel(uint32_t s, uint32_t seed) {
        if (is_sobol) sobol.startPixel(s, seed);
    }
    void incrementSample() {
        if (is_sobol) sobol.incrementSample();
    }
    void incrementPadding() {
        if (is_sobol) sobol.incrementPadding();
    }
    float sample1D() { return is_sobol ? sobol.sample1D() : rng->randomFloat(); }
    Vec2f sample2D() {
        Vec2f r;
        if (is_sobol) {
            sobol.sample2D(r.v);
        } else {
            r.v[0] = rng->randomFloat();
            r.v[1] = rng->randomFloat();
        }
        return r;
    }
    Vec4f sample3D() {
        if (is_sobol) return sobol.sample3D();
        Vec4f r;
        r[0] = rng->randomFloat();
        r[1] = rng->randomFloat();
        r[2] = rng->randomFloat();
        r[3] = 0.f;
        return r;
    }
    Vec4f sample4D() {
        if (is_sobol) return sobol.sample4D();
        Vec4f r;
        r[0] = rng->randomFloat();
        r[1] = rng->randomFloat();
        r[2] = rng->randomFloat();
        r[3] = rng->randomFloat();
        return r;
    }
};

}  // namespace zo
