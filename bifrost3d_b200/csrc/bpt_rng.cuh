// Counter-based sampling: pcg2d pixel hash and hash-based Owen-scrambled 4D Sobol points.
// Integer arithmetic; must be bit-exact against the reference:
//   pcg2d                      extensions/OptiXRenderer/OptiXRenderer/RNG.h:127-144
//   cessen_owen_hash           RNG.h:150-157
//   PracticalScrambledSobol    RNG.h:238-293 (direction numbers RNG.h:39-75)
// The reference evaluates the Sobol matrix product with a 4x32 masked-XOR loop. The four
// generator matrices are fixed, so here each dimension is evaluated in closed form:
//   dim 0: identity matrix on reversed bits  -> brev(index)
//   dim 1: Pascal matrix mod 2 (columns v_k = v_{k-1} ^ (v_{k-1} >> 1)); evaluated by the
//          5-step GF(2) superset transform below.
//   dims 2,3: masked-XOR loop over direction numbers held in constant memory. The loop counter is
//          warp-uniform, so every read is a constant-cache broadcast. The numbers are not copied from
//          the reference: they are generated at compile time from the Joe-Kuo primitive polynomials
//          (dim 2: x^2+x+1, m = {1,3}; dim 3: x^3+x+1, m = {1,3,1}) and checked against the oracle.
// ---------------------------------------------------------------------------
// The arithmetic restated in this file follows Bifrost3D (https://github.com/papaboo/Bifrost3D), which carries this notice:
//   Copyright (C) Bifrost. See AUTHORS.txt for authors.
//   This program is open source and distributed under the New BSD License. See LICENSE.txt for more detail.
// The notice and the licence terms are reproduced in NOTICE.md at the root of this repository.
// ---------------------------------------------------------------------------
#pragma once
#include "bpt_math.cuh"

namespace bpt {

struct SobolDirections {
    uint32_t v[2][32];
};

// Joe-Kuo recurrence: m_k = 2 a_1 m_{k-1} ^ 4 a_2 m_{k-2} ^ ... ^ 2^s m_{k-s} ^ m_{k-s}; v_k = m_k << (31 - k).
constexpr SobolDirections make_sobol_directions() {
    SobolDirections d = {};
    { // dimension 2: degree s = 2, a = {1}
        uint32_t m[32] = {1u, 3u};
        for (int k = 2; k < 32; ++k)
            m[k] = (2u * m[k - 1]) ^ (4u * m[k - 2]) ^ m[k - 2];
        for (int k = 0; k < 32; ++k)
            d.v[0][k] = m[k] << (31 - k);
    }
    { // dimension 3: degree s = 3, a = {0, 1}
        uint32_t m[32] = {1u, 3u, 1u};
        for (int k = 3; k < 32; ++k)
            m[k] = (4u * m[k - 2]) ^ (8u * m[k - 3]) ^ m[k - 3];
        for (int k = 0; k < 32; ++k)
            d.v[1][k] = m[k] << (31 - k);
    }
    return d;
}

static __constant__ SobolDirections c_sobol_directions = make_sobol_directions();

BPT_HD uint2 pcg2d(uint32_t x, uint32_t y) {
    x = x * 1664525u + 1013904223u;
    y = y * 1664525u + 1013904223u;
    x += y * 1664525u;
    y += x * 1664525u;
    x ^= x >> 16u;
    y ^= y >> 16u;
    x += y * 1664525u;
    y += x * 1664525u;
    x ^= x >> 16u;
    y ^= y >> 16u;
    return make_uint2(x, y);
}

BPT_D uint32_t owen_hash(uint32_t x, uint32_t seed) {
    x ^= x * 0x3d20adeau;
    x += seed;
    x *= (seed >> 16) | 1u;
    x ^= x * 0x05526c56u;
    x ^= x * 0x53a22864u;
    return x;
}

BPT_D uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }

BPT_D uint32_t nested_uniform_scramble(uint32_t x, uint32_t seed) {
    return __brev(owen_hash(__brev(x), seed));
}

// Sobol dimension 1: out = XOR over set bits k of index of v_k, v_0 = 1<<31, v_k = v_{k-1} ^ (v_{k-1}>>1),
// i.e. v_k = (1+S)^k v_0 with S a one-bit right shift. The coefficients of (1+S)^k mod 2 are Pascal's
// triangle mod 2, and by Lucas' theorem C(k, j) is odd iff j is a bit-subset of k. Output bit j (from
// the top) is therefore the XOR of index bits k over all supersets k of j: a GF(2) zeta transform over
// the 5-bit bit-position, done in 5 masked shift-XOR steps.
BPT_HD uint32_t sobol_dim1_natural(uint32_t y) {
    y ^= (y >> 1) & 0x55555555u;
    y ^= (y >> 2) & 0x33333333u;
    y ^= (y >> 4) & 0x0f0f0f0fu;
    y ^= (y >> 8) & 0x00ff00ffu;
    y ^= (y >> 16) & 0x0000ffffu;
    return y;
}
BPT_D uint32_t sobol_dim1(uint32_t index) { return __brev(sobol_dim1_natural(index)); }

BPT_D uint32_t sobol_dim23(int d, uint32_t index) {
    uint32_t r = 0u;
#pragma unroll
    for (int bit = 0; bit < 32; ++bit)
        r ^= (0u - ((index >> bit) & 1u)) & c_sobol_directions.v[d][bit];
    return r;
}

BPT_CALL1 uint4 sobol_sample4ui(uint32_t accumulation_count, uint32_t pixel_hash, uint32_t dimension) {
    uint32_t seed = pcg2d(pixel_hash, dimension).x;
    uint32_t index = nested_uniform_scramble(accumulation_count, seed);
    uint4 xs;
    xs.x = __brev(index);
    xs.y = sobol_dim1(index);
    xs.z = sobol_dim23(0, index);
    xs.w = sobol_dim23(1, index);
    xs.x = nested_uniform_scramble(xs.x, hash_combine(seed, 0u));
    xs.y = nested_uniform_scramble(xs.y, hash_combine(seed, 1u));
    xs.z = nested_uniform_scramble(xs.z, hash_combine(seed, 2u));
    xs.w = nested_uniform_scramble(xs.w, hash_combine(seed, 3u));
    return xs;
}

BPT_D float4 sobol_sample4f(uint32_t accumulation_count, uint32_t pixel_hash, uint32_t dimension) {
    const float normalizer = 1.0f / 4294967296.0f;
    uint4 u = sobol_sample4ui(accumulation_count, pixel_hash, dimension);
    // uint -> float conversion rounds to nearest even exactly as the host's cast does.
    return make_float4(float(u.x) * normalizer, float(u.y) * normalizer, float(u.z) * normalizer, float(u.w) * normalizer);
}

// RngSamplingDimension, Types.h:422-427
enum : uint32_t { DIM_CAMERA = 0, DIM_NEE = 1, DIM_BSDF = 2, DIM_ROULETTE = 3 /* not used by the reference */, DIM_MAX = 8 };

BPT_D float4 path_rng_sample4f(uint32_t accumulation_count, uint32_t pixel_hash, uint32_t bounces, uint32_t sampling_dimension) {
    return sobol_sample4f(accumulation_count, pixel_hash, DIM_MAX * bounces + sampling_dimension);
}

} // namespace bpt
