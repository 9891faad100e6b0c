/*
 * romis_rng.h -- the counter-based random stream both sides are driven by.
 *
 * The reference draws from three unsynchronised sources:
 *   R1  std::mt19937 re-seeded per pixel + std::uniform_int_distribution  (src/scene/light.cpp:49-51,66)
 *   R2  rand() in the light samplers                                      (src/scene/light.cpp:20,28-29)
 *   R3  rand() in Reservoir::update                                       (src/rendering/reservoir.cpp:24)
 *   R4  one shared std::mt19937 + uniform_int_distribution(-r, r)         (src/rendering/render_utils.cpp:89-91,109-110)
 * which makes it non-deterministic.  north_star's parity mode injects ONE counter-based generator
 * into the reference through a harness; this header is that generator.  Every draw is addressed by
 *
 *     (seed, frame, stage, pixel, stream, counter)
 *
 *   stage   ROMIS_STAGE_INITIAL | ROMIS_STAGE_TEMPORAL | ROMIS_STAGE_SPATIAL0 + pass
 *   pixel   y * W + x in GLOBAL image coordinates (so row-band shards reproduce the 1-GPU stream)
 *   stream  ROMIS_STREAM_ENGINE: the draws the reference takes from a std:: engine (R1, R4)
 *           ROMIS_STREAM_RAND:   the draws the reference takes from rand()      (R2, R3)
 *   counter position of the draw within (pixel, stage, stream), in the reference's program order:
 *           initial:  ENGINE c = candidate index;  RAND c = running count of rand() calls for the
 *                     pixel (0/1/2 sampler draws, then 1 update draw, per candidate)
 *           temporal: RAND c = 0 .. 2N-1 (one per Reservoir::update)
 *           spatial:  ENGINE c = 2*neighbour + {0: dx, 1: dy};  RAND c = running update count
 *
 * Mappings (restated on both sides, see oracle/ref_harness/rng_shim.h):
 *   rand()                    -> bits >> 1                     (RAND_MAX = 2^31 - 1, as glibc)
 *   uniform float (A.2)       -> float(rand()) / 2147483648.0f (linearMap, src/utils/utils.cpp:26-31)
 *   uniform_int_distribution  -> a + ((u64(bits) * (b - a + 1)) >> 32)   (multiply-shift; libstdc++'s
 *                                Lemire rejection step is dropped so every draw costs exactly one
 *                                engine call -- it would redraw with probability < range / 2^32)
 */
#ifndef ROMIS_RNG_H
#define ROMIS_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define ROMIS_RNG_HD __host__ __device__ __forceinline__
#else
#define ROMIS_RNG_HD static inline
#endif

enum { ROMIS_STAGE_INITIAL = 0, ROMIS_STAGE_TEMPORAL = 1, ROMIS_STAGE_SPATIAL0 = 2 /* + pass, < 64 */,
       /* R-MIS (renderRMIS, reference src/rendering/render.cpp:64-119) */
       ROMIS_STAGE_RMIS_NEIGH = 64,         /* neighbour index grid: ENGINE c = running count of engine calls of the pixel */
       ROMIS_STAGE_RMIS_INITIAL0 = 128      /* + iteration: genInitialSamples of that iteration, counters as ROMIS_STAGE_INITIAL */ };
enum { ROMIS_STREAM_ENGINE = 0, ROMIS_STREAM_RAND = 1 };

/* 32-bit finaliser (two odd multipliers, three xorshifts) */
ROMIS_RNG_HD uint32_t romis_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

/* 64-bit key of a (seed, frame, stage, pixel, stream) stream, as two 32-bit words. */
typedef struct romis_stream_key { uint32_t k0, k1; } romis_stream_key;

ROMIS_RNG_HD romis_stream_key romis_rng_stream(uint64_t seed, uint32_t frame, uint32_t stage,
                                              uint32_t pixel, uint32_t stream) {
    uint32_t s0 = (uint32_t)seed, s1 = (uint32_t)(seed >> 32);
    uint32_t a = romis_mix32(s0 ^ 0x9e3779b9u);
    a = romis_mix32(a ^ (frame * 0x85ebca6bu + 0x165667b1u));
    a = romis_mix32(a ^ ((stage * 2u + stream) * 0xc2b2ae35u + 0x27d4eb2fu));
    uint32_t b = romis_mix32(s1 ^ 0x7f4a7c15u ^ a);
    b = romis_mix32(b + pixel * 0x9e3779b1u);
    a = romis_mix32(a ^ pixel ^ (b >> 7));
    romis_stream_key k; k.k0 = a; k.k1 = b;
    return k;
}

/* The counter-th 32-bit draw of a stream: one finaliser over k0 + counter * golden, with the second key
 * word folded in between its two multiply rounds. */
ROMIS_RNG_HD uint32_t romis_rng_bits(romis_stream_key k, uint32_t counter) {
    uint32_t x = k.k0 + counter * 0x9e3779b1u;
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= k.k1;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

/* rand() replacement: [0, 2^31 - 1] */
ROMIS_RNG_HD int32_t romis_rng_rand(romis_stream_key k, uint32_t counter) {
    return (int32_t)(romis_rng_bits(k, counter) >> 1);
}

/* linearMap(float(rand()), 0, RAND_MAX, 0, 1) with RAND_MAX = 2147483647 -> float 2147483648.0f */
ROMIS_RNG_HD float romis_rand_to_unit(int32_t r) {
#if defined(__CUDA_ARCH__)
    /* Same bits with one multiplication: r >= 0, so float(r) is +0 or >= 1; dividing by 2^31 and multiplying by 2^-31 are
     * both exact; x - 0, x * 1 and, for x >= +0, x + 0 return x.  (nvcc keeps the `+ 0.0f` otherwise: it cannot know x != -0.) */
    return (float)r * 4.656612873077392578125e-10f;
#else
    return (((float)r - 0.0f) / (2147483648.0f - 0.0f)) * (1.0f - 0.0f) + 0.0f;
#endif
}

/* uniform integer in [a, b] from one 32-bit draw */
ROMIS_RNG_HD int32_t romis_rng_uniform_int(uint32_t bits, int32_t a, int32_t b) {
    uint32_t range = (uint32_t)(b - a) + 1u;
    return a + (int32_t)(((uint64_t)bits * (uint64_t)range) >> 32);
}

#endif /* ROMIS_RNG_H */
