// reservoir.cuh -- the per-pixel multi-sample reservoir of the reference (Reservoir, src/rendering/reservoir.h:28-73) as
// register-resident state shared by the pass kernels: init, update, finish (W), store, and one stream entry of combine*.
//
// One thread owns one pixel: the weighted-reservoir stream of a pixel is sequential by definition (every update's accept
// test depends on the running wSum, reservoir.cpp:22-25), and with >= 2 M pixels per frame the grid is wide enough
// without intra-pixel parallelism (DESIGN.md 4 on why splitting a pixel over lanes does not pay).  Random draws are
// addressed by (pixel, stage, stream, counter) (include/romis_rng.h), so results do not depend on launch shape or bands.
//
// NT > 0: numSamplesInReservoir is the compile-time constant NT and sub-reservoirs live in registers (predicated writes
// instead of dynamic indexing).  NT == 0: runtime N <= 32, sub-reservoirs in local memory (generic fallback).
#pragma once
#include "device_common.cuh"

namespace romis {

// Loop over sub-reservoirs: compile-time trip count (fully unrolled, register-resident) when NT > 0,
// runtime trip count over local-memory arrays in the generic NT == 0 fallback.
#define ROMIS_FOR_SUB(j, NT, N) _Pragma("unroll") for (int j = 0; j < ((NT) > 0 ? (NT) : (N)); j++)

template <int NT> struct SubRes {
    static constexpr int CAP = NT > 0 ? NT : 32;
    uint32_t light[CAP]; float u[CAP], v[CAP], W[CAP], wSum[CAP];
    float chosen[CAP];  // chosenSampleWeights (reservoir.cpp:27): only R-OMIS reads it, dead code in every other kernel
    float pdf[CAP];     // target pdf of the held sample at THIS pixel, as evaluated when it was accepted (see res_finish)
    uint32_t M[CAP];
    uint32_t cnt[CAP];  // routed sums of the sources' M, saturating at 2^32-1 (= the clamp of the exact sum: res_take_counts)
};

// min(a + b, 2^32 - 1): a chain of these equals the exact (size_t) sum clamped once at the end, one register instead of two
__device__ __forceinline__ uint32_t sat_add_u32(uint32_t a, uint32_t b) { const uint32_t s = a + b; return s < a ? 0xffffffffu : s; }

// Reservoir::Reservoir (reservoir.h:29-32)
template <int NT> __device__ __forceinline__ void res_init(SubRes<NT>& r, int N) {
    ROMIS_FOR_SUB(j, NT, N) {
        r.light[j] = ROMIS_NO_LIGHT; r.u[j] = 0.0f; r.v[j] = 0.0f; r.W[j] = 0.0f; r.wSum[j] = FLT_MIN; r.M[j] = 1u; r.cnt[j] = 0u; r.pdf[j] = 0.0f; r.chosen[j] = 0.0f;
    }
}

// Reservoir::update (reservoir.cpp:10-32); returns the sub-reservoir that took the sample
template <int NT> __device__ __forceinline__ int res_update(SubRes<NT>& r, int N, uint32_t light, float u, float v, float pdf, float weight,
                                                            romis_stream_key rk, uint32_t& rc) {
    int idx = 0; float smallest = FLT_MAX;
    ROMIS_FOR_SUB(j, NT, N) { if (r.wSum[j] < smallest) { idx = j; smallest = r.wSum[j]; } }
    // A weight of exactly 0 (light behind the surface, occluded or empty source sample) leaves wSum unchanged and makes
    // the accept test `u < 0 / wSum` false for every u in [0, 1]: only the counters move.  Taking that case out also
    // keeps 0 / x off nvcc's slow IEEE-division path (FCHK rejects a zero numerator).
    const bool zero = weight == 0.0f;
    const uint32_t draw = rc++;                                     // reservoir.cpp:24: one rand() per update, always
    if (NT > 0) {
        // The chosen sub-reservoir differs from lane to lane: its wSum is selected into one register first so that the warp
        // runs ONE addition, division and accept test (a branch per sub-reservoir made it run them once per distinct idx),
        // then predicated writes put the results back (and keep the arrays in registers).
        float ws = r.wSum[0];
        ROMIS_FOR_SUB(j, NT, N) { if (j > 0 && j == idx) ws = r.wSum[j]; }
        bool accept = false;
        if (!zero) {
            const float rnd = romis_rand_to_unit(romis_rng_rand(rk, draw));
            ws += weight;
            accept = rnd < (weight / ws);
        }
        ROMIS_FOR_SUB(j, NT, N) {
            const bool me = j == idx;
            r.M[j] += me ? 1u : 0u;
            if (me) r.wSum[j] = ws;                                 // zero weight: ws is the value it already holds
            if (me && accept) { r.light[j] = light; r.u[j] = u; r.v[j] = v; r.pdf[j] = pdf; r.chosen[j] = weight; }
        }
        return idx;
    }
    r.M[idx] += 1u;                                                 // generic N: sub-reservoirs in local memory
    if (!zero) {
        const float rnd = romis_rand_to_unit(romis_rng_rand(rk, draw));
        r.wSum[idx] += weight;
        if (rnd < (weight / r.wSum[idx])) { r.light[idx] = light; r.u[idx] = u; r.v[idx] = v; r.pdf[idx] = pdf; r.chosen[idx] = weight; }
    }
    return idx;
}

template <int NT> __device__ __forceinline__ void res_store(const ResBuf& b, int lrow, int x, const SubRes<NT>& r, int N) {
    ROMIS_FOR_SUB(j, NT, N) {
        res_rec(b, lrow, j)[x] = make_uint4(r.light[j], __float_as_uint(r.u[j]), __float_as_uint(r.v[j]), __float_as_uint(r.W[j]));
        res_m(b, lrow, j)[x] = r.M[j];
        res_pdf(b, lrow, j)[x] = r.pdf[j];
    }
}

// targetPDF of the held sample y_j at this pixel (light.cpp:88, reservoir.cpp:59,98).  The reference evaluates it again
// here; the sample was accepted by an update whose weight was built from exactly that evaluation (same pixel, same light
// sample, a pure function), so the value kept at acceptance has the same bits.  A sub-reservoir that never accepted holds
// the default LightSample (position = colour = 0, reservoir.h:18-21), which is evaluated here.
template <int NT> __device__ __forceinline__ float res_held_pdf(const SubRes<NT>& r, int j, const PixCtx& c, bool es) {
    if (r.light[j] == ROMIS_NO_LIGHT) return target_pdf(c, es, V3(0, 0, 0), V3(0, 0, 0));
    return r.pdf[j];
}

// W_j = pdf == 0 ? 0 : (1/pdf) * (1/M_j) * wSum_j   (light.cpp:89-93, reservoir.cpp:57-65)
template <int NT> __device__ __forceinline__ void res_finish(SubRes<NT>& r, int N, const SceneDev& sc, const PixCtx& c, bool es) {
    ROMIS_FOR_SUB(j, NT, N) {
        float pdf = res_held_pdf(r, j, c, es);
        r.pdf[j] = pdf;
        float Wj = 0.0f;
        if (pdf != 0.0f) Wj = (1.0f / pdf) * (1.0f / (float)r.M[j]) * r.wSum[j];   // a real branch: 1/M with M = 0 stays unevaluated
        r.W[j] = Wj;
    }
}

// One stream entry of combineBiased / combineUnbiased (reservoir.cpp:42-53): w = (pdf * W_i) * float(M_i).
// own_pdf >= 0: the entry is this pixel's own reservoir of this frame, whose pdf at this pixel is stored with the record.
template <int NT> __device__ __forceinline__ void stream_sample(SubRes<NT>& r, int N, const SceneDev& sc, const PixCtx& c, bool es,
                                                                uint4 rec, uint32_t Mi, romis_stream_key rk, uint32_t& rc, float own_pdf = -1.0f) {
    float u = __uint_as_float(rec.y), v = __uint_as_float(rec.z), Wi = __uint_as_float(rec.w);
    float pdf;
    if (own_pdf >= 0.0f && rec.x != ROMIS_NO_LIGHT) pdf = own_pdf;
    else { v3 pos, col; light_sample(sc, rec.x, u, v, pos, col); pdf = target_pdf(c, es, pos, col); }
    int idx = res_update(r, N, rec.x, u, v, pdf, pdf * Wi * (float)Mi, rk, rc);
    if (NT > 0) { ROMIS_FOR_SUB(j, NT, N) { if (j == idx) r.cnt[j] = sat_add_u32(r.cnt[j], Mi); } }
    else r.cnt[idx] = sat_add_u32(r.cnt[idx], Mi);
}

// sampleNums = routed sums of the sources' M (reservoir.cpp:54,82); saturates at 2^32-1
template <int NT> __device__ __forceinline__ void res_take_counts(SubRes<NT>& r, int N) {
    ROMIS_FOR_SUB(j, NT, N) r.M[j] = r.cnt[j];
}

}  // namespace romis
