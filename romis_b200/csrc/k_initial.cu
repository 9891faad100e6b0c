// k_initial.cu -- initial RIS pass (genInitialSamples / genCanonicalSamples, reference src/scene/light.cpp:39-99).
#include "reservoir.cuh"
#include "launch.hpp"

namespace romis {

// ------------------------------------------------------------------------------------------------
// initial RIS: M candidates per pixel, + visibility reuse
// ------------------------------------------------------------------------------------------------
// EXTRA (R-OMIS): also writes wSums and chosenSampleWeights (reservoir.h:38-41), which
// arbitraryUnbiasedContributionWeightReciprocal reads (render_utils.cpp:245-257), as N planes each.
template <int NT, bool EXTRA>
__global__ void __launch_bounds__(ROMIS_LBT_INITIAL, ROMIS_MINB_INITIAL) initial_kernel(SceneDev sc, FrameDev fr, GBufDev g, ResBuf out, float* __restrict__ wsum, float* __restrict__ chosen) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    pdl_wait();                                     // the G-buffer comes from the kernel before this one
    pdl_launch_dependents();
    const int N = NT > 0 ? NT : (int)fr.f.numSamplesInReservoir;
    const bool es = fr.f.enableShading != 0;
    const uint32_t pixel = (uint32_t)y * (uint32_t)fr.W + (uint32_t)x;
    SubRes<NT> r; res_init(r, N);
    const size_t plane = (size_t)fr.W * fr.H;
    auto store_extra = [&]() {
        if (EXTRA) { ROMIS_FOR_SUB(j, NT, N) { wsum[(size_t)j * plane + pixel] = r.wSum[j]; chosen[(size_t)j * plane + pixel] = r.chosen[j]; } }
    };
    if (sc.n_lights == 0) { res_store(out, y - fr.ey0, x, r, N); store_extra(); return; }          // light.cpp:46: M_j stays 1
    PixCtx c = make_ctx(sc, fr, g, x, y);
    if (c.miss) {
        // every candidate weighs p^ / (1/L) = 0: wSums never leave FLT_MIN, so sub-reservoir 0 wins every strict-'<'
        // argmin (reservoir.cpp:12-19) and takes all M updates, nothing is ever accepted, W = 0
        ROMIS_FOR_SUB(j, NT, N) r.M[j] = 0u;
        r.M[0] = fr.f.initialLightSamples;
        res_store(out, y - fr.ey0, x, r, N);
        store_extra();
        return;
    }
    romis_stream_key ek = romis_rng_stream(fr.seed, fr.frame, fr.initial_stage, pixel, ROMIS_STREAM_ENGINE);
    romis_stream_key rk = romis_rng_stream(fr.seed, fr.frame, fr.initial_stage, pixel, ROMIS_STREAM_RAND);
    uint32_t rc = 0;
    ROMIS_FOR_SUB(j, NT, N) r.M[j] = 0u;               // light.cpp:58-60
    const float nLights = (float)sc.n_lights, invPdf = 1.0f / nLights;              // light.cpp:80
    // L a power of two (512, 65 536, 2^20 in the BASELINE configs): 1/L is exact and pdf / 2^-k and pdf * 2^k are the same
    // real number, rounded (or overflowed) once either way: the division is a multiplication
    const bool pow2L = (sc.n_lights & (sc.n_lights - 1)) == 0 && sc.n_lights <= (1 << 24);
    const uint32_t Mcand = fr.f.initialLightSamples;
    for (uint32_t i = 0; i < Mcand; i++) {
        uint32_t li = (uint32_t)romis_rng_uniform_int(romis_rng_bits(ek, i), 0, sc.n_lights - 1);
        uint32_t type = __float_as_uint(__ldg(&sc.lights[6 * (size_t)li].x));
        float u = 0.0f, v = 0.0f;
        if (type != ROMIS_LIGHT_POINT) u = romis_rand_to_unit(romis_rng_rand(rk, rc++));        // light.cpp:20 / :28
        if (type == ROMIS_LIGHT_PARALLELOGRAM) v = romis_rand_to_unit(romis_rng_rand(rk, rc++)); // light.cpp:29
        v3 pos, col; light_sample<true>(sc, li, u, v, pos, col);
        float pdf = target_pdf(c, es, pos, col);
        float w = 0.0f;                                     // +0 / (1/L) = +0: keep 0 / x off the slow division path
        if (pdf != 0.0f) w = pow2L ? pdf * nLights : pdf / invPdf;
        res_update(r, N, li, u, v, pdf, w, rk, rc);
    }
    // light.cpp:85-95: visibility reuse zeroes W of occluded samples, otherwise W = (1/pdf)(1/M)wSum
    res_finish(r, N, sc, c, es);
    if (fr.f.initialSamplesVisibilityCheck) {
        ROMIS_FOR_SUB(j, NT, N) {
            if (r.W[j] == 0.0f) continue;                   // already 0 (pdf = 0 or nothing accepted): the ray cannot change it
            v3 pos, col; light_sample<true>(sc, r.light[j], r.u[j], r.v[j], pos, col);
            if (!visible(sc, c, pos)) r.W[j] = 0.0f;
        }
    }
    res_store(out, y - fr.ey0, x, r, N);
    store_extra();
}


void launch_initial(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& out,
                    float* wsum, float* chosen) {
    if (wsum && chosen) { ROMIS_DISPATCH_N(N, (launch_pdl(initial_kernel<NT, true>, grid, block, s, sc, fr, g, out, wsum, chosen))); }
    else { ROMIS_DISPATCH_N(N, (launch_pdl(initial_kernel<NT, false>, grid, block, s, sc, fr, g, out, (float*)nullptr, (float*)nullptr))); }
}
}  // namespace romis
