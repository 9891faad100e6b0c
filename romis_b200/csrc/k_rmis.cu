// k_rmis.cu -- R-MIS mode (renderRMIS, reference src/rendering/render.cpp:64-119): the neighbour index grid
// (generateResampleIndicesGrid, src/rendering/neighbour_selection.cpp:107-122), the per-iteration gather over the k+1
// neighbourhood pixels with equal or balance-heuristic MIS weights (render.cpp:79-112; generalisedBalanceHeuristic,
// render_utils.cpp:179-187) and combineToScreen (render_utils.cpp:68-85).  The initial RIS of every iteration is
// initial_kernel (k_initial.cu) under the iteration's own random-stream stage.
#include <cstdlib>
#include "reservoir.cuh"
#include "launch.hpp"
#include "romis_cod.h"
#include "cod_fixed.hpp"

namespace romis {

// libstdc++ 13 uniform_int_distribution on a 32-bit engine: Lemire's method WITH its rejection step
// (bits/uniform_int_dist.h:255-281), as std::sample instantiates it inside the reference.  Value in [0, range).
__device__ __forceinline__ uint32_t lemire32(romis_stream_key ek, uint32_t& ec, uint32_t range) {
    uint64_t product = (uint64_t)romis_rng_bits(ek, ec++) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        const uint32_t threshold = (0u - range) % range;
        while (low < threshold) { product = (uint64_t)romis_rng_bits(ek, ec++) * (uint64_t)range; low = (uint32_t)product; }
    }
    return (uint32_t)(product >> 32);
}

// areSimilar (neighbour_selection.cpp:7-22); `own` is the canonical pixel (lhs).  A miss pixel keeps the value-initialised
// geometryId 0.  The normal test compares the dot product with the ANGLE in radians (:18), as the reference does.
__device__ __forceinline__ bool are_similar(const romis_rmis_params& rp, uint32_t n_meshes, float4 own, uint32_t own_mesh, float4 nb, uint32_t nb_mesh) {
    if (rp.neighbourSameGeometry) {
        uint32_t gl = own_mesh == n_meshes ? 0u : own_mesh, gr = nb_mesh == n_meshes ? 0u : nb_mesh;
        if (gl != gr) return false;
    }
    float depthFracDiff = fabsf(1.0f - (own.x / nb.x));
    if (depthFracDiff > rp.neighbourMaxDepthDifferenceFraction) return false;
    float normalsDot = dot3(V3(own.y, own.z, own.w), V3(nb.y, nb.z, nb.w));
    if (normalsDot < rp.neighbourMaxNormalAngleDifferenceRadians) return false;
    return true;
}

// The same predicate without early exits, for the window scan (441 evaluations per pixel at r = 10): the three criteria are
// evaluated side by side and combined, which costs a few arithmetic instructions more than the early exits save in branches.
__device__ __forceinline__ bool are_similar_flat(bool sameGeometry, float maxDepthFrac, float minNormalsDot, uint32_t n_meshes,
                                                 float4 own, uint32_t own_gid, float4 nb, uint32_t nb_mesh) {
    const uint32_t gr = nb_mesh == n_meshes ? 0u : nb_mesh;
    const float depthFracDiff = fabsf(1.0f - (own.x / nb.x));
    const float normalsDot = dot3(V3(own.y, own.z, own.w), V3(nb.y, nb.z, nb.w));
    // the negations keep the reference's behaviour for NaNs: `a > b` and `a < b` are both false for a NaN
    return (!sameGeometry || own_gid == gr) & !(depthFracDiff > maxDepthFrac) & !(normalsDot < minNormalsDot);
}

// The window of one pixel, classified: bit i of `mask` (scan order over the rows of the clipped window, wx pixels per row) =
// "similar"; the pixel's own bit stays clear and `own` remembers where it is (it belongs to neither class).
struct RmisWindow { int x0, y0, wx, count, own; uint32_t mask[ROMIS_RMIS_WORDS]; };

// word `wi` of the membership mask of class `cls` (true: similar, false: dissimilar = clear bits inside the window, own pixel excluded)
__device__ __forceinline__ uint32_t class_word(const RmisWindow& w, bool cls, int wi) {
    uint32_t v = w.mask[wi];
    if (cls) return v;
    v = ~v;
    const int rest = w.count - wi * 32;                                 // window pixels from this word on
    if (rest < 32) v &= (1u << rest) - 1u;
    if ((w.own >> 5) == wi) v &= ~(1u << (w.own & 31));
    return v;
}

// Walks the members of a class in scan order: select(m) returns the packed (y << 16 | x) position of the m-th member, for
// non-decreasing m (the cursor only moves forward, so a whole class costs one pass over the mask words).
struct ClassCursor {
    int wi = 0; uint32_t base = 0;
    __device__ __forceinline__ uint32_t select(const RmisWindow& w, bool cls, uint32_t m) {
        uint32_t wv = class_word(w, cls, wi); uint32_t c = (uint32_t)__popc(wv);
        while (m - base >= c) { base += c; wi++; wv = class_word(w, cls, wi); c = (uint32_t)__popc(wv); }
        uint32_t t = wv;
        for (uint32_t q = m - base; q != 0u; q--) t &= t - 1u;          // drop the q lowest members of the word
        const int i = wi * 32 + (__ffs((int)t) - 1);
        // i / wx for i < 3721, wx <= 61: (i + 0.5) / wx is at least 0.5 / 61 away from an integer, far beyond the approximation's error
        const int row = (int)__fdividef((float)i + 0.5f, (float)w.wx);
        return ((uint32_t)(w.y0 + row) << 16) | (uint32_t)(w.x0 + (i - row * w.wx));
    }
};

// std::sample (libstdc++ selection sampling, bits/stl_algo.h:5841-5905) of n out of the `size` window pixels of class `cls`,
// in scan order, without materialising the list: one decision per member, two decisions per engine call while unsampled^2
// fits the 32-bit engine range (__gen_two_uniform_ints), in MEMBER space -- window positions are only looked up for the (at
// most k) members that are taken.  all = true copies the class without draws (neighbour_selection.cpp:80).  Appends packed
// (y << 16 | x) entries at plane `no` of the pixel's neighbour column.
__device__ __forceinline__ void emit_class(const RmisWindow& w, bool cls, uint32_t size, uint32_t n, bool all,
                                           romis_stream_key ek, uint32_t& ec, uint32_t* __restrict__ col, size_t plane, int& no, int cap) {
    if (size == 0u) return;
    if (n > size) n = size;
    if (all) n = size;
    if (n == 0u) return;
    // The decisions of different pixels fall at different members, so a lane that looked its pick up on the spot would do so
    // alone (measured: a third of the kernel's issue slots at 1.2 active lanes).  Picks are only noted (member index) and the
    // window positions are looked up after the sampling loop, by all lanes of the warp together.
    uint32_t taken[ROMIS_MAX_K + 1]; int nt = 0;
    const int room = cap - no;
    auto emit = [&](uint32_t m) { if (nt < room) taken[nt] = m; nt++; --n; };    // the grid holds k + 1 planes (host rejects what needs more)
    auto flush = [&]() {
        ClassCursor cur;
        const int cnt = nt < room ? nt : room;
        for (int t = 0; t < cnt; t++) col[(size_t)(no + t) * plane] = cur.select(w, cls, taken[t]);
        no += nt;
    };
    if (all) {
        ClassCursor cur;
        for (uint32_t m = 0; m < size; m++) { if (no < cap) col[(size_t)no * plane] = cur.select(w, cls, m); no++; }
        return;
    }
    uint32_t unsampled = size, m = 0;
    if (0xffffffffu / unsampled >= unsampled) {                         // two decisions per engine call (:5866-5885)
        while (n != 0u && unsampled >= 2u) {
            const uint32_t b1 = unsampled - 1u;
            const uint32_t xx = lemire32(ek, ec, unsampled * b1);
            // xx / b1 and xx % b1: xx < unsampled * b1 <= 3721 * 3720 < 2^24 (r <= ROMIS_RMIS_MAX_R), so both operands are
            // exact floats, the quotient is below 3721 and the approximate division is off by less than one: one fix-up step
            uint32_t p0 = (uint32_t)__fdividef((float)xx, (float)b1);
            int32_t rem = (int32_t)(xx - p0 * b1);
            if (rem < 0) { p0--; rem += (int32_t)b1; } else if (rem >= (int32_t)b1) { p0++; rem -= (int32_t)b1; }
            unsampled -= 2u;
            if (p0 < n) { emit(m); if (n == 0u) break; }
            if ((uint32_t)rem < n) emit(m + 1u);
            m += 2u;
        }
    }
    while (n != 0u && unsampled != 0u) {                                // one decision per engine call (:5888-5895)
        if (lemire32(ek, ec, unsampled) < n) emit(m);
        --unsampled; m++;
    }
    flush();
}

// generateResampleIndicesGrid: indicesRandom (neighbour_selection.cpp:24-45) / indicesSimilarity (:47-105)
__global__ void __launch_bounds__(256) rmis_neighbours_kernel(SceneDev sc, FrameDev fr, GBufDev g, RmisDev rm) {
    int x, y; thread_pixel<true>(x, y);
    y += fr.y0;                                     // the band's own rows; the window below may reach into its halo rows
    if (x >= fr.W || y >= fr.y1) return;
    const int k = (int)fr.f.numNeighboursToSample, r = (int)fr.f.spatialResampleRadius;
    const size_t p = (size_t)y * fr.W + x;          // R-MIS planes are indexed by the GLOBAL pixel, G-buffer and reservoirs by band row
    const size_t lp = (size_t)(y - fr.ey0) * fr.W + x;
    uint32_t* col = rm.nb + p;
    romis_stream_key ek = romis_rng_stream(fr.seed, fr.frame, ROMIS_STAGE_RMIS_NEIGH, (uint32_t)p, ROMIS_STREAM_ENGINE);
    uint32_t ec = 0;
    int no = 0;
    const int wx0 = max(0, x - r), wx1 = min(fr.W - 1, x + r), wy0 = max(0, y - r), wy1 = min(fr.H - 1, y + r);
    col[0] = ((uint32_t)y << 16) | (uint32_t)x; no = 1;                            // :40 / :71 the pixel itself, always
    if (rm.p.neighbourSelectionStrategy == ROMIS_NEIGHBOURS_RANDOM) {
        for (int i = 0; i < k; i++) {
            // `glm::ivec2(distrX(gen), distrY(gen))` (:42): the order of the two draws is unspecified in C++; g++, which
            // builds the compiled reference the oracle is pinned against, evaluates them right to left: y first
            int ny = romis_rng_uniform_int(romis_rng_bits(ek, ec++), wy0, wy1);
            int nx = romis_rng_uniform_int(romis_rng_bits(ek, ec++), wx0, wx1);
            col[(size_t)no * rm.plane] = ((uint32_t)ny << 16) | (uint32_t)nx; no++;
        }
    } else {
        // classify the window (:59-72): one bit per window pixel, 32 pixels per mask word
        RmisWindow w;
        w.x0 = wx0; w.y0 = wy0; w.wx = wx1 - wx0 + 1; w.count = w.wx * (wy1 - wy0 + 1); w.own = (y - wy0) * w.wx + (x - wx0);
        const float4 own = g.tn[lp]; const uint32_t own_mesh = g.mesh[lp];
        const uint32_t own_gid = own_mesh == (uint32_t)sc.n_meshes ? 0u : own_mesh;    // a miss pixel keeps the value-initialised geometryId 0
        const bool sameGeometry = rm.p.neighbourSameGeometry != 0;
        const float maxDepthFrac = rm.p.neighbourMaxDepthDifferenceFraction, minNormalsDot = rm.p.neighbourMaxNormalAngleDifferenceRadians;
        uint32_t ns = 0, word = 0; int bit = 0, wi = 0;
        for (int ny = wy0; ny <= wy1; ny++) {
            const float4* __restrict__ trow = g.tn + (size_t)(ny - fr.ey0) * fr.W + wx0;
            const uint32_t* __restrict__ mrow = g.mesh + (size_t)(ny - fr.ey0) * fr.W + wx0;
            _Pragma("unroll 3") for (int cx = 0; cx < w.wx; cx++) {
                const bool s = are_similar_flat(sameGeometry, maxDepthFrac, minNormalsDot, (uint32_t)sc.n_meshes, own, own_gid, trow[cx], mrow[cx]);
                word |= (s ? 1u : 0u) << bit;
                if (++bit == 32) { w.mask[wi++] = word; ns += (uint32_t)__popc(word); word = 0u; bit = 0; }
            }
        }
        if (bit) { w.mask[wi] = word; ns += (uint32_t)__popc(word); }
        {   // the pixel itself is in neither class (:64)
            uint32_t& ow = w.mask[w.own >> 5];
            if ((ow >> (w.own & 31)) & 1u) { ow &= ~(1u << (w.own & 31)); ns--; }
        }
        const uint32_t nd = (uint32_t)w.count - 1u - ns;
        const uint32_t ku = (uint32_t)k;
        if (rm.p.neighbourSelectionStrategy == ROMIS_NEIGHBOURS_SIMILAR) {          // :79-85
            if (ns < ku) {
                emit_class(w, true, ns, ns, true, ek, ec, col, rm.plane, no, rm.K1);
                emit_class(w, false, nd, ku - ns, false, ek, ec, col, rm.plane, no, rm.K1);
            } else emit_class(w, true, ns, ku, false, ek, ec, col, rm.plane, no, rm.K1);
        } else {                                                                    // EqualSimilarDissimilar :94-103
            uint32_t similarsSampled = min(ku / 2u + 1u, ns);
            const uint32_t desiredDissimilars = ku - similarsSampled;
            if (desiredDissimilars > nd) similarsSampled += ku - nd - similarsSampled;
            emit_class(w, true, ns, similarsSampled, false, ek, ec, col, rm.plane, no, rm.K1);
            emit_class(w, false, nd, ku - similarsSampled, false, ek, ec, col, rm.plane, no, rm.K1);
        }
    }
    for (; no < rm.K1; no++) col[(size_t)no * rm.plane] = 0xffffffffu;
}

// Hit point and unit view vector of every pixel, once per frame (GBufDev::pv): exactly make_ctx's arithmetic.
__global__ void __launch_bounds__(256) ctx_kernel(SceneDev sc, FrameDev fr, GBufDev g, int row0, int row1) {
    int x, y; thread_pixel<false>(x, y);
    y += row0;                                      // band rows plus halo rows
    if (x >= fr.W || y >= row1) return;
    GBufDev g0 = g; g0.pv = nullptr;
    const PixCtx c = make_ctx<false>(sc, fr, g0, x, y);
    const size_t p = (size_t)(y - fr.ey0) * fr.W + x;      // like the G-buffer (make_ctx<true> reads it back so)
    g.pv[2 * p] = make_float4(c.P.x, c.P.y, c.P.z, 0.0f);
    g.pv[2 * p + 1] = make_float4(c.Vv.x, c.Vv.y, c.Vv.z, 0.0f);
}

// One iteration's gather (render.cpp:79-112): every pixel shades the samples of its neighbourhood pixels at ITS OWN hit
// point, each weighted by the MIS weight and the sample's outputWeight, with a shadow ray per sample.
template <int NT>
__global__ void __launch_bounds__(256, ROMIS_MINB_GATHER) rmis_gather_kernel(SceneDev sc, FrameDev fr, GBufDev g, ResBuf in, RmisDev rm) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    const int N = NT > 0 ? NT : (int)fr.f.numSamplesInReservoir;
    const bool es = fr.f.enableShading != 0;
    const size_t p = (size_t)y * fr.W + x;
    PixCtx c = make_ctx<true>(sc, fr, g, x, y);
    // A miss pixel shades every sample to exactly +0 (see target_pdf) and the balance weight is 0 / (FLT_MIN + ...) = 0:
    // the iteration adds +0 to the accumulator, which leaves it unchanged.
    if (c.miss) return;
    uint32_t q[ROMIS_MAX_K + 1]; int n = 0;
    for (int a = 0; a < rm.K1; a++) { uint32_t e = rm.nb[(size_t)a * rm.plane + p]; if (e != 0xffffffffu) q[n++] = e; }
    const bool balance = rm.p.misWeightRMIS == ROMIS_MIS_BALANCE;
    const float equalWeight = 1.0f / (float)n;                                      // render.cpp:97 (size_t -> float)
    v3 finalColor = V3(0, 0, 0);
    for (int a = 0; a < n; a++) {
        const int qy = (int)(q[a] >> 16), qx = (int)(q[a] & 0xffffu);
        // The pixel's own reservoir under initialSamplesVisibilityCheck: the initial pass shot this very ray (same hit point, same
        // sample) and zeroed W when it was blocked (light.cpp:86-87), so W != 0 below says "visible" without a second ray.
        const bool own_checked = fr.f.initialSamplesVisibilityCheck != 0 && q[a] == q[0];
        _Pragma("unroll 1") for (int j = 0; j < N; j++) {
            const uint4 rec = res_rec(in, qy - fr.ey0, j)[qx];
            const float Wj = __uint_as_float(rec.w);
            // W == 0 contributes (+-0) whatever the weight and the visibility: the sum keeps its bits
            if (Wj == 0.0f) continue;
            v3 pos, col; light_sample<true>(sc, rec.x, __uint_as_float(rec.y), __uint_as_float(rec.z), pos, col);
            const v3 shading = compute_shading(c, es, pos, col);
            if (shading.x == 0.0f && shading.y == 0.0f && shading.z == 0.0f) continue;      // same: adds (+-0)
            float misWeight = equalWeight;
            if (balance) {                                                          // render_utils.cpp:179-187
                const float numerator = length3(shading);                           // targetPDF at this pixel
                float denominator = FLT_MIN;
                for (int b = 0; b < n; b++) {
                    const int by = (int)(q[b] >> 16), bx = (int)(q[b] & 0xffffu);
                    PixCtx cb = make_ctx<true>(sc, fr, g, bx, by);
                    denominator += target_pdf(cb, es, pos, col);
                }
                misWeight = numerator / denominator;
            }
            v3 sampleColor = (own_checked || visible(sc, c, pos)) ? shading : V3(0, 0, 0);  // render.cpp:103-105
            finalColor = add3(finalColor, div3(scale3(scale3(sampleColor, misWeight), Wj), (float)N));      // :106
        }
    }
    float4 acc = rm.acc[p];
    rm.acc[p] = make_float4(acc.x + finalColor.x, acc.y + finalColor.y, acc.z + finalColor.z, 0.0f);        // :111
}

// combineToScreen (render_utils.cpp:68-85): average over the iterations, tone map, Screen layout
__global__ void __launch_bounds__(256) rmis_combine_kernel(FrameDev fr, RmisDev rm, float* __restrict__ rgb) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    const float4 acc = rm.acc[(size_t)y * fr.W + x];
    v3 color = div3(V3(acc.x, acc.y, acc.z), (float)rm.p.maxIterationsMIS);
    if (fr.f.enableToneMapping) color = tone_map(color, fr.f);                      // tone_mapping.cpp:8-11
    size_t i = (size_t)(fr.H - 1 - y) * fr.W + x;
    rgb[3 * i] = color.x; rgb[3 * i + 1] = color.y; rgb[3 * i + 2] = color.z;
}

// ------------------------------------------------------------------------------------------------
// R-OMIS (renderROMIS, reference src/rendering/render.cpp:121-265)
// ------------------------------------------------------------------------------------------------
// One iteration's accumulation (render.cpp:148-222): for every sample (pixel a of the neighbourhood, sub-reservoir j) the
// column vector of all k+1 sampling techniques evaluated at that sample
// (arbitraryUnbiasedContributionWeightReciprocal, render_utils.cpp:245-257), scaled, goes into the pixel's technique
// matrix (outer product) and, weighted by the shaded sample, into the three contribution vectors.  Matrix and vectors
// live in global memory as planes over the pixels (coalesced read-modify-write); the accumulation order per element is
// the reference's: iterations, then a, then j.
// PROG (useProgressiveROMIS, render.cpp:133-139,160-170,188-200): the running estimate additionally takes, per pixel of the
// neighbourhood, the current alpha components and, per sample, (f - sum_b alpha_b w_b) / sum_b (N / (k+1)) w_b over the total
// sample count -- with the reference's INTEGER N / (k+1) (:139), i.e. a division by FLT_MIN whenever N < k + 1.
template <int NT, bool PROG>
__global__ void __launch_bounds__(256, ROMIS_MINB_RMIS) romis_accumulate_kernel(SceneDev sc, FrameDev fr, GBufDev g, ResBuf in, RmisDev rm) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    constexpr int CAP = SubRes<NT>::CAP;
    const int N = NT > 0 ? NT : (int)fr.f.numSamplesInReservoir;
    const int K1 = rm.K1;
    const bool es = fr.f.enableShading != 0;
    const size_t p = (size_t)y * fr.W + x;
    const float nLights = (float)sc.n_lights, invPdf = 1.0f / nLights;
    const bool pow2L = (sc.n_lights & (sc.n_lights - 1)) == 0 && sc.n_lights <= (1 << 24);     // see initial_kernel
    const float Nf = (float)fr.f.numSamplesInReservoir;
    PixCtx c = make_ctx<true>(sc, fr, g, x, y);
    uint32_t q[ROMIS_COD_MAX];
    // per distribution b and sub-reservoir j, once instead of once per sample: 1 / M and wSum - chosenSampleWeight
    // (render_utils.cpp:250-254)
    float invM[ROMIS_COD_MAX][CAP], wRest[ROMIS_COD_MAX][CAP];
    for (int a = 0; a < K1; a++) {
        q[a] = rm.nb[(size_t)a * rm.plane + p];
        const int by = (int)(q[a] >> 16), bx = (int)(q[a] & 0xffffu);
        const size_t bp = (size_t)by * fr.W + bx;
        ROMIS_FOR_SUB(j, NT, N) {
            invM[a][j] = 1.0f / (float)res_m(in, by - fr.ey0, j)[bx];
            wRest[a][j] = rm.wsum[(size_t)j * rm.plane + bp] - rm.chosen[(size_t)j * rm.plane + bp];
        }
    }
    float Al[PROG ? 3 * ROMIS_COD_MAX : 1];
    v3 fin = V3(0, 0, 0);
    const float invTotalSamples = 1.0f / (float)(int32_t)((uint32_t)K1 * fr.f.numSamplesInReservoir);          // :138,199
    const float fractionOfTotalSamples = (float)(int32_t)(fr.f.numSamplesInReservoir / (uint32_t)K1);           // :139
    if (PROG) {
        for (int i = 0; i < 3 * K1; i++) Al[i] = rm.alpha[(size_t)i * rm.plane + p];
        const float4 f4 = rm.acc[p]; fin = V3(f4.x, f4.y, f4.z);
    }
    for (int a = 0; a < K1; a++) {                                                  // render.cpp:165
        if (PROG) fin = add3(fin, V3(Al[0 * K1 + a], Al[1 * K1 + a], Al[2 * K1 + a]));     // :168-170
        const int ay = (int)(q[a] >> 16), ax = (int)(q[a] & 0xffffu);
        v3 spos[CAP], scol[CAP];
        ROMIS_FOR_SUB(j, NT, N) {
            const uint4 rec = res_rec(in, ay - fr.ey0, j)[ax];
            light_sample<true>(sc, rec.x, __uint_as_float(rec.y), __uint_as_float(rec.z), spos[j], scol[j]);
        }
        // The samples shaded at this pixel (:184-186), once: the value is also distribution 0's target pdf (plane 0 of the
        // neighbour grid is the pixel itself, neighbour_selection.cpp:40,71).
        v3 shade[CAP];
        _Pragma("unroll 1") for (int j = 0; j < N; j++) shade[j] = c.miss ? V3(0, 0, 0) : compute_shading(c, es, spos[j], scol[j]);
        float V[CAP][ROMIS_COD_MAX];                                                // colVecW of every sample of pixel a
        for (int b = 0; b < K1; b++) {                                              // :178-181, distribution b at all N samples
            const int by = (int)(q[b] >> 16), bx = (int)(q[b] & 0xffffu);
            PixCtx cb;
            const bool own = b == 0;
            // distribution a at pixel a's own samples: the target pdf the initial pass stored with the record (res_finish)
            const bool stored = b == a && sc.n_lights != 0;
            float im[CAP], wr[CAP];                                 // requested before the context so that the latencies overlap
            ROMIS_FOR_SUB(j, NT, N) { im[j] = invM[b][j]; wr[j] = wRest[b][j]; }
            if (!own && !stored) cb = make_ctx<true>(sc, fr, g, bx, by);
            _Pragma("unroll 1") for (int j = 0; j < N; j++) {      // one copy of the evaluation: the kernel is instruction-cache bound
                float w = 0.0f;
                float imj = im[0], wrj = wr[0];
                if (NT > 0) { ROMIS_FOR_SUB(jj, NT, N) { if (jj == j) { imj = im[jj]; wrj = wr[jj]; } } }
                else { imj = im[j]; wrj = wr[j]; }
                float pdf;
                if (own) pdf = c.miss ? 0.0f : length3(shade[j]);
                else if (stored) pdf = res_pdf(in, by - fr.ey0, j)[bx];
                else pdf = target_pdf(cb, es, spos[j], scol[j]);
                if (pdf != 0.0f) {                                                  // render_utils.cpp:248-256
                    const float mock = pow2L ? pdf * nLights : pdf / invPdf;
                    const float arbitraryWeight = (1.0f / pdf) * imj * (wrj + mock);
                    w = 1.0f / arbitraryWeight;
                }
                V[j][b] = w;
            }
        }
        v3 sampleColor[CAP]; float scaleFactor[CAP];
        _Pragma("unroll 1") for (int j = 0; j < N; j++) {                           // :173
            // a zero result adds (+-0) to the contribution vectors: no shadow ray.  (Reusing the initial pass's verdict for the
            // pixel's own samples, as the R-MIS gather does, costs this kernel more in registers than the two rays: 22.0 -> 22.5 ms.)
            v3 sc_j = V3(0, 0, 0);
            const v3 shading = shade[j];
            if (!(shading.x == 0.0f && shading.y == 0.0f && shading.z == 0.0f) && visible(sc, c, spos[j])) sc_j = shading;
            sampleColor[j] = sc_j;
            if (PROG) {                                                             // :190-200
                v3 sumAlphaProducts = V3(0, 0, 0); float sumSampleFractionProducts = FLT_MIN;
                for (int b = 0; b < K1; b++) {
                    sumAlphaProducts = add3(sumAlphaProducts, scale3(V3(Al[0 * K1 + b], Al[1 * K1 + b], Al[2 * K1 + b]), V[j][b]));
                    sumSampleFractionProducts += fractionOfTotalSamples * V[j][b];
                }
                fin = add3(fin, scale3(sub3(div3(sc_j, sumSampleFractionProducts), div3(sumAlphaProducts, sumSampleFractionProducts)), invTotalSamples));
            }
            float sf = FLT_MIN;                                                     // :203-205
            for (int b = 0; b < K1; b++) sf += Nf * V[j][b];
            sf = 1.0f / sf;
            for (int b = 0; b < K1; b++) V[j][b] *= sf;                             // :208
            scaleFactor[j] = sf;
        }
        // :209-214 for the N samples of pixel a in one pass over the system: every element is read once, takes its N terms
        // in sample order (the order the reference adds them in) and is written once.  The matrix is a sum of outer products
        // v v^T: element (b, i) receives the same products in the same order as (i, b), so only the upper triangle is kept;
        // the solve and the parity read-back mirror it.
        for (int i = 0; i < K1; i++) {
            // three elements of the row in flight at a time: they are separate planes in HBM, and one load-add-store per
            // element waited for each of them in turn (24.7 -> 23.1 ms per frame; requesting the contribution elements ahead
            // of the row as well, or a whole row through predicated slots, did not pay)
            for (int b = i; b < K1; b += 3) {
                const bool h1 = b + 1 < K1, h2 = b + 2 < K1;
                float* const e0 = rm.tech + (size_t)(i * K1 + b) * rm.plane + p;
                float* const e1 = e0 + rm.plane; float* const e2 = e1 + rm.plane;
                float t0 = *e0, t1 = h1 ? *e1 : 0.0f, t2 = h2 ? *e2 : 0.0f;
                ROMIS_FOR_SUB(j, NT, N) {
                    const float vi = V[j][i];
                    t0 += vi * V[j][b];
                    if (h1) t1 += vi * V[j][b + 1];
                    if (h2) t2 += vi * V[j][b + 2];
                }
                *e0 = t0; if (h1) *e1 = t1; if (h2) *e2 = t2;
            }
            float cx = rm.contrib[(size_t)(0 * K1 + i) * rm.plane + p], cy = rm.contrib[(size_t)(1 * K1 + i) * rm.plane + p],
                  cz = rm.contrib[(size_t)(2 * K1 + i) * rm.plane + p];
            ROMIS_FOR_SUB(j, NT, N) {
                const float scaleColVecConst = scaleFactor[j] * V[j][i];            // :211
                cx += sampleColor[j].x * scaleColVecConst; cy += sampleColor[j].y * scaleColVecConst; cz += sampleColor[j].z * scaleColVecConst;
            }
            rm.contrib[(size_t)(0 * K1 + i) * rm.plane + p] = cx; rm.contrib[(size_t)(1 * K1 + i) * rm.plane + p] = cy;
            rm.contrib[(size_t)(2 * K1 + i) * rm.plane + p] = cz;
        }
    }
    if (PROG) rm.acc[p] = make_float4(fin.x, fin.y, fin.z, 0.0f);
}

// The direct estimator's final step (render.cpp:233-262): three minimum-norm least-squares solves per pixel
// (solveSystem = completeOrthogonalDecomposition().solve, render_utils.h:52 -> include/romis_cod.h), component sums,
// tone mapping, Screen layout.
// alphas_only: the progressive estimator's per-iteration update of the alpha vectors (:160-164) instead of the image.
// Common tail of both solve kernels: component sums, tone mapping, Screen layout (render.cpp:247-262)
__device__ __forceinline__ void romis_write_pixel(const FrameDev& fr, float* __restrict__ rgb, int x, int y, const float* sum) {
    v3 color = V3(sum[0], sum[1], sum[2]);
    if (fr.f.enableToneMapping) color = tone_map(color, fr.f);                      // tone_mapping.cpp:8-11
    size_t i = (size_t)(fr.H - 1 - y) * fr.W + x;
    rgb[3 * i] = color.x; rgb[3 * i + 1] = color.y; rgb[3 * i + 2] = color.z;
}

__global__ void __launch_bounds__(128) romis_solve_kernel(FrameDev fr, RmisDev rm, float* __restrict__ rgb, int alphas_only) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    const int K1 = rm.K1;
    const size_t p = (size_t)y * fr.W + x;
    float b[ROMIS_COD_MAX], xs[ROMIS_COD_MAX];
    romis_cod cod;                                                                  // symmetric: straight into the factorisation's storage
    bool any = false;
    for (int i = 0; i < K1; i++)                                                    // upper triangle stored, see the accumulation
        for (int b = i; b < K1; b++) {
            const float t = rm.tech[(size_t)(i * K1 + b) * rm.plane + p];
            cod.qr[i * K1 + b] = cod.qr[b * K1 + i] = t;
            any |= t != 0.0f;
        }
    // An all-zero system (a pixel whose neighbourhood saw nothing: the misses, a quarter of the nightclub frame) has rank 0 -- every
    // column norm, pivot and tau is 0, no diagonal entry exceeds the threshold 0 -- and the solution is x = 0 whatever b is.
    if (any) romis_cod_factor(&cod, K1); else { cod.n = K1; cod.rank = 0; }
    float sum[3];
    for (int ch = 0; ch < 3; ch++) {
        if (any) for (int i = 0; i < K1; i++) b[i] = rm.contrib[(size_t)(ch * K1 + i) * rm.plane + p];
        romis_cod_solve(&cod, b, xs);
        if (alphas_only) { for (int i = 0; i < K1; i++) rm.alpha[(size_t)(ch * K1 + i) * rm.plane + p] = xs[i]; continue; }
        float s = 0.0f;
        for (int i = 0; i < K1; i++) s += xs[i];                                    // :247-252
        sum[ch] = s;
    }
    if (alphas_only) return;
    romis_write_pixel(fr, rgb, x, y, sum);
}

// The same for a system size known at compile time (K1 = 6: the reference's default k = 5): cod_fixed.hpp, the decomposition in
// registers.  Bit-identical to the generic kernel (tests/test_cod_fixed.py on the CPU, tests/test_gpu_romis.py against the oracle).
template <int K1>
__global__ void __launch_bounds__(128) romis_solve_fixed_kernel(FrameDev fr, RmisDev rm, float* __restrict__ rgb, int alphas_only) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    const size_t p = (size_t)y * fr.W + x;
    CodFixed<K1> cod;
    bool any = false;
    _Pragma("unroll") for (int i = 0; i < K1; i++) {
        _Pragma("unroll") for (int b = i; b < K1; b++) {
            const float t = rm.tech[(size_t)(i * K1 + b) * rm.plane + p];
            cod.qr[i][b] = t; cod.qr[b][i] = t;
            any |= t != 0.0f;
        }
    }
    float sum[3] = {0.0f, 0.0f, 0.0f};
    if (any) {
        cod_fixed_solve3<K1>(cod,
            [&](int ch, float* b) { _Pragma("unroll") for (int i = 0; i < K1; i++) b[i] = rm.contrib[(size_t)(ch * K1 + i) * rm.plane + p]; },
            [&](int ch, const float* xs) {
                if (alphas_only) { _Pragma("unroll") for (int i = 0; i < K1; i++) rm.alpha[(size_t)(ch * K1 + i) * rm.plane + p] = xs[i]; return; }
                float s = 0.0f;
                _Pragma("unroll") for (int i = 0; i < K1; i++) s += xs[i];              // :247-252
                if (ch == 0) sum[0] = s; else if (ch == 1) sum[1] = s; else sum[2] = s;
            });
    } else if (alphas_only) {                       // all-zero system: rank 0, x = 0 (see the generic kernel)
        for (int i = 0; i < 3 * K1; i++) rm.alpha[(size_t)i * rm.plane + p] = 0.0f;
    }
    if (alphas_only) return;
    romis_write_pixel(fr, rgb, x, y, sum);
}

void launch_ctx(cudaStream_t s, dim3 grid, dim3 block, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, int row0, int row1) {
    ctx_kernel<<<grid, block, 0, s>>>(sc, fr, g, row0, row1);
}
void launch_rmis_neighbours(cudaStream_t s, dim3 grid, dim3 block, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const RmisDev& rm) {
    rmis_neighbours_kernel<<<grid, block, 0, s>>>(sc, fr, g, rm);
}
void launch_rmis_gather(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& in, const RmisDev& rm) {
    ROMIS_DISPATCH_N(N, (rmis_gather_kernel<NT><<<grid, block, 0, s>>>(sc, fr, g, in, rm)));
}
void launch_romis_accumulate(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& in, const RmisDev& rm) {
    if (rm.p.useProgressiveROMIS) { ROMIS_DISPATCH_N(N, (romis_accumulate_kernel<NT, true><<<grid, block, 0, s>>>(sc, fr, g, in, rm))); }
    else { ROMIS_DISPATCH_N(N, (romis_accumulate_kernel<NT, false><<<grid, block, 0, s>>>(sc, fr, g, in, rm))); }
}
void launch_romis_solve(cudaStream_t s, dim3 grid, dim3 block, const FrameDev& fr, const RmisDev& rm, float* rgb, bool alphas_only) {
    dim3 b(32, 4), gr((fr.W + 31) / 32, (fr.y1 - fr.y0 + 3) / 4);
    // ROMIS_SOLVE_GENERIC=1 (environment): the local-memory routine for every size (A/B runs)
    static const bool generic_only = [] { const char* e = std::getenv("ROMIS_SOLVE_GENERIC"); return e && std::atoi(e) != 0; }();
    if (rm.K1 == 6 && !generic_only) romis_solve_fixed_kernel<6><<<gr, b, 0, s>>>(fr, rm, rgb, alphas_only ? 1 : 0);
    else romis_solve_kernel<<<gr, b, 0, s>>>(fr, rm, rgb, alphas_only ? 1 : 0);
}
void launch_rmis_combine(cudaStream_t s, dim3 grid, dim3 block, const FrameDev& fr, const RmisDev& rm, float* rgb) {
    rmis_combine_kernel<<<grid, block, 0, s>>>(fr, rm, rgb);
}
}  // namespace romis
