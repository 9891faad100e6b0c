// k_spatial.cu -- spatial reuse pass (spatialReuse, reference src/rendering/render_utils.cpp:87-140).
#include "reservoir.cuh"
#include "launch.hpp"

namespace romis {

// ------------------------------------------------------------------------------------------------
// spatial reuse, one pass: k neighbours from a (2r+1)^2 window of the previous iteration, self last
// ------------------------------------------------------------------------------------------------
// ES: enableShading is known to be on (the reference's default): the flag and the material's kd, which only the unshaded
// path reads, stop occupying registers in a kernel that is short of them (measured: -3.5 %).  !ES reads the flag at run time.
template <int NT, bool UNBIASED, bool ES>
__device__ __forceinline__ void spatial_pixel(const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& in, const ResBuf& out, int pass,
                                              int x, int y, const HaloDev* hd, const FineDev& fd, int by0, int by1) {
    const int N = NT > 0 ? NT : (int)fr.f.numSamplesInReservoir;
    const bool es = ES || fr.f.enableShading != 0;
    const int lrow = y - fr.ey0;
    const uint32_t pixel = (uint32_t)y * (uint32_t)fr.W + (uint32_t)x;
    constexpr int CAP = SubRes<NT>::CAP;
    const int k = (int)fr.f.numNeighboursToSample, rad = (int)fr.f.spatialResampleRadius;
    PixCtx c = make_ctx(sc, fr, g, x, y);
    romis_stream_key ek = romis_rng_stream(fr.seed, fr.frame, ROMIS_STAGE_SPATIAL0 + (uint32_t)pass, pixel, ROMIS_STREAM_ENGINE);
    romis_stream_key rk = romis_rng_stream(fr.seed, fr.frame, ROMIS_STAGE_SPATIAL0 + (uint32_t)pass, pixel, ROMIS_STREAM_RAND);
    uint32_t rc = 0;
    SubRes<NT> r; res_init(r, N);
    // Phase 1 -- pick the stream: k draws (x before y, always 2k draws, render_utils.cpp:109-110), clamp to the image, and in
    // biased mode apply the hard-coded depth / normal heuristics (:114-118).  Only the 16-byte {t, n} of each neighbour is
    // touched here and the iterations are independent, so their gathers overlap.  Self goes last (:124).
    uint32_t stream[ROMIS_MAX_K + 1];   // (local row << 16) | x  -- width and band height are validated to fit 16 bits
    int ns = 0;
    const float4 own_tn = g.tn[(size_t)lrow * fr.W + x];
    for (int nb = 0; nb < k; nb++) {
        int dx = romis_rng_uniform_int(romis_rng_bits(ek, 2u * nb), -rad, rad);
        int dy = romis_rng_uniform_int(romis_rng_bits(ek, 2u * nb + 1u), -rad, rad);
        int nx = min(max(x + dx, 0), fr.W - 1);
        int ny = min(max(y + dy, 0), fr.H - 1);
        int nrow = ny - fr.ey0;
        bool keep = true;
        if (!UNBIASED) {
            float4 ntn = g.tn[(size_t)nrow * fr.W + nx];
            float depthFracDiff = fabsf(1.0f - (ntn.x / own_tn.x));
            float normalsDot = dot3(V3(ntn.y, ntn.z, ntn.w), V3(own_tn.y, own_tn.z, own_tn.w));
            keep = !(depthFracDiff > 0.1f || normalsDot < 0.90630778703f);
        }
        if (keep) stream[ns++] = ((uint32_t)nrow << 16) | (uint32_t)nx;
    }
    // Everything above read the G-buffer only, which is older than the previous kernel; `in` is the previous kernel's output.
    fine_wait(fd, by0, by1);
    pdl_launch_dependents();
    // Phase 2 -- stream the selected reservoirs through Reservoir::update in order (reservoir.cpp:42-53); the records of
    // entry s+1 are fetched while entry s is being evaluated.  Self goes last (:124) and outside the loop: every lane of
    // the warp reaches it together, and its target pdf at this pixel is stored with the record, so the warp skips the
    // evaluation as a whole.
    uint4 rec[CAP], nrec[CAP]; uint32_t Mi[CAP], nMi[CAP];
    if (ns > 0) {
        int srow = (int)(stream[0] >> 16), sx = (int)(stream[0] & 0xffffu);
        ROMIS_FOR_SUB(j, NT, N) { nrec[j] = res_rec(in, srow, j)[sx]; nMi[j] = res_m(in, srow, j)[sx]; }
    }
    for (int s = 0; s < ns; s++) {
        ROMIS_FOR_SUB(j, NT, N) { rec[j] = nrec[j]; Mi[j] = nMi[j]; }
        if (s + 1 < ns) {
            int srow = (int)(stream[s + 1] >> 16), sx = (int)(stream[s + 1] & 0xffffu);
            ROMIS_FOR_SUB(j, NT, N) { nrec[j] = res_rec(in, srow, j)[sx]; nMi[j] = res_m(in, srow, j)[sx]; }
        }
        ROMIS_FOR_SUB(j, NT, N) stream_sample(r, N, sc, c, es, rec[j], Mi[j], rk, rc);
    }
    ROMIS_FOR_SUB(j, NT, N) { rec[j] = res_rec(in, lrow, j)[x]; Mi[j] = res_m(in, lrow, j)[x]; }
    ROMIS_FOR_SUB(j, NT, N) stream_sample(r, N, sc, c, es, rec[j], Mi[j], rk, rc, res_pdf(in, lrow, j)[x]);
    stream[ns++] = ((uint32_t)lrow << 16) | (uint32_t)x;       // the unbiased normalisation below walks the whole stream
    res_take_counts(r, N);
    if (!UNBIASED) {
        res_finish(r, N, sc, c, es);
    } else {
        // reservoir.cpp:84-103: Z_j = sum of totalM(s) over stream reservoirs s whose OWN pixel sees y_j with pdf > 0
        uint64_t Z[CAP];
        v3 spos[CAP], scol[CAP];
        ROMIS_FOR_SUB(j, NT, N) { Z[j] = 0ull; light_sample(sc, r.light[j], r.u[j], r.v[j], spos[j], scol[j]); }
        for (int s = 0; s < ns; s++) {
            int srow = (int)(stream[s] >> 16), sx = (int)(stream[s] & 0xffffu);
            int sy = srow + fr.ey0;
            PixCtx cs = make_ctx(sc, fr, g, sx, sy);
            uint64_t tot = 0;
            ROMIS_FOR_SUB(j, NT, N) tot += res_m(in, srow, j)[sx];
            ROMIS_FOR_SUB(j, NT, N) {
                float pdf = target_pdf(cs, es, spos[j], scol[j]);
                // reservoir.cpp:89-92: pdf *= visibility; pdf > 0 counts.  The ray only matters when pdf > 0.
                if (pdf > 0.0f && (!fr.f.spatialReuseVisibilityCheck || visible(sc, cs, spos[j]))) Z[j] += tot;
            }
        }
        ROMIS_FOR_SUB(j, NT, N) {
            float pdf = res_held_pdf(r, j, c, es);
            r.pdf[j] = pdf;
            r.W[j] = (pdf == 0.0f || Z[j] == 0ull) ? 0.0f : (1.0f / pdf) * (1.0f / (float)Z[j]) * r.wSum[j];
        }
    }
    res_store(out, lrow, x, r, N);
    if (hd && hd->push) {
        // this pixel's row is a boundary row of the band: the neighbouring band reads it as a halo row in the next pass
        // (record and M only: the stored pdf is read for a band's own pixels alone)
        _Pragma("unroll") for (int e = 0; e < 2; e++) {
            if (!hd->peer_out[e] || !(e == 0 ? y < fr.y0 + hd->r : y >= fr.y1 - hd->r)) continue;
            ResBuf pb; pb.base = hd->peer_out[e]; pb.row_stride = hd->peer_stride[e]; pb.W = out.W; pb.N = out.N;
            const int prow = y - hd->peer_ey0[e];
            ROMIS_FOR_SUB(j, NT, N) {
                res_rec(pb, prow, j)[x] = make_uint4(r.light[j], __float_as_uint(r.u[j]), __float_as_uint(r.v[j]), __float_as_uint(r.W[j]));
                res_m(pb, prow, j)[x] = r.M[j];
            }
        }
    }
}

template <int NT, bool UNBIASED, bool ES>
__global__ void __launch_bounds__(256, ROMIS_MINB_SPATIAL) spatial_kernel(SceneDev sc, FrameDev fr, GBufDev g, ResBuf in, ResBuf out, int pass, FineDev fd) {
    int x, y; thread_pixel<true>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    const int by0 = fr.y0 + (int)(blockIdx.y * blockDim.y), by1 = by0 + (int)blockDim.y;
    spatial_pixel<NT, UNBIASED, ES>(sc, fr, g, in, out, pass, x, y, nullptr, fd, by0, by1);
    fine_signal(fd, by0, by1);
}

// The same pass for a band with peer-mapped neighbours (fused halo exchange).  Row groups next to a band edge are launched
// first: each waits until the neighbour's token of the PREVIOUS stage has arrived (its boundary rows are in my halo, and it has
// finished reading the halo rows of the buffer I am about to write into), runs, stores its rows twice (here and into the
// neighbour's halo) and the last such block of an edge publishes this stage's token.  Interior row groups never wait.
template <int NT, bool UNBIASED, bool ES>
__global__ void __launch_bounds__(256, ROMIS_MINB_SPATIAL) spatial_halo_kernel(SceneDev sc, FrameDev fr, GBufDev g, ResBuf in, ResBuf out, int pass, HaloDev hd, FineDev fd) {
    int by = (int)blockIdx.y;
    if (by >= hd.nl) by = by < hd.nl + hd.nh ? hd.gh0 + (by - hd.nl) : hd.nl + (by - hd.nl - hd.nh);
    const int gy0 = fr.y0 + by * (int)blockDim.y, gy1 = gy0 + (int)blockDim.y;
    const bool edge0 = hd.wait_flag[0] != nullptr && gy0 < fr.y0 + hd.r;
    const bool edge1 = hd.wait_flag[1] != nullptr && gy1 > fr.y1 - hd.r;
    if (edge0 || edge1) {
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            if (edge0) halo_spin(hd.wait_flag[0], hd.wait_token, hd.err);
            if (edge1) halo_spin(hd.wait_flag[1], hd.wait_token, hd.err);
            fence_acquire_sys();                    // the neighbour's rows in my halo: no stale line of them may stay in this SM's L1
        }
        __syncthreads();
    }
    int x, y; thread_pixel<true>(x, y, by);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;            // the barriers below are among the live threads
    spatial_pixel<NT, UNBIASED, ES>(sc, fr, g, in, out, pass, x, y, &hd, fd, gy0, gy1);
    fine_signal(fd, gy0, gy1);
    if (edge0 || edge1) {
        // Release pattern: every thread's stores (own buffer and the neighbour's halo over NVLink), the barrier, then ONE release
        // fence in thread 0 -- cumulative over what the barrier made it observe -- and its counter update; the last block of an edge
        // reads all the others' updates, fences both ways and publishes the token.  (Round 2: this was a fence.sc.sys by every
        // thread of every edge block, each one also invalidating the L1 under the two other blocks resident on the SM.)
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            fence_release_sys();
            _Pragma("unroll") for (int e = 0; e < 2; e++) {
                if (!(e == 0 ? edge0 : edge1)) continue;
                if (atomicAdd(&hd.counter[e], 1u) == hd.edge_blocks[e] - 1u) {
                    hd.counter[e] = 0u;
                    __threadfence_system();
                    *(volatile uint32_t*)hd.sig_flag[e] = hd.token;
                    __threadfence_system();
                }
            }
        }
    }
}


void launch_spatial(cudaStream_t s, dim3 grid, dim3 block, int N, bool unbiased, const SceneDev& sc, const FrameDev& fr, const GBufDev& g,
                    const ResBuf& in, const ResBuf& out, int pass, const FineDev& fd) {
    const bool es = fr.f.enableShading != 0;
    if (unbiased && es) { ROMIS_DISPATCH_N(N, (launch_pdl(spatial_kernel<NT, true, true>, grid, block, s, sc, fr, g, in, out, pass, fd))); }
    else if (unbiased) { ROMIS_DISPATCH_N(N, (launch_pdl(spatial_kernel<NT, true, false>, grid, block, s, sc, fr, g, in, out, pass, fd))); }
    else if (es) { ROMIS_DISPATCH_N(N, (launch_pdl(spatial_kernel<NT, false, true>, grid, block, s, sc, fr, g, in, out, pass, fd))); }
    else { ROMIS_DISPATCH_N(N, (launch_pdl(spatial_kernel<NT, false, false>, grid, block, s, sc, fr, g, in, out, pass, fd))); }
}

void launch_spatial_halo(cudaStream_t s, dim3 grid, dim3 block, int N, bool unbiased, const SceneDev& sc, const FrameDev& fr, const GBufDev& g,
                         const ResBuf& in, const ResBuf& out, int pass, const HaloDev& hd, const FineDev& fd) {
    const bool es = fr.f.enableShading != 0;
    if (unbiased && es) { ROMIS_DISPATCH_N(N, (launch_pdl(spatial_halo_kernel<NT, true, true>, grid, block, s, sc, fr, g, in, out, pass, hd, fd))); }
    else if (unbiased) { ROMIS_DISPATCH_N(N, (launch_pdl(spatial_halo_kernel<NT, true, false>, grid, block, s, sc, fr, g, in, out, pass, hd, fd))); }
    else if (es) { ROMIS_DISPATCH_N(N, (launch_pdl(spatial_halo_kernel<NT, false, true>, grid, block, s, sc, fr, g, in, out, pass, hd, fd))); }
    else { ROMIS_DISPATCH_N(N, (launch_pdl(spatial_halo_kernel<NT, false, false>, grid, block, s, sc, fr, g, in, out, pass, hd, fd))); }
}
}  // namespace romis
