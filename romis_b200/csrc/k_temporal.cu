// k_temporal.cu -- temporal reuse pass (temporalReuse, reference src/rendering/render_utils.cpp:142-177).
#include "reservoir.cuh"
#include "launch.hpp"

namespace romis {

// ------------------------------------------------------------------------------------------------
// temporal reuse: same-pixel predecessor, M clamp, biased combine of {current, predecessor}
// ------------------------------------------------------------------------------------------------
template <int NT, bool ES>        // ES: enableShading known to be on, see spatial_kernel
__global__ void __launch_bounds__(ROMIS_LBT_TEMPORAL, ROMIS_MINB_TEMPORAL) temporal_kernel(SceneDev sc, FrameDev fr, GBufDev g, ResBuf cur, ResBuf prev, ResBuf out, FineDev fd, HaloDev hd) {
    int x, y; thread_pixel<false>(x, y);
    y += fr.y0;
    if (x >= fr.W || y >= fr.y1) return;
    const int by0 = fr.y0 + (int)(blockIdx.y * blockDim.y), by1 = by0 + (int)blockDim.y;
    pdl_wait();                                     // `cur` comes from the kernel before this one
    pdl_launch_dependents();
    const int N = NT > 0 ? NT : (int)fr.f.numSamplesInReservoir;
    const bool es = ES || fr.f.enableShading != 0;
    const int lrow = y - fr.ey0;
    const uint32_t pixel = (uint32_t)y * (uint32_t)fr.W + (uint32_t)x;
    constexpr int CAP = SubRes<NT>::CAP;
    uint4 crec[CAP], prec[CAP]; uint32_t cM[CAP], pM[CAP]; float cpdf[CAP];
    uint64_t curTotal = 0, prevTotal = 0;
    ROMIS_FOR_SUB(j, NT, N) {
        crec[j] = res_rec(cur, lrow, j)[x]; cM[j] = res_m(cur, lrow, j)[x]; cpdf[j] = res_pdf(cur, lrow, j)[x];
        prec[j] = res_rec(prev, lrow, j)[x]; pM[j] = res_m(prev, lrow, j)[x];
        curTotal += cM[j]; prevTotal += pM[j];
    }
    // render_utils.cpp:156-163: cap = clampM * totalM(cur) + 1; every non-empty predecessor sub-reservoir gets M = cap
    uint64_t cap = (uint64_t)fr.f.temporalClampM * curTotal + 1ull;
    if (prevTotal > cap) {
        uint32_t cap32 = cap > 0xffffffffull ? 0xffffffffu : (uint32_t)cap;
        ROMIS_FOR_SUB(j, NT, N) { if (pM[j] != 0u) pM[j] = cap32; }
    }
    PixCtx c = make_ctx(sc, fr, g, x, y);
    romis_stream_key rk = romis_rng_stream(fr.seed, fr.frame, ROMIS_STAGE_TEMPORAL, pixel, ROMIS_STREAM_RAND);
    uint32_t rc = 0;
    SubRes<NT> r; res_init(r, N);
    ROMIS_FOR_SUB(j, NT, N) stream_sample(r, N, sc, c, es, crec[j], cM[j], rk, rc, cpdf[j]);    // :169 current first (pdf at this pixel known)
        ROMIS_FOR_SUB(j, NT, N) stream_sample(r, N, sc, c, es, prec[j], pM[j], rk, rc);    // then the predecessor
    res_take_counts(r, N);
    res_finish(r, N, sc, c, es);
    res_store(out, lrow, x, r, N);
    fine_signal(fd, by0, by1);
    // Stage 0 of the frame's halo exchange (romis_gpu.cu "fused halo exchange"): the first spatial pass of the neighbouring bands
    // reads this pass's boundary rows, so the blocks that hold them store them a second time into the neighbours' halo rows --
    // once the neighbour's last token of the PREVIOUS frame says it has finished with those rows -- and the last such block of
    // an edge publishes the stage token.  No copy kernel between the temporal and the first spatial pass.
    const bool edge0 = hd.peer_out[0] != nullptr && by0 < fr.y0 + hd.r;
    const bool edge1 = hd.peer_out[1] != nullptr && by1 > fr.y1 - hd.r;
    if (edge0 || edge1) {
        if (edge0) halo_spin(hd.wait_flag[0], hd.wait_token, hd.err);
        if (edge1) halo_spin(hd.wait_flag[1], hd.wait_token, hd.err);
        _Pragma("unroll") for (int e = 0; e < 2; e++) {
            if (!(e == 0 ? edge0 && y < fr.y0 + hd.r : edge1 && y >= fr.y1 - hd.r)) continue;
            ResBuf pb; pb.base = hd.peer_out[e]; pb.row_stride = hd.peer_stride[e]; pb.W = out.W; pb.N = out.N;
            const int prow = y - hd.peer_ey0[e];
            ROMIS_FOR_SUB(j, NT, N) {
                res_rec(pb, prow, j)[x] = make_uint4(r.light[j], __float_as_uint(r.u[j]), __float_as_uint(r.v[j]), __float_as_uint(r.W[j]));
                res_m(pb, prow, j)[x] = r.M[j];
            }
        }
        __syncthreads();                            // release pattern as in spatial_halo_kernel
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


void launch_temporal(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g,
                     const ResBuf& cur, const ResBuf& prev, const ResBuf& out, const FineDev& fd, const HaloDev& hd) {
    if (fr.f.enableShading) { ROMIS_DISPATCH_N(N, (launch_pdl(temporal_kernel<NT, true>, grid, block, s, sc, fr, g, cur, prev, out, fd, hd))); }
    else { ROMIS_DISPATCH_N(N, (launch_pdl(temporal_kernel<NT, false>, grid, block, s, sc, fr, g, cur, prev, out, fd, hd))); }
}
}  // namespace romis
