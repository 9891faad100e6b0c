// launch.hpp -- host-callable launchers of the pass kernels (one translation unit per pass so they build in parallel).
#pragma once
#include "device_common.cuh"

#include <cstdlib>
#include <utility>

namespace romis {
// Launch with programmatic stream serialisation (see pdl_wait in device_common.cuh); ROMIS_PDL=0 in the environment turns it off.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args) {
    static const bool on = [] { const char* e = std::getenv("ROMIS_PDL"); return !e || std::atoi(e) != 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = on ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
void launch_primary(cudaStream_t s, dim3 grid, dim3 block, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, int row0, int row1);
void launch_initial(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& out,
                    float* wsum = nullptr, float* chosen = nullptr);
void launch_temporal(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g,
                     const ResBuf& cur, const ResBuf& prev, const ResBuf& out, const FineDev& fd, const HaloDev& hd);
void launch_spatial(cudaStream_t s, dim3 grid, dim3 block, int N, bool unbiased, const SceneDev& sc, const FrameDev& fr, const GBufDev& g,
                    const ResBuf& in, const ResBuf& out, int pass, const FineDev& fd);
void launch_spatial_halo(cudaStream_t s, dim3 grid, dim3 block, int N, bool unbiased, const SceneDev& sc, const FrameDev& fr, const GBufDev& g,
                         const ResBuf& in, const ResBuf& out, int pass, const HaloDev& hd, const FineDev& fd);
void launch_shade(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& in, float* rgb,
                  const FineDev& fd);
void launch_ctx(cudaStream_t s, dim3 grid, dim3 block, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, int row0, int row1);
void launch_rmis_neighbours(cudaStream_t s, dim3 grid, dim3 block, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const RmisDev& rm);
void launch_rmis_gather(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& in, const RmisDev& rm);
void launch_romis_accumulate(cudaStream_t s, dim3 grid, dim3 block, int N, const SceneDev& sc, const FrameDev& fr, const GBufDev& g, const ResBuf& in, const RmisDev& rm);
void launch_romis_solve(cudaStream_t s, dim3 grid, dim3 block, const FrameDev& fr, const RmisDev& rm, float* rgb, bool alphas_only);
void launch_rmis_combine(cudaStream_t s, dim3 grid, dim3 block, const FrameDev& fr, const RmisDev& rm, float* rgb);
void launch_trace(cudaStream_t s, int n, const SceneDev& sc, const float* o, const float* d, const float* tfar, int any_hit,
                  uint8_t* hit, float* t, float* u, float* v, uint32_t* tri);
void launch_division_selftest(cudaStream_t s, const float* num, const float* den, int n, float* out_fast, float* out_ref);
void launch_row_hits(cudaStream_t s, const GBufDev& g, int W, int H, int n_meshes, uint32_t* rows);
void launch_halo_push(cudaStream_t s, const void* src_low, void* dst_low, size_t bytes_low, const void* src_high, void* dst_high, size_t bytes_high,
                      const uint32_t* war_a, const uint32_t* war_b, uint32_t war_token, uint32_t* sig_a, uint32_t* sig_b, uint32_t token,
                      const uint32_t* wait_a, const uint32_t* wait_b, uint32_t* err, unsigned int* ticket);
void launch_signal(cudaStream_t s, uint32_t* a, uint32_t va, uint32_t* b, uint32_t vb);
void launch_wait(cudaStream_t s, const uint32_t* a, uint32_t va, const uint32_t* b, uint32_t vb, uint32_t* err);
void launch_dump(cudaStream_t s, dim3 grid, dim3 block, const SceneDev& sc, const FrameDev& fr, const ResBuf& in, int N, const uint32_t* arch_orig,
                 uint32_t* light, float* u, float* v, float* W, uint32_t* M, float* pos, float* col);
void launch_light_archive(cudaStream_t s, const float4* lights, float4* arch, uint32_t* remap, const uint32_t* idx, const uint32_t* slot, int n,
                          const ResBuf& hist, int lrow0, int rows, int N, uint32_t n_remap, uint8_t* mark, uint32_t n_slots);
}  // namespace romis

// numSamplesInReservoir 1..4 get register-resident instantiations; anything else the generic one
#define ROMIS_DISPATCH_N(N, CALL)               \
    switch (N) {                                \
        case 1: { constexpr int NT = 1; CALL; } break;  \
        case 2: { constexpr int NT = 2; CALL; } break;  \
        case 3: { constexpr int NT = 3; CALL; } break;  \
        case 4: { constexpr int NT = 4; CALL; } break;  \
        default: { constexpr int NT = 0; CALL; } break; \
    }
