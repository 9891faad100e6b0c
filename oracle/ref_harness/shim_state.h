/* shim_state.h -- state of the injected random stream (see rng_shim.h / shims.cpp). TEST ORACLE ONLY. */
#pragma once
#include <cstdint>
#include <vector>

enum ShimMode { SHIM_PARITY = 0, SHIM_TIMING = 1, SHIM_ASIS = 2 /* the reference's own rand() / random_device / mt19937 */ };
enum { SHIM_STAGE_NONE = -1,        // stages without draws: primary rays, final shading
       SHIM_STAGE_PRIMARY_THEN_NEIGH = -2 };   // renderRMIS: primary rays, then the neighbour index grid (no bar of its own)

struct ShimState {
    int mode = SHIM_PARITY;
    uint64_t seed = 0;
    uint32_t frame = 0;
    int W = 0, H = 0, N = 1, k = 0;
    std::vector<int> stage_queue;   // stage of each upcoming progressbar construction
    size_t stage_pos = 0;
    int stage = SHIM_STAGE_NONE;
    long pixel = -1;
    uint32_t engine_ctr = 0, rand_ctr = 0;
    long engine_total = 0, rand_total = 0;
    int rows_done = 0;
};
extern ShimState g_shim;
extern int g_tracer_mode;
extern "C" void romis_shim_stage_begin(int rows);
extern "C" void romis_shim_row_done(void);
