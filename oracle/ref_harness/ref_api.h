/* ref_api.h -- C entry points of oracle/_ref/libromis_ref.so (the compiled REFERENCE). TEST ORACLE ONLY. */
#pragma once
#include <stdint.h>
#include "romis_gpu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ref_camera_desc {    /* CameraConfig of the reference (src/utils/config.h:21-26) */
    float fov_deg;
    float distance;
    float look_at[3];
    float rotation_deg[3];
} ref_camera_desc;

typedef struct ref_reservoir_dump { /* [N][H][W] each; every pointer nullable */
    float* position;                /* [N][H][W][3] */
    float* color;                   /* [N][H][W][3] */
    float* W;
    uint64_t* M;
    float* wSum;
    float* chosenW;
} ref_reservoir_dump;

typedef struct ref_frame_dump {
    float* gbuffer_t; float* gbuffer_normal; float* gbuffer_texcoord; uint32_t* gbuffer_mesh;
    float* ray_dir; float* ray_origin;
    ref_reservoir_dump* initial;
    ref_reservoir_dump* temporal;
    ref_reservoir_dump* spatial[8];  /* per pass, only with REF_FLAG_SPLIT_SPATIAL */
    ref_reservoir_dump* final_;     /* grid returned by renderReSTIR (after the last spatial pass) */
} ref_frame_dump;

typedef struct ref_timings { double primary_ms, initial_ms, temporal_ms, spatial_ms, shade_ms, total_ms, grid_copy_ms; } ref_timings;

enum { REF_FLAG_WHOLE_FRAME = 1, REF_FLAG_TIMING_RNG = 2, REF_FLAG_SPLIT_SPATIAL = 4,
       REF_FLAG_ASIS_RNG = 8 /* timing with the reference's own random sources: glibc rand(), std::random_device + std::mt19937 per pixel */ };

const char* ref_last_error(void);
int ref_set_tracer_mode(int mode);
int ref_load_prebuilt(int scene_type, const char* data_dir);
int ref_set_scene(const romis_mesh_desc* meshes, int n_meshes, const romis_texture* textures, int n_textures);
int ref_set_lights(const romis_light* lights, int n);
int ref_scene_info(int* n_meshes, int* n_textures, int* n_lights);
int ref_mesh_info(int i, uint32_t* n_vertices, uint32_t* n_triangles, romis_material* mat);
int ref_mesh_data(int i, romis_vertex* v, uint32_t* tris);
int ref_texture_info(int i, int* w, int* h);
int ref_texture_data(int i, float* px);
int ref_lights_data(romis_light* out);
int ref_make_camera(const ref_camera_desc* c, int width, int height, romis_camera* out);
int ref_reset_history(void);
int ref_render_frame_rmis(const romis_features* f, const romis_rmis_params* rp, const ref_camera_desc* cam, int W, int H,
                          const romis_rng* rng, float* out_rgb, int32_t* neigh_xy, uint32_t* neigh_count);
/* renderROMIS (reference src/rendering/render.cpp:121-265), called whole.  matrices: [H][W][K1][K1] row-major (K1 = k + 1),
 * contributions: [H][W][3][K1] (red, green, blue), both as they stand after the LAST iteration (captured through the
 * visualiseAlphas call of render.cpp:228); any pointer may be NULL. */
int ref_render_frame_romis(const romis_features* f, const romis_rmis_params* rp, const ref_camera_desc* cam, int W, int H,
                           const romis_rng* rng, float* out_rgb, float* matrices, float* contributions);
int ref_set_mis_timing(int on);
int ref_num_threads(void);
/* OpenMP threads of the timing runs (n <= 0: all processors); overrides OMP_NUM_THREADS, which torchrun sets to 1 */
int ref_set_num_threads(int n);
int ref_render_frame(const romis_features* f, const ref_camera_desc* cam, int W, int H, int history_valid,
                     const romis_rng* rng, int flags, ref_frame_dump* dump, float* out_rgb, ref_timings* tm);
#ifdef __cplusplus
}
#endif
