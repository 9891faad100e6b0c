/*
 * romis_gpu.h -- C-ABI of libromis_gpu.so: the B200 replacement for the body of the reference's
 * ReSTIR frame, `renderReSTIR` (reference src/rendering/render.cpp:28-62, declared render.h:25-28).
 *
 * The reference has no FFI layer: the hot path sits behind plain C++ free functions called from
 * src/main.cpp:164,220 and src/ui/ui.cpp:161.  A maintainer replaces the body of renderReSTIR by
 * the marshalling shown in INTEGRATION.md; every entry point below names the reference code it
 * stands in for.  Plain pointers and sizes only, no C++ or torch types, no exceptions: every call
 * returns ROMIS_OK (0) or a negative romis_status and leaves a message in romis_last_error().
 *
 * A context is bound to ONE CUDA device (or, created with several device ids, to one row band per device) and is not re-entrant (the reference's CLI mode calls
 * renderRayTraced from one thread per camera, main.cpp:213-230: create one context per thread).
 * There is no CPU fallback: romis_create fails when no CUDA device is usable.
 */
#ifndef ROMIS_GPU_H
#define ROMIS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ROMIS_ABI_VERSION 1

typedef enum romis_status {
    ROMIS_OK = 0,
    ROMIS_ERR_INVALID = -1,     /* bad argument / parameter outside the supported domain */
    ROMIS_ERR_CUDA = -2,        /* CUDA runtime error (message in romis_last_error) */
    ROMIS_ERR_STATE = -3,       /* call order (no scene, no frame in flight, ...) */
    ROMIS_ERR_NOMEM = -4
} romis_status;

typedef struct romis_ctx romis_ctx;

/* ---- scene: mirrors framework Vertex / Material / Mesh (reference framework/include/framework/mesh.h:14-43) ---- */
typedef struct romis_vertex {
    float position[3];
    float normal[3];
    float texcoord[2];
} romis_vertex;

typedef struct romis_material {
    float kd[3];
    float ks[3];
    float shininess;
    float transparency;
    int32_t kd_texture;         /* index into the textures passed to romis_upload_scene, -1 = none */
} romis_material;

typedef struct romis_mesh_desc {
    const romis_vertex* vertices;
    uint32_t n_vertices;
    const uint32_t* triangles;  /* 3 vertex indices per triangle (Mesh::triangles) */
    uint32_t n_triangles;
    romis_material material;    /* one material per mesh; geometryId == mesh index (embree_interface.cpp:46-47) */
} romis_mesh_desc;

typedef struct romis_texture {  /* framework Image (image.h): float RGB, row-major, width*height*3 */
    const float* pixels;
    int32_t width, height;
} romis_texture;

/* ---- lights: tagged POD mirror of PointLight / SegmentLight / ParallelogramLight (reference src/utils/common.h:72-87) ---- */
enum { ROMIS_LIGHT_POINT = 0, ROMIS_LIGHT_SEGMENT = 1, ROMIS_LIGHT_PARALLELOGRAM = 2 };
typedef struct romis_light {
    uint32_t type;
    float p0[3];    /* point: position      | segment: endpoint0 | parallelogram: v0     */
    float e1[3];    /*                      | segment: endpoint1 | parallelogram: edge01 */
    float e2[3];    /*                      |                    | parallelogram: edge02 */
    float c0[3];    /* point: color         | segment: color0    | parallelogram: color0 */
    float c1[3];    /*                      | segment: color1    | parallelogram: color1 */
    float c2[3];    /*                                           | parallelogram: color2 */
    float c3[3];    /*                                           | parallelogram: color3 */
} romis_light;

/* ---- parameters: POD mirror of the hot fields of Features (reference src/utils/common.h:89-136), same names ---- */
typedef struct romis_features {
    uint32_t enableShading;                 /* common.h:91  */
    uint32_t enableTextureMapping;          /* common.h:96  */
    uint32_t initialSamplesVisibilityCheck; /* common.h:104 "visibility reuse" */
    uint32_t numSamplesInReservoir;         /* common.h:105 N, 1..32 */
    uint32_t initialLightSamples;           /* common.h:106 M, >= 1 */
    uint32_t numNeighboursToSample;         /* common.h:107 k, 0..32 */
    uint32_t spatialResampleRadius;         /* common.h:108 r, pixels */
    uint32_t unbiasedCombination;           /* common.h:124 */
    uint32_t spatialReuse;                  /* common.h:125 */
    uint32_t spatialReuseVisibilityCheck;   /* common.h:126 */
    uint32_t temporalReuse;                 /* common.h:127 */
    uint32_t spatialResamplingPasses;       /* common.h:130 P */
    uint32_t temporalClampM;                /* common.h:131 */
    uint32_t enableToneMapping;             /* common.h:134 */
    float gamma;                            /* common.h:135 */
    float exposure;                         /* common.h:136 */
} romis_features;

/* ---- R-MIS / R-OMIS (renderRMIS, renderROMIS, reference src/rendering/render.cpp:64-265): the fields of Features only those modes read ---- */
enum { ROMIS_MIS_EQUAL = 0, ROMIS_MIS_BALANCE = 1 };                                /* MISWeightRMIS, common.h:31-34 */
enum { ROMIS_NEIGHBOURS_RANDOM = 0, ROMIS_NEIGHBOURS_SIMILAR = 1, ROMIS_NEIGHBOURS_DISSIMILAR = 2,
       ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR = 3 };                             /* NeighbourSelectionStrategy, common.h:36-41 */
typedef struct romis_rmis_params {
    uint32_t maxIterationsMIS;                          /* common.h:116 */
    uint32_t misWeightRMIS;                             /* common.h:118 */
    uint32_t neighbourSelectionStrategy;                /* common.h:117 */
    uint32_t neighbourSameGeometry;                     /* common.h:111 */
    float neighbourMaxDepthDifferenceFraction;          /* common.h:112 */
    float neighbourMaxNormalAngleDifferenceRadians;     /* common.h:113 (compared against the normals' dot product as is, neighbour_selection.cpp:18) */
    uint32_t useProgressiveROMIS;                       /* common.h:119 (R-OMIS only) */
    uint32_t progressiveUpdateMod;                      /* common.h:120 (R-OMIS only) */
} romis_rmis_params;

/* Camera: what Trackball::generateRay needs (reference framework/src/trackball.cpp:75-78,105-114).
 * origin = Trackball::position(); quat = glm::quat(rotationEulerAngles) as (w, x, y, z);
 * half_height = tan(fovy/2), half_width = aspect * half_height (trackball.cpp:26-27). */
typedef struct romis_camera {
    float origin[3];
    float quat[4];
    float half_width;
    float half_height;
} romis_camera;

/* Counter-based random stream (include/romis_rng.h): one seed per run, one frame index per frame. */
typedef struct romis_rng {
    uint64_t seed;
    uint32_t frame;
    uint32_t reserved;
} romis_rng;

/* ---- lifetime ---- */
/* device_ids / n_devices: n_devices == 1 (or 0 with a NULL list: device 0) binds the context to one GPU.
 * n_devices > 1 makes a MULTI-DEVICE context for the reference's actual caller -- one thread of one process (main.cpp:164,
 * ui.cpp:161): the frame is cut into one row band per device (equal cost from the per-row hit profile), the bands' boundary
 * reservoir rows travel between neighbouring devices over NVLink inside the spatial pass (cudaDeviceEnablePeerAccess), and
 * romis_render_frame fills the caller's single out_rgb; romis_render_frame_rmis / _romis cut their frames the same way (each band
 * renders its halo rows itself: nothing is exchanged) and leave the temporal history of a ReSTIR sequence alone, as the
 * reference's previousFrameGrid outlives them (render.cpp:268-280).  Results are bit-identical to the one-device frame.  On a
 * multi-device context the per-device calls (stepwise frames, bands, peer_*, render_frame_device, the R-MIS / R-OMIS parity
 * read-backs) return ROMIS_ERR_INVALID; the same device may be listed twice (two bands on one GPU: a test configuration).
 * Replaces: EmbreeInterface construction + the implicit process state of the reference (main.cpp:56-65). */
int romis_create(const int* device_ids, int n_devices, romis_ctx** out);
void romis_destroy(romis_ctx* ctx);
/* Last error text of `ctx`; ctx == NULL returns the last romis_create failure. */
const char* romis_last_error(const romis_ctx* ctx);
int romis_abi_version(void);

/* ---- scene / lights ---- */
/* Builds the BVH on the host and uploads geometry, materials and textures.  Replaces
 * EmbreeInterface::initScene / changeScene (embree_interface.cpp:30-56).  Resets temporal history. */
int romis_upload_scene(romis_ctx* ctx, const romis_mesh_desc* meshes, int n_meshes,
                       const romis_texture* textures, int n_textures);
/* Uploads scene.lights.  The reference reads them fresh every frame (light.cpp:46-66) and the UI edits them without
 * notification (ui.cpp:172-261), so call this every frame: every light is compared with what the device holds and only the
 * changed ones are sent (an unchanged table costs one pass over the caller's array and no transfer).
 * Edits and the temporal history: the reference's reservoirs hold LightSample{position, color} by value (reservoir.h:18-26), so
 * a history sample keeps the position / colour its light had when it was drawn, whatever happens to the light afterwards
 * (render_utils.cpp:154-170).  This call preserves exactly that: the old record of an edited or removed light is archived on
 * the device and the history samples drawn from it keep evaluating against the archived record.  Adding, removing, moving or
 * recolouring lights therefore never resets the history. */
int romis_upload_lights(romis_ctx* ctx, const romis_light* lights, int n_lights);
/* Same, for callers that know what they edited (the UI's light controls edit one selected light, ui.cpp:172-261): only
 * lights [first_dirty, first_dirty + n_dirty) are examined, everything else is taken as unchanged -- O(n_dirty) host work
 * instead of O(n_lights).  n_dirty == 0 is a no-op.  If n_lights differs from the current table the whole table is examined. */
int romis_upload_lights_range(romis_ctx* ctx, const romis_light* lights, int n_lights, int first_dirty, int n_dirty);
/* Archive bookkeeping (only needed with several contexts rendering bands of ONE frame: halo rows carry archive slots from band
 * to band, so every context must recycle the same slots at the same time).  With `auto` on (default) a context recycles the
 * slots its own history no longer holds at the next edit.  Hosts of banded contexts switch it off and, before every light
 * upload: fetch every context's marks (1 = slot still held by its history; n_slots = 0 when no edit is pending), OR them
 * across the contexts and hand the result to romis_light_archive_release on every context. */
int romis_set_light_archive_auto(romis_ctx* ctx, int on);
int romis_light_archive_marks(romis_ctx* ctx, uint8_t* marks, int capacity, int* n_slots);
int romis_light_archive_release(romis_ctx* ctx, const uint8_t* keep, int n_slots);
int romis_light_archive_size(romis_ctx* ctx, int* n_slots, int* n_held);

/* ---- the frame ---- */
/* One ReSTIR frame = renderReSTIR (render.cpp:28-62): primary hits -> initial RIS (+ visibility
 * reuse) -> temporal reuse -> spatial passes -> shade + tone map.  history_valid == 0 is the
 * reference's previousFrameGrid == nullptr (render.cpp:35); the returned ReservoirGrid becomes the
 * device-resident history of the next call.  out_rgb (nullable) receives width*height*3 floats in
 * Screen::pixels() layout, i.e. row (height-1-y) (screen.cpp:37-43); it may be pageable or pinned
 * (romis_host_alloc) host memory.  A change of width, height or numSamplesInReservoir drops the
 * history (the reference would read out of bounds). */
int romis_render_frame(romis_ctx* ctx, const romis_features* features, const romis_camera* camera,
                       int width, int height, int history_valid, const romis_rng* rng,
                       float* out_rgb);
int romis_reset_history(romis_ctx* ctx);
/* Same frame, result left on the device; *dev_rgb receives the device pointer of the RGB image
 * (valid until the next frame).  For callers that present from device memory. */
int romis_render_frame_device(romis_ctx* ctx, const romis_features* features, const romis_camera* camera,
                              int width, int height, int history_valid, const romis_rng* rng,
                              const float** dev_rgb);
int romis_synchronize(romis_ctx* ctx);
/* Largest c (with margin) such that pow(x, shininess) as this path evaluates it (include/romis_detmath.h) is +-0 or NaN for
 * every |x| <= c, or 0 when no such bound >= 0.05 exists: outside that lobe computeShading's specular term (reference
 * src/rendering/shading.cpp:21-28) is exactly zero and the kernels skip it.  Pure function, no context; exported so
 * that the tests can check the claim against the oracle's pow. */
float romis_specular_cutoff(float shininess);

/* One R-MIS frame = renderRMIS (render.cpp:64-119): primary hits, a neighbour index grid (k neighbours per pixel within the
 * spatial radius: random, or chosen by similarity of depth / normal / geometry, neighbour_selection.cpp), then
 * maxIterationsMIS rounds of { initial RIS per pixel; every pixel shades the samples of its k+1 neighbourhood pixels with
 * equal or balance-heuristic MIS weights and a shadow ray each }, averaged and tone mapped (combineToScreen,
 * render_utils.cpp:68-85).  No temporal state.  With a row band set (romis_set_band) only the band's rows are rendered and
 * written -- the band renders the radius rows around it itself, so bands need no exchange and, assembled, equal the whole frame
 * bit for bit; a multi-device context does exactly that with one band per device.
 * ROMIS_NEIGHBOURS_DISSIMILAR is rejected: the reference passes a negative count to std::sample there
 * (neighbour_selection.cpp:88-93), which is undefined behaviour. */
int romis_render_frame_rmis(romis_ctx* ctx, const romis_features* features, const romis_rmis_params* rmis,
                            const romis_camera* camera, int width, int height, const romis_rng* rng, float* out_rgb);
/* Parity read-back of the last R-MIS frame's neighbour grid: xy[H][W][k+1][2] (entry 0 is the pixel itself, unused
 * entries are -1) and count[H][W]. */
int romis_download_rmis_neighbours(romis_ctx* ctx, int32_t* xy, uint32_t* count);

/* One R-OMIS frame = renderROMIS (render.cpp:121-265): primary hits and neighbour index grid as R-MIS, then
 * maxIterationsMIS rounds of { initial RIS per pixel; every pixel adds, for each sample of its k+1 neighbourhood pixels, the
 * scaled column of all k+1 techniques' contribution-weight reciprocals (arbitraryUnbiasedContributionWeightReciprocal,
 * render_utils.cpp:245-257) to its (k+1)x(k+1) technique matrix and, times the shaded sample, to three contribution vectors },
 * then per pixel three minimum-norm least-squares solves (Eigen's completeOrthogonalDecomposition().solve, render_utils.h:52;
 * here include/romis_cod.h) whose components are summed, tone mapped and written in Screen layout (direct estimator).  With
 * useProgressiveROMIS the solves run before every progressiveUpdateMod-th iteration instead and a running estimate built from
 * the current alphas is averaged over the iterations (render.cpp:160-200,232).  numNeighboursToSample <= 10; every pixel's
 * window must hold k other pixels (the reference reads out of bounds otherwise, render.cpp:165).  Row bands and multi-device
 * contexts as for R-MIS. */
int romis_render_frame_romis(romis_ctx* ctx, const romis_features* features, const romis_rmis_params* rmis,
                             const romis_camera* camera, int width, int height, const romis_rng* rng, float* out_rgb);
/* Parity read-back of the last R-OMIS frame: matrices[H][W][k+1][k+1] (row-major) and contributions[H][W][3][k+1]
 * (red, green, blue) as accumulated over all iterations. */
int romis_download_romis_system(romis_ctx* ctx, float* matrices, float* contributions);

/* ---- row-band sharding (one context per GPU) ---- */
/* This context renders rows [y0, y1) of the height passed to the frame calls.  Pixels, RNG keys and
 * the output layout stay in global image coordinates, so N bands reproduce the 1-GPU frame bit for
 * bit.  Default (or y0 = y1 = 0): whole frame. */
int romis_set_band(romis_ctx* ctx, int y0, int y1);
/* Pixels per image row whose primary ray hits geometry (hits_per_row: height entries).  Work per pixel is concentrated in
 * hit pixels, so hosts cut the frame into equal-COST bands with this profile; every rank computes the same numbers. */
int romis_row_hit_counts(romis_ctx* ctx, const romis_camera* camera, int width, int height, uint32_t* hits_per_row);
/* Stepwise frame for banded rendering: begin = primary (band + radius halo rows) + initial +
 * temporal; then per spatial pass: the caller moves halo rows between neighbouring bands
 * (romis_halo_region, any transport), calls romis_frame_spatial_pass; end = shade + read-back of the
 * band's rows into out_rgb (full-frame layout; only the band's rows are written). */
int romis_frame_begin(romis_ctx* ctx, const romis_features* features, const romis_camera* camera,
                      int width, int height, int history_valid, const romis_rng* rng);
int romis_frame_spatial_pass(romis_ctx* ctx, int pass);
int romis_frame_end(romis_ctx* ctx, float* out_rgb);
/* Halo rows (r = spatialResampleRadius) of the reservoir buffer the NEXT spatial pass reads, as contiguous device ranges:
 *   ROMIS_HALO_SEND_LOW   my rows [y0, y0 + r)      -> to the band below (smaller y), its ROMIS_HALO_RECV_HIGH
 *   ROMIS_HALO_SEND_HIGH  my rows [y1 - r, y1)      -> to the band above, its ROMIS_HALO_RECV_LOW
 *   ROMIS_HALO_RECV_LOW   rows [max(0, y0 - r), y0)    <- from the band below
 *   ROMIS_HALO_RECV_HIGH  rows [y1, min(H, y1 + r))    <- from the band above
 * A range is empty (bytes = 0) at the image border.  Valid between romis_frame_begin and romis_frame_end; re-query
 * before every pass (the buffers ping-pong). */
enum { ROMIS_HALO_SEND_LOW = 0, ROMIS_HALO_SEND_HIGH = 1, ROMIS_HALO_RECV_LOW = 2, ROMIS_HALO_RECV_HIGH = 3 };
int romis_halo_region(romis_ctx* ctx, int which, void** dev_ptr, size_t* bytes);
/* CUDA stream (cudaStream_t) the context launches on, so the caller can order its transport. */
int romis_stream(romis_ctx* ctx, void** cuda_stream);

/* Peer-mapped halos (one process per GPU, all GPUs of one NVLink/NVSwitch node).  Instead of the caller moving halo rows,
 * romis_frame_spatial_pass pushes this band's boundary rows straight into the neighbouring bands' halo rows (CUDA IPC
 * mapped device memory) and orders the passes with flag words in device memory: no NCCL call, no host synchronisation.
 *   1. every rank: romis_set_band, romis_upload_scene, romis_band_prepare (allocates the frame buffers now)
 *   2. every rank: romis_peer_export -> ROMIS_PEER_BLOB_BYTES opaque bytes; exchange them out of band (any transport)
 *   3. every rank: romis_peer_attach(blob of the band below or NULL, blob of the band above or NULL)
 * All ranks must then issue the same sequence of frames.  A change of resolution / N / band needs a detach and a new
 * prepare-export-attach round.  romis_peer_error reports whether a flag wait ever timed out (a neighbour died). */
#define ROMIS_PEER_BLOB_BYTES 512
int romis_band_prepare(romis_ctx* ctx, const romis_features* features, int width, int height);
int romis_peer_export(romis_ctx* ctx, void* blob);
int romis_peer_attach(romis_ctx* ctx, const void* low_blob, const void* high_blob);
int romis_peer_detach(romis_ctx* ctx);
int romis_peer_error(romis_ctx* ctx, int* timed_out);

/* ---- parity / debug read-back (SURVEY.md 8b romis_download_reservoirs) ---- */
enum { ROMIS_PASS_INITIAL = 0, ROMIS_PASS_TEMPORAL = 1, ROMIS_PASS_SPATIAL0 = 2 /* + pass */, ROMIS_PASS_FINAL = 1000 };
typedef struct romis_reservoir_dump {   /* every pointer nullable; arrays are [N][height][width] (band rows only are written) */
    uint32_t* light_id;
    float* u;
    float* v;
    float* W;           /* outputWeight */
    uint32_t* M;        /* sampleNums   */
    float* position;    /* [N][H][W][3] LightSample::position recomputed from (light, u, v) */
    float* color;       /* [N][H][W][3] */
} romis_reservoir_dump;
typedef struct romis_gbuffer_dump {     /* [height][width] */
    float* t;           /* Ray::t, FLT_MAX on miss */
    float* normal;      /* [H][W][3] interpolated, un-normalised */
    float* texcoord;    /* [H][W][2] */
    uint32_t* mesh;     /* geometryId; n_meshes on miss */
} romis_gbuffer_dump;
/* Capture mode keeps a copy of the reservoir buffer after every stage so that
 * romis_download_reservoirs can return per-stage state (costs bandwidth; off by default, in which
 * case only ROMIS_PASS_FINAL is available). */
int romis_set_capture(romis_ctx* ctx, int enable);
int romis_download_reservoirs(romis_ctx* ctx, int pass_id, romis_reservoir_dump* out);
int romis_download_gbuffer(romis_ctx* ctx, romis_gbuffer_dump* out);
/* Raw ray queries against the uploaded scene: closestHit / anyHit (embree_interface.cpp:58-90).
 * origins/dirs: n*3 floats, tfar: n floats.  any_hit != 0: hit[i] = occluded.  Otherwise t,u,v,tri
 * (global triangle index) of the closest hit; hit[i] = 0 leaves the others untouched. */
int romis_trace_rays(romis_ctx* ctx, const float* origins, const float* dirs, const float* tfar, int n,
                     int any_hit, uint8_t* hit, float* t, float* u, float* v, uint32_t* tri);

/* Self-test of the kernels' three-divisions-by-one-denominator routine (csrc/device_common.cuh div3_shared) against the plain
 * IEEE division, both on the device: num n*3 floats, den n floats; out_fast / out_ref n*3 floats each, to be compared bit for
 * bit by the caller.  No reference counterpart: the reference divides with glm's operator/ (shading.cpp:33). */
int romis_selftest_division(romis_ctx* ctx, const float* num, const float* den, int n, float* out_fast, float* out_ref);

/* ---- measurement ---- */
typedef struct romis_timings {          /* device time of the last frame, milliseconds (CUDA events) */
    float primary_ms;
    float initial_ms;
    float temporal_ms;
    float spatial_ms[8];
    float shade_ms;
    float total_ms;                     /* first kernel start .. last kernel end (no read-back) */
    int32_t n_spatial;
    int32_t n_launches;                 /* kernels launched for the frame */
    float exchange_ms[8];               /* peer-mapped halo push + wait before each spatial pass (not part of spatial_ms) */
    /* R-MIS / R-OMIS frames (initial_ms = sum over the iterations): */
    float neighbours_ms;                /* neighbour index grid */
    float gather_ms;                    /* R-MIS gather / R-OMIS accumulation, sum over the iterations */
    float resolve_ms;                   /* R-MIS combineToScreen / R-OMIS per-pixel solves */
} romis_timings;
/* Per-stage events are recorded only when enabled; total_ms is always available. */
int romis_set_stage_timing(romis_ctx* ctx, int enable);
int romis_last_frame_timings(romis_ctx* ctx, romis_timings* out);

/* ---- pinned host memory for out_rgb (Screen::pixels() storage) ---- */
void* romis_host_alloc(size_t bytes);
void romis_host_free(void* p);
/* Page-lock / release memory the caller owns (e.g. the std::vector behind Screen::pixels(), screen.h): out_rgb copies into
 * registered memory run as asynchronous DMA overlapped with shading, pageable memory goes through the driver's staging. */
int romis_host_register(void* p, size_t bytes);
int romis_host_unregister(void* p);

#ifdef __cplusplus
}
#endif
#endif /* ROMIS_GPU_H */
