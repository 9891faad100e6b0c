/*
 * ref_api.cpp -- C entry points of oracle/_ref/libromis_ref.so: the REFERENCE's own hot path
 * (unmodified translation units from /root/reference, see oracle/Makefile) behind the same POD
 * types as include/romis_gpu.h, so tests and bench.py can run it next to the CUDA path.
 *
 * TEST ORACLE ONLY (tests/, __graft_entry__.smoke, bench.py's cpu_baseline / --impl reference).
 *
 * ref_render_frame does what renderReSTIR does (reference src/rendering/render.cpp:28-62), by
 * calling the reference's stage functions genPrimaryRayHits, genInitialSamples, temporalReuse,
 * spatialReuse, finalShading + exposureToneMapping + Screen::setPixel one by one so that the state
 * after every stage can be dumped (mode 0), or by calling renderReSTIR itself (mode 1).
 */
#include <framework/trackball.h>
#include <framework/window.h>
#include <post_processing/tone_mapping.h>
#include <rendering/render.h>
#include <rendering/neighbour_selection.h>
#include <rendering/render_utils.h>
#include <rendering/reservoir.h>
#include <rendering/screen.h>
#include <scene/light.h>
#include <scene/scene.h>
#include <utils/common.h>
#include <utils/utils.h>

#include <chrono>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "romis_gpu.h"
#include "romis_rng.h"
#include "shim_state.h"
#include "ref_api.h"

namespace {
Scene g_scene;
std::unique_ptr<EmbreeInterface> g_embree;
std::shared_ptr<ReservoirGrid> g_prev;
std::string g_err;
std::vector<std::shared_ptr<Image>> g_textures;

Features toFeatures(const romis_features& f) {
    Features o;                                   // defaults of common.h:89-136
    o.rayTraceMode                  = RayTraceMode::ReSTIR;
    o.enableShading                 = f.enableShading != 0;
    o.enableTextureMapping          = f.enableTextureMapping != 0;
    o.initialSamplesVisibilityCheck = f.initialSamplesVisibilityCheck != 0;
    o.numSamplesInReservoir         = f.numSamplesInReservoir;
    o.initialLightSamples           = f.initialLightSamples;
    o.numNeighboursToSample         = f.numNeighboursToSample;
    o.spatialResampleRadius         = f.spatialResampleRadius;
    o.unbiasedCombination           = f.unbiasedCombination != 0;
    o.spatialReuse                  = f.spatialReuse != 0;
    o.spatialReuseVisibilityCheck   = f.spatialReuseVisibilityCheck != 0;
    o.temporalReuse                 = f.temporalReuse != 0;
    o.spatialResamplingPasses       = f.spatialResamplingPasses;
    o.temporalClampM                = f.temporalClampM;
    o.enableToneMapping             = f.enableToneMapping != 0;
    o.gamma                         = f.gamma;
    o.exposure                      = f.exposure;
    return o;
}

void v3(float* d, const glm::vec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
glm::vec3 g3(const float* s) { return glm::vec3(s[0], s[1], s[2]); }

void dumpGrid(const ReservoirGrid& grid, int W, int H, int N, ref_reservoir_dump* d) {
    if (!d) return;
    for (int j = 0; j < N; j++) for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const Reservoir& r = grid[y][x];
        size_t i = (size_t(j) * H + y) * W + x;
        if (d->position) v3(d->position + 3 * i, r.outputSamples[j].lightSample.position);
        if (d->color)    v3(d->color + 3 * i, r.outputSamples[j].lightSample.color);
        if (d->W)        d->W[i] = r.outputSamples[j].outputWeight;
        if (d->M)        d->M[i] = (uint64_t)r.sampleNums[j];
        if (d->wSum)     d->wSum[i] = r.wSums[j];
        if (d->chosenW)  d->chosenW[i] = r.chosenSampleWeights[j];
    }
}

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
}

// ---- capture hook: renderROMIS calls visualiseAlphas after every iteration when saveAlphasVisualisation is set
// (render.cpp:227-229); the reference's own definition is compiled under another name (oracle/Makefile) ----
namespace { float* g_cap_matrices = nullptr; float* g_cap_contrib = nullptr; int g_mis_timing = 0; }
void visualiseAlphas(const MatrixGrid& techniqueMatrices, const VectorGrid& contributionVectorsRed,
                     const VectorGrid& contributionVectorsGreen, const VectorGrid& contributionVectorsBlue,
                     const glm::ivec2& windowResolution, const Features& features) {
    const int W = windowResolution.x, H = windowResolution.y, K1 = (int)features.numNeighboursToSample + 1;
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const size_t p = size_t(y) * W + x;
        if (g_cap_matrices) for (int i = 0; i < K1; i++) for (int j = 0; j < K1; j++)
            g_cap_matrices[(p * K1 + i) * K1 + j] = techniqueMatrices[y][x](i, j);
        if (g_cap_contrib) for (int i = 0; i < K1; i++) {
            g_cap_contrib[(p * 3 + 0) * K1 + i] = contributionVectorsRed[y][x](i);
            g_cap_contrib[(p * 3 + 1) * K1 + i] = contributionVectorsGreen[y][x](i);
            g_cap_contrib[(p * 3 + 2) * K1 + i] = contributionVectorsBlue[y][x](i);
        }
    }
}

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

int ref_set_tracer_mode(int mode) { g_tracer_mode = mode; if (g_embree) g_embree->changeScene(g_scene); return 0; }

int ref_load_prebuilt(int scene_type, const char* data_dir) {
    try {
        g_scene = loadScenePrebuilt(static_cast<SceneType>(scene_type), std::filesystem::path(data_dir));
        g_embree = std::make_unique<EmbreeInterface>(g_scene);
        g_prev.reset();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

int ref_set_scene(const romis_mesh_desc* meshes, int n_meshes, const romis_texture* textures, int n_textures) {
    try {
        Scene s; s.type = SceneType::SingleTriangle; s.lights = g_scene.lights;
        g_textures.clear();
        // Image has only a file-loading constructor (framework/include/framework/image.h); build it raw.
        for (int t = 0; t < n_textures; t++) {
            Image* img = static_cast<Image*>(::operator new(sizeof(Image)));
            new (&img->pixels) std::vector<glm::vec3>();
            img->width = textures[t].width; img->height = textures[t].height;
            size_t n = size_t(img->width) * img->height;
            img->pixels.resize(n);
            for (size_t i = 0; i < n; i++) img->pixels[i] = g3(textures[t].pixels + 3 * i);
            g_textures.emplace_back(img, [](Image* p) { p->pixels.~vector(); ::operator delete(p); });
        }
        for (int m = 0; m < n_meshes; m++) {
            Mesh mesh;
            mesh.vertices.resize(meshes[m].n_vertices);
            for (uint32_t i = 0; i < meshes[m].n_vertices; i++) {
                const romis_vertex& v = meshes[m].vertices[i];
                mesh.vertices[i].position = g3(v.position);
                mesh.vertices[i].normal = g3(v.normal);
                mesh.vertices[i].texCoord = glm::vec2(v.texcoord[0], v.texcoord[1]);
            }
            mesh.triangles.resize(meshes[m].n_triangles);
            for (uint32_t i = 0; i < meshes[m].n_triangles; i++)
                mesh.triangles[i] = glm::uvec3(meshes[m].triangles[3 * i], meshes[m].triangles[3 * i + 1], meshes[m].triangles[3 * i + 2]);
            const romis_material& mt = meshes[m].material;
            mesh.material.kd = g3(mt.kd); mesh.material.ks = g3(mt.ks);
            mesh.material.shininess = mt.shininess; mesh.material.transparency = mt.transparency;
            if (mt.kd_texture >= 0 && mt.kd_texture < n_textures) mesh.material.kdTexture = g_textures[mt.kd_texture];
            s.meshes.push_back(std::move(mesh));
        }
        g_scene = std::move(s);
        g_embree = std::make_unique<EmbreeInterface>(g_scene);
        g_prev.reset();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

int ref_set_lights(const romis_light* lights, int n) {
    g_scene.lights.clear();
    for (int i = 0; i < n; i++) {
        const romis_light& l = lights[i];
        switch (l.type) {
        case ROMIS_LIGHT_POINT:   g_scene.lights.emplace_back(PointLight { g3(l.p0), g3(l.c0) }); break;
        case ROMIS_LIGHT_SEGMENT: g_scene.lights.emplace_back(SegmentLight { g3(l.p0), g3(l.e1), g3(l.c0), g3(l.c1) }); break;
        case ROMIS_LIGHT_PARALLELOGRAM:
            g_scene.lights.emplace_back(ParallelogramLight { g3(l.p0), g3(l.e1), g3(l.e2), g3(l.c0), g3(l.c1), g3(l.c2), g3(l.c3) }); break;
        default: g_err = "bad light type"; return -1;
        }
    }
    return 0;
}

// ---- scene export (fixture generation: tests/golden/gen_golden.py) ----
int ref_scene_info(int* n_meshes, int* n_textures, int* n_lights) {
    *n_meshes = (int)g_scene.meshes.size();
    std::vector<const Image*> tex;
    for (const Mesh& m : g_scene.meshes) if (m.material.kdTexture) {
        bool seen = false; for (const Image* t : tex) seen |= (t == m.material.kdTexture.get());
        if (!seen) tex.push_back(m.material.kdTexture.get());
    }
    *n_textures = (int)tex.size();
    *n_lights = (int)g_scene.lights.size();
    return 0;
}
static std::vector<const Image*> sceneTextures() {
    std::vector<const Image*> tex;
    for (const Mesh& m : g_scene.meshes) if (m.material.kdTexture) {
        bool seen = false; for (const Image* t : tex) seen |= (t == m.material.kdTexture.get());
        if (!seen) tex.push_back(m.material.kdTexture.get());
    }
    return tex;
}
int ref_mesh_info(int i, uint32_t* n_vertices, uint32_t* n_triangles, romis_material* mat) {
    const Mesh& m = g_scene.meshes.at(i);
    *n_vertices = (uint32_t)m.vertices.size(); *n_triangles = (uint32_t)m.triangles.size();
    v3(mat->kd, m.material.kd); v3(mat->ks, m.material.ks);
    mat->shininess = m.material.shininess; mat->transparency = m.material.transparency;
    mat->kd_texture = -1;
    auto tex = sceneTextures();
    for (size_t t = 0; t < tex.size(); t++) if (tex[t] == m.material.kdTexture.get()) mat->kd_texture = (int)t;
    return 0;
}
int ref_mesh_data(int i, romis_vertex* v, uint32_t* tris) {
    const Mesh& m = g_scene.meshes.at(i);
    for (size_t k = 0; k < m.vertices.size(); k++) {
        v3(v[k].position, m.vertices[k].position); v3(v[k].normal, m.vertices[k].normal);
        v[k].texcoord[0] = m.vertices[k].texCoord.x; v[k].texcoord[1] = m.vertices[k].texCoord.y;
    }
    for (size_t k = 0; k < m.triangles.size(); k++) for (int c = 0; c < 3; c++) tris[3 * k + c] = m.triangles[k][c];
    return 0;
}
int ref_texture_info(int i, int* w, int* h) { auto t = sceneTextures(); *w = t.at(i)->width; *h = t.at(i)->height; return 0; }
int ref_texture_data(int i, float* px) {
    auto t = sceneTextures(); const Image* im = t.at(i);
    for (size_t k = 0; k < im->pixels.size(); k++) v3(px + 3 * k, im->pixels[k]);
    return 0;
}
int ref_lights_data(romis_light* out) {
    for (size_t i = 0; i < g_scene.lights.size(); i++) {
        romis_light l; std::memset(&l, 0, sizeof l);
        const auto& v = g_scene.lights[i];
        if (std::holds_alternative<PointLight>(v)) {
            const auto& p = std::get<PointLight>(v); l.type = ROMIS_LIGHT_POINT; v3(l.p0, p.position); v3(l.c0, p.color);
        } else if (std::holds_alternative<SegmentLight>(v)) {
            const auto& s = std::get<SegmentLight>(v); l.type = ROMIS_LIGHT_SEGMENT;
            v3(l.p0, s.endpoint0); v3(l.e1, s.endpoint1); v3(l.c0, s.color0); v3(l.c1, s.color1);
        } else {
            const auto& p = std::get<ParallelogramLight>(v); l.type = ROMIS_LIGHT_PARALLELOGRAM;
            v3(l.p0, p.v0); v3(l.e1, p.edge01); v3(l.e2, p.edge02);
            v3(l.c0, p.color0); v3(l.c1, p.color1); v3(l.c2, p.color2); v3(l.c3, p.color3);
        }
        out[i] = l;
    }
    return 0;
}

// Camera exactly as main.cpp:58-59 builds it; exports what the C-ABI's romis_camera carries.
int ref_make_camera(const ref_camera_desc* c, int width, int height, romis_camera* out) {
    Window window("ref", glm::ivec2(width, height), OpenGLVersion::GL2, false);
    Trackball camera { &window, glm::radians(c->fov_deg), c->distance };
    camera.setCamera(g3(c->look_at), glm::radians(g3(c->rotation_deg)), c->distance);
    v3(out->origin, camera.position());
    glm::quat q = glm::quat(camera.rotationEulerAngles());
    out->quat[0] = q.w; out->quat[1] = q.x; out->quat[2] = q.y; out->quat[3] = q.z;
    out->half_height = std::tan(glm::radians(c->fov_deg) / 2.0f);               // trackball.cpp:26
    out->half_width  = window.getAspectRatio() * out->half_height;               // trackball.cpp:27
    return 0;
}

int ref_reset_history(void) { g_prev.reset(); return 0; }

// renderRMIS itself (reference src/rendering/render.cpp:64-119), called whole; the neighbour index grid is additionally
// produced by a separate call of generateResampleIndicesGrid under the same random-stream keys so that it can be dumped.
int ref_render_frame_rmis(const romis_features* f, const romis_rmis_params* rp, const ref_camera_desc* cam, int W, int H,
                          const romis_rng* rng, float* out_rgb, int32_t* neigh_xy, uint32_t* neigh_count) {
    if (!g_embree) { g_err = "no scene"; return -1; }
    try {
        Features features = toFeatures(*f);
        features.rayTraceMode = RayTraceMode::RMIS;
        features.maxIterationsMIS = rp->maxIterationsMIS;
        features.misWeightRMIS = static_cast<MISWeightRMIS>(rp->misWeightRMIS);
        features.neighbourSelectionStrategy = static_cast<NeighbourSelectionStrategy>(rp->neighbourSelectionStrategy);
        features.neighbourSameGeometry = rp->neighbourSameGeometry != 0;
        features.neighbourMaxDepthDifferenceFraction = rp->neighbourMaxDepthDifferenceFraction;
        features.neighbourMaxNormalAngleDifferenceRadians = rp->neighbourMaxNormalAngleDifferenceRadians;
        Window window("ref", glm::ivec2(W, H), OpenGLVersion::GL2, false);
        Screen screen(glm::ivec2(W, H), false);
        Trackball camera { &window, glm::radians(cam->fov_deg), cam->distance };
        camera.setCamera(g3(cam->look_at), glm::radians(g3(cam->rotation_deg)), cam->distance);
        ShimState& s = g_shim;
        s.mode = g_mis_timing ? SHIM_TIMING : SHIM_PARITY; s.seed = rng->seed; s.frame = rng->frame; s.W = W; s.H = H;
        s.N = (int)features.numSamplesInReservoir; s.k = (int)features.numNeighboursToSample;
#ifdef _OPENMP
        omp_set_num_threads(g_mis_timing ? omp_get_num_procs() : 1);
#endif
        NullBuf nb; std::streambuf* old = std::cout.rdbuf(&nb);
        const uint32_t K1 = features.numNeighboursToSample + 1U;
        if (neigh_xy || neigh_count) {
            s.stage_queue = { SHIM_STAGE_PRIMARY_THEN_NEIGH }; s.stage_pos = 0; s.stage = SHIM_STAGE_NONE;
            PrimaryHitGrid primaryHits = genPrimaryRayHits(g_scene, camera, *g_embree, screen, features);
            ResampleIndicesGrid grid = generateResampleIndicesGrid(primaryHits, screen.resolution(), features);
            for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
                const auto& v = grid[y][x]; size_t p = size_t(y) * W + x;
                if (neigh_count) neigh_count[p] = (uint32_t)v.size();
                if (neigh_xy) for (uint32_t i = 0; i < K1; i++) {
                    neigh_xy[(p * K1 + i) * 2 + 0] = i < v.size() ? v[i].x : -1;
                    neigh_xy[(p * K1 + i) * 2 + 1] = i < v.size() ? v[i].y : -1;
                }
            }
        }
        s.stage_queue.clear(); s.stage_pos = 0; s.stage = SHIM_STAGE_NONE;
        s.stage_queue.push_back(SHIM_STAGE_PRIMARY_THEN_NEIGH);                         // genPrimaryRayHits, then the index grid
        for (uint32_t it = 0; it < features.maxIterationsMIS; it++) {
            s.stage_queue.push_back(ROMIS_STAGE_RMIS_INITIAL0 + (int)it);               // genInitialSamples
            s.stage_queue.push_back(SHIM_STAGE_NONE);                                   // the gather loop's bar (render.cpp:75)
        }
        s.stage_queue.push_back(SHIM_STAGE_NONE);                                       // combineToScreen
        renderRMIS(g_scene, camera, *g_embree, screen, features);
        std::cout.rdbuf(old);
        if (!g_mis_timing && s.stage_pos != s.stage_queue.size()) { g_err = "stage queue not consumed"; return -2; }
        if (out_rgb) std::memcpy(out_rgb, screen.pixels().data(), size_t(W) * H * 3 * sizeof(float));
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

// 1: ref_render_frame_rmis / _romis run for TIMING: thread-safe non-parity random stream, OpenMP on all host cores
int ref_set_mis_timing(int on) { g_mis_timing = on != 0; return 0; }

// renderROMIS itself (reference src/rendering/render.cpp:121-265), called whole
int ref_render_frame_romis(const romis_features* f, const romis_rmis_params* rp, const ref_camera_desc* cam, int W, int H,
                           const romis_rng* rng, float* out_rgb, float* matrices, float* contributions) {
    if (!g_embree) { g_err = "no scene"; return -1; }
    try {
        Features features = toFeatures(*f);
        features.rayTraceMode = RayTraceMode::ROMIS;
        features.maxIterationsMIS = rp->maxIterationsMIS;
        features.neighbourSelectionStrategy = static_cast<NeighbourSelectionStrategy>(rp->neighbourSelectionStrategy);
        features.neighbourSameGeometry = rp->neighbourSameGeometry != 0;
        features.neighbourMaxDepthDifferenceFraction = rp->neighbourMaxDepthDifferenceFraction;
        features.neighbourMaxNormalAngleDifferenceRadians = rp->neighbourMaxNormalAngleDifferenceRadians;
        features.useProgressiveROMIS = rp->useProgressiveROMIS != 0;
        features.progressiveUpdateMod = rp->progressiveUpdateMod;
        features.saveAlphasVisualisation = (matrices || contributions);      // -> the capture hook above
        Window window("ref", glm::ivec2(W, H), OpenGLVersion::GL2, false);
        Screen screen(glm::ivec2(W, H), false);
        Trackball camera { &window, glm::radians(cam->fov_deg), cam->distance };
        camera.setCamera(g3(cam->look_at), glm::radians(g3(cam->rotation_deg)), cam->distance);
        ShimState& s = g_shim;
        s.mode = g_mis_timing ? SHIM_TIMING : SHIM_PARITY; s.seed = rng->seed; s.frame = rng->frame; s.W = W; s.H = H;
        s.N = (int)features.numSamplesInReservoir; s.k = (int)features.numNeighboursToSample;
#ifdef _OPENMP
        omp_set_num_threads(g_mis_timing ? omp_get_num_procs() : 1);
#endif
        NullBuf nb; std::streambuf* old = std::cout.rdbuf(&nb);
        s.stage_queue.clear(); s.stage_pos = 0; s.stage = SHIM_STAGE_NONE;
        s.stage_queue.push_back(SHIM_STAGE_PRIMARY_THEN_NEIGH);                         // genPrimaryRayHits, then the index grid
        for (uint32_t it = 0; it < features.maxIterationsMIS; it++) {
            s.stage_queue.push_back(ROMIS_STAGE_RMIS_INITIAL0 + (int)it);               // genInitialSamples
            s.stage_queue.push_back(SHIM_STAGE_NONE);                                   // the accumulation loop's bar (render.cpp:144)
        }
        s.stage_queue.push_back(SHIM_STAGE_NONE);                                       // combineToScreen / the summation loop's bar
        g_cap_matrices = matrices; g_cap_contrib = contributions;
        renderROMIS(g_scene, camera, *g_embree, screen, features);
        g_cap_matrices = nullptr; g_cap_contrib = nullptr;
        std::cout.rdbuf(old);
        if (!g_mis_timing && s.stage_pos != s.stage_queue.size()) { g_err = "stage queue not consumed"; return -2; }
        if (out_rgb) std::memcpy(out_rgb, screen.pixels().data(), size_t(W) * H * 3 * sizeof(float));
    } catch (const std::exception& e) { g_cap_matrices = nullptr; g_cap_contrib = nullptr; g_err = e.what(); return -1; }
    return 0;
}

#ifdef ROMIS_WITH_DROPIN
}  // extern "C"
// integration/render_restir_gpu.cpp compiled with -DROMIS_DROPIN_NAME=renderReSTIR_gpu
ReservoirGrid renderReSTIR_gpu(std::shared_ptr<ReservoirGrid> previousFrameGrid, const Scene& scene, const Trackball& camera,
                               const EmbreeInterface& embreeInterface, Screen& screen, const Features& features);
void renderRMIS_gpu(const Scene& scene, const Trackball& camera, const EmbreeInterface& embreeInterface, Screen& screen, const Features& features);
void renderROMIS_gpu(const Scene& scene, const Trackball& camera, const EmbreeInterface& embreeInterface, Screen& screen, const Features& features);
extern "C" void romis_dropin_set_rng(uint64_t seed, uint32_t frame);
extern "C" void romis_dropin_set_half_extents(float halfWidth, float halfHeight);
extern "C" void romis_dropin_release_screen(void);
// the Screens of this harness live for one call (the application's lives as long as its window): the drop-in's page-lock
// registration of Screen::pixels() must end before the vector is freed
struct ScreenPinGuard { ~ScreenPinGuard() { romis_dropin_release_screen(); } };
extern "C" {
// The reference's own Scene / Trackball / Screen / Features objects driven through the GPU drop-in.
int ref_render_frame_dropin(const romis_features* f, const ref_camera_desc* cam, int W, int H, int history_valid,
                            const romis_rng* rng, float* out_rgb) {
    if (!g_embree) { g_err = "no scene"; return -1; }
    try {
        const Features features = toFeatures(*f);
        Window window("ref", glm::ivec2(W, H), OpenGLVersion::GL2, false);
        Screen screen(glm::ivec2(W, H), false);
        Trackball camera { &window, glm::radians(cam->fov_deg), cam->distance };
        camera.setCamera(g3(cam->look_at), glm::radians(g3(cam->rotation_deg)), cam->distance);
        const float halfH = std::tan(glm::radians(cam->fov_deg) / 2.0f);            // trackball.cpp:26-27
        romis_dropin_set_half_extents(window.getAspectRatio() * halfH, halfH);
        romis_dropin_set_rng(rng->seed, rng->frame);
        static std::shared_ptr<ReservoirGrid> prevGpu;
        if (!history_valid) prevGpu.reset();
        ScreenPinGuard unpin;
        ReservoirGrid grid = renderReSTIR_gpu(prevGpu, g_scene, camera, *g_embree, screen, features);
        prevGpu = std::make_shared<ReservoirGrid>(grid);                            // main.cpp:165
        if (out_rgb) std::memcpy(out_rgb, screen.pixels().data(), size_t(W) * H * 3 * sizeof(float));
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}
// End-to-end timing of the drop-in as the application runs it: ONE Screen / Trackball living across the frames (main.cpp:56-65),
// renderReSTIR_gpu called frame after frame with the grid it returned, wall clock around every call.  pin = 0 leaves
// Screen::pixels() pageable.  edit_light != 0: one light's colour changes before every frame (ui.cpp:172-261).
extern "C" void romis_dropin_set_pin_screen(int on);
int ref_dropin_bench(const romis_features* f, const ref_camera_desc* cam, int W, int H, int frames, int pin, int edit_light, double* out_ms) {
    if (!g_embree) { g_err = "no scene"; return -1; }
    try {
        const Features features = toFeatures(*f);
        Window window("ref", glm::ivec2(W, H), OpenGLVersion::GL2, false);
        Screen screen(glm::ivec2(W, H), false);
        Trackball camera { &window, glm::radians(cam->fov_deg), cam->distance };
        camera.setCamera(g3(cam->look_at), glm::radians(g3(cam->rotation_deg)), cam->distance);
        const float halfH = std::tan(glm::radians(cam->fov_deg) / 2.0f);
        romis_dropin_set_half_extents(window.getAspectRatio() * halfH, halfH);
        romis_dropin_set_rng(0x5eed, 0);
        romis_dropin_set_pin_screen(pin);
        ScreenPinGuard unpin;
        std::shared_ptr<ReservoirGrid> prev;
        NullBuf nb; std::streambuf* old = std::cout.rdbuf(&nb);
        for (int fr = 0; fr < frames; fr++) {
            if (edit_light && !g_scene.lights.empty()) {
                auto& l = g_scene.lights[0];
                const float s = 1.0f + 1e-3f * float(1 + fr % 7);
                if (auto* p = std::get_if<PointLight>(&l)) p->color *= s;
                else if (auto* q = std::get_if<SegmentLight>(&l)) q->color0 *= s;
                else std::get<ParallelogramLight>(l).color0 *= s;
            }
            auto t0 = std::chrono::steady_clock::now();
            ReservoirGrid grid = renderReSTIR_gpu(prev, g_scene, camera, *g_embree, screen, features);
            prev = std::make_shared<ReservoirGrid>(std::move(grid));                // main.cpp:165
            out_ms[fr] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
        std::cout.rdbuf(old);
        romis_dropin_set_pin_screen(1);
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

// mode 0: renderRMIS_gpu, 1: renderROMIS_gpu -- the reference's own objects through the GPU bodies of the other two estimators
int ref_render_frame_mis_dropin(int mode, const romis_features* f, const romis_rmis_params* rp, const ref_camera_desc* cam, int W, int H,
                                const romis_rng* rng, float* out_rgb) {
    if (!g_embree) { g_err = "no scene"; return -1; }
    try {
        Features features = toFeatures(*f);
        features.rayTraceMode = mode ? RayTraceMode::ROMIS : RayTraceMode::RMIS;
        features.maxIterationsMIS = rp->maxIterationsMIS;
        features.misWeightRMIS = static_cast<MISWeightRMIS>(rp->misWeightRMIS);
        features.neighbourSelectionStrategy = static_cast<NeighbourSelectionStrategy>(rp->neighbourSelectionStrategy);
        features.neighbourSameGeometry = rp->neighbourSameGeometry != 0;
        features.neighbourMaxDepthDifferenceFraction = rp->neighbourMaxDepthDifferenceFraction;
        features.neighbourMaxNormalAngleDifferenceRadians = rp->neighbourMaxNormalAngleDifferenceRadians;
        features.useProgressiveROMIS = rp->useProgressiveROMIS != 0;
        features.progressiveUpdateMod = rp->progressiveUpdateMod;
        Window window("ref", glm::ivec2(W, H), OpenGLVersion::GL2, false);
        Screen screen(glm::ivec2(W, H), false);
        Trackball camera { &window, glm::radians(cam->fov_deg), cam->distance };
        camera.setCamera(g3(cam->look_at), glm::radians(g3(cam->rotation_deg)), cam->distance);
        const float halfH = std::tan(glm::radians(cam->fov_deg) / 2.0f);            // trackball.cpp:26-27
        romis_dropin_set_half_extents(window.getAspectRatio() * halfH, halfH);
        romis_dropin_set_rng(rng->seed, rng->frame);
        ScreenPinGuard unpin;
        if (mode) renderROMIS_gpu(g_scene, camera, *g_embree, screen, features);
        else renderRMIS_gpu(g_scene, camera, *g_embree, screen, features);
        if (out_rgb) std::memcpy(out_rgb, screen.pixels().data(), size_t(W) * H * 3 * sizeof(float));
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}
#endif
static int g_threads = 0;       // 0 = all processors
static int timing_threads() {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_num_procs();
#else
    return 1;
#endif
}
int ref_num_threads(void) { return timing_threads(); }
int ref_set_num_threads(int n) { g_threads = n > 0 ? n : 0; return 0; }

int ref_render_frame(const romis_features* f, const ref_camera_desc* cam, int W, int H, int history_valid,
                     const romis_rng* rng, int flags, ref_frame_dump* dump, float* out_rgb, ref_timings* tm) {
    if (!g_embree) { g_err = "no scene"; return -1; }
    try {
        const Features features = toFeatures(*f);
        const bool whole = (flags & REF_FLAG_WHOLE_FRAME) != 0;     // call renderReSTIR itself
        const bool asis = (flags & REF_FLAG_ASIS_RNG) != 0;         // the reference's own random sources (timing only)
        const bool timing = asis || (flags & REF_FLAG_TIMING_RNG) != 0;     // thread-safe non-parity RNG, OpenMP allowed
        Window window("ref", glm::ivec2(W, H), OpenGLVersion::GL2, false);
        Screen screen(glm::ivec2(W, H), false);
        Trackball camera { &window, glm::radians(cam->fov_deg), cam->distance };
        camera.setCamera(g3(cam->look_at), glm::radians(g3(cam->rotation_deg)), cam->distance);
        if (!history_valid) g_prev.reset();
        const bool doTemporal = features.temporalReuse && g_prev;

        ShimState& s = g_shim;
        s.mode = asis ? SHIM_ASIS : timing ? SHIM_TIMING : SHIM_PARITY;
        s.seed = rng->seed; s.frame = rng->frame; s.W = W; s.H = H;
        s.N = (int)features.numSamplesInReservoir; s.k = (int)features.numNeighboursToSample;
        s.stage_queue.clear(); s.stage_pos = 0; s.stage = SHIM_STAGE_NONE;
        s.stage_queue.push_back(SHIM_STAGE_NONE);                                   // genPrimaryRayHits
        s.stage_queue.push_back(ROMIS_STAGE_INITIAL);                               // genInitialSamples
        if (doTemporal) s.stage_queue.push_back(ROMIS_STAGE_TEMPORAL);              // temporalReuse
        if (features.spatialReuse) for (uint32_t p = 0; p < features.spatialResamplingPasses; p++)
            s.stage_queue.push_back(ROMIS_STAGE_SPATIAL0 + (int)p);                 // spatialReuse, one bar per pass
        s.stage_queue.push_back(SHIM_STAGE_NONE);                                   // final shading loop
#ifdef _OPENMP
        omp_set_num_threads(timing ? timing_threads() : 1);
#endif
        // the reference prints a banner per stage (render.cpp:32,40; render_utils.cpp:18,39,93,97,146)
        NullBuf nb; std::streambuf* old = std::cout.rdbuf(&nb);
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        auto t0 = now();
        ReservoirGrid grid;
        if (whole) {
            grid = renderReSTIR(g_prev, g_scene, camera, *g_embree, screen, features);
            if (tm) { std::memset(tm, 0, sizeof *tm); tm->total_ms = ms(t0, now()); }
        } else {
            PrimaryHitGrid primaryHits = genPrimaryRayHits(g_scene, camera, *g_embree, screen, features);
            auto t1 = now();
            if (dump && dump->gbuffer_t) {
                for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
                    const RayHit& rh = primaryHits[y][x]; size_t i = size_t(y) * W + x;
                    dump->gbuffer_t[i] = rh.ray.t;
                    if (dump->gbuffer_normal) v3(dump->gbuffer_normal + 3 * i, rh.hit.normal);
                    if (dump->gbuffer_texcoord) { dump->gbuffer_texcoord[2 * i] = rh.hit.texCoord.x; dump->gbuffer_texcoord[2 * i + 1] = rh.hit.texCoord.y; }
                    if (dump->gbuffer_mesh) dump->gbuffer_mesh[i] = rh.ray.t == std::numeric_limits<float>::max() ? (uint32_t)g_scene.meshes.size() : rh.hit.geometryId;
                    if (dump->ray_dir) v3(dump->ray_dir + 3 * i, rh.ray.direction);
                    if (dump->ray_origin) v3(dump->ray_origin + 3 * i, rh.ray.origin);
                }
            }
            grid = genInitialSamples(primaryHits, g_scene, *g_embree, features, screen.resolution());
            auto t2 = now();
            const int N = s.N;
            if (dump) dumpGrid(grid, W, H, N, dump->initial);
            if (doTemporal) temporalReuse(grid, *g_prev, *g_embree, screen, features);
            auto t3 = now();
            if (dump && doTemporal) dumpGrid(grid, W, H, N, dump->temporal);
            if (features.spatialReuse) {
                if (flags & REF_FLAG_SPLIT_SPATIAL) {
                    // one spatialReuse call per pass (same result: prevIteration is re-copied from the
                    // grid at the top of every call, render_utils.cpp:95,138) so each pass can be dumped
                    Features one = features; one.spatialResamplingPasses = 1U;
                    for (uint32_t p = 0; p < features.spatialResamplingPasses; p++) {
                        spatialReuse(grid, *g_embree, screen, one);
                        if (dump && p < 8) dumpGrid(grid, W, H, N, dump->spatial[p]);
                    }
                } else spatialReuse(grid, *g_embree, screen, features);
            }
            auto t4 = now();
            // final shading loop of renderReSTIR (render.cpp:38-58), restated because it is inline there
            {
                glm::ivec2 windowResolution = screen.resolution();
                romis_shim_stage_begin(windowResolution.y);
                #ifdef _OPENMP
                #pragma omp parallel for schedule(guided)
                #endif
                for (int y = 0; y < windowResolution.y; y++) {
                    for (int x = 0; x != windowResolution.x; x++) {
                        const Reservoir& reservoir = grid[y][x];
                        glm::vec3 finalColor = finalShading(reservoir, reservoir.cameraRay, *g_embree, features);
                        if (features.enableToneMapping) { finalColor = exposureToneMapping(finalColor, features); }
                        screen.setPixel(x, y, finalColor);
                    }
                }
            }
            auto t5 = now();
            if (tm) {
                tm->primary_ms = ms(t0, t1); tm->initial_ms = ms(t1, t2); tm->temporal_ms = ms(t2, t3);
                tm->spatial_ms = ms(t3, t4); tm->shade_ms = ms(t4, t5); tm->total_ms = ms(t0, t5);
            }
        }
        std::cout.rdbuf(old);
        if (dump) dumpGrid(grid, W, H, s.N, dump->final_);
        if (out_rgb) std::memcpy(out_rgb, screen.pixels().data(), size_t(W) * H * 3 * sizeof(float));
        auto tc = now();
        g_prev = std::make_shared<ReservoirGrid>(std::move(grid));     // main.cpp:165
        if (tm) tm->grid_copy_ms = ms(tc, now());
        if (s.mode == SHIM_PARITY && s.stage_pos != s.stage_queue.size()) { g_err = "stage queue not consumed"; return -2; }
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
    return 0;
}

}  // extern "C"
