// render_restir_gpu.cpp -- the reference-side host code of the drop-in: a replacement BODY for
//
//     ReservoirGrid renderReSTIR(std::shared_ptr<ReservoirGrid> previousFrameGrid, const Scene& scene,
//                                const Trackball& camera, const EmbreeInterface& embreeInterface,
//                                Screen& screen, const Features& features);      (reference src/rendering/render.h:25-28)
//
// written against the reference's OWN headers (Scene, Mesh, Trackball, Screen, Features, ReservoirGrid) and the C-ABI of
// include/romis_gpu.h.  A maintainer drops this file into src/rendering/, removes the body of renderReSTIR from
// render.cpp (lines 28-62) and links libromis_gpu.so; nothing else in the renderer, UI or scene loader changes
// (INTEGRATION.md).  In this repo it is compiled by `make -C oracle dropin` into oracle/_ref/libromis_dropin.so together
// with the reference's translation units, and tests/test_gpu_dropin.py checks that the reference's own Scene / Trackball
// / Screen objects driven through it produce the reference's image bit for bit.
//
// Compile with -DROMIS_DROPIN_NAME=<symbol> to emit the function under another name (the test library keeps the
// reference's CPU renderReSTIR next to it).
#include <rendering/render.h>
#include <rendering/reservoir.h>
#include <rendering/screen.h>
#include <scene/scene.h>
#include <utils/common.h>
#include <framework/trackball.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "romis_gpu.h"

#ifndef ROMIS_DROPIN_NAME
#define ROMIS_DROPIN_NAME renderReSTIR
#endif

namespace {

struct GpuState {
    romis_ctx* ctx = nullptr;
    const void* sceneKey = nullptr;     // identity of the uploaded geometry
    uint64_t sceneSig = 0;
    bool sceneDirty = true;             // romis_dropin_invalidate_scene (hook next to EmbreeInterface::changeScene)
    uint64_t seed = 0x524f4d4953ull;    // "ROMIS"
    uint32_t frame = 0;
    // lights: a persistent POD copy of scene.lights; with the UI hook (romis_dropin_lights_dirty) only the marked range is
    // converted and examined, without it the whole table is converted and compared every frame
    std::vector<romis_light> lights;
    bool lightsHooked = false, lightsAllDirty = true;
    int dirtyFirst = 0, dirtyEnd = 0;
    // Screen::pixels() storage registered as page-locked memory, so the read-back is an asynchronous DMA into it
    void* pinnedPtr = nullptr; size_t pinnedBytes = 0; bool pinScreen = true;
    ~GpuState() {
        if (pinnedPtr) romis_host_unregister(pinnedPtr);
        if (ctx) romis_destroy(ctx);
    }
};
// one context per calling thread: the reference's CLI mode renders one camera per std::thread (main.cpp:213-230)
thread_local GpuState g;

void check(int rc, const char* what) {
    if (rc != ROMIS_OK) throw std::runtime_error(std::string(what) + ": " + romis_last_error(g.ctx));   // render.cpp:278 throws too
}

// FNV-1a over everything romis_upload_scene consumes: vertices, indices, materials, texture identity and size.  One pass over
// the geometry per frame (a few hundred KB for the reference's scenes) against a frame of rendering; a maintainer who hooks
// romis_dropin_invalidate_scene next to EmbreeInterface::changeScene (embree_interface.cpp:53-56) can drop it.
uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
uint64_t geometrySignature(const Scene& scene) {
    uint64_t h = 1469598103934665603ull;
    const size_t nm = scene.meshes.size();
    h = fnv(h, &nm, sizeof nm);
    for (const Mesh& m : scene.meshes) {
        const size_t nv = m.vertices.size(), nt = m.triangles.size();
        h = fnv(h, &nv, sizeof nv); h = fnv(h, &nt, sizeof nt);
        if (nv) h = fnv(h, m.vertices.data(), nv * sizeof(Vertex));
        if (nt) h = fnv(h, m.triangles.data(), nt * sizeof(glm::uvec3));
        h = fnv(h, &m.material.kd.x, 12); h = fnv(h, &m.material.ks.x, 12);
        h = fnv(h, &m.material.shininess, 4); h = fnv(h, &m.material.transparency, 4);
        const Image* img = m.material.kdTexture.get();
        h = fnv(h, &img, sizeof img);
        if (img) { h = fnv(h, &img->width, sizeof img->width); h = fnv(h, &img->height, sizeof img->height); }
    }
    return h;
}

void uploadScene(romis_ctx* ctx, const Scene& scene) {
    std::vector<romis_mesh_desc> descs(scene.meshes.size());
    std::vector<std::vector<uint32_t>> tris(scene.meshes.size());
    std::vector<const Image*> images;
    std::vector<std::vector<float>> pixels;
    std::vector<romis_texture> textures;
    static_assert(sizeof(Vertex) == sizeof(romis_vertex), "Vertex {vec3 position, vec3 normal, vec2 texCoord} maps 1:1 (mesh.h:14-20)");
    for (size_t i = 0; i < scene.meshes.size(); i++) {
        const Mesh& m = scene.meshes[i];
        for (const glm::uvec3& t : m.triangles) { tris[i].push_back(t.x); tris[i].push_back(t.y); tris[i].push_back(t.z); }
        romis_mesh_desc& d = descs[i];
        d.vertices = reinterpret_cast<const romis_vertex*>(m.vertices.data()); d.n_vertices = (uint32_t)m.vertices.size();
        d.triangles = tris[i].data(); d.n_triangles = (uint32_t)m.triangles.size();
        std::memcpy(d.material.kd, &m.material.kd.x, 12); std::memcpy(d.material.ks, &m.material.ks.x, 12);
        d.material.shininess = m.material.shininess; d.material.transparency = m.material.transparency;
        d.material.kd_texture = -1;
        if (m.material.kdTexture) {
            const Image* img = m.material.kdTexture.get();
            size_t k = 0;
            while (k < images.size() && images[k] != img) k++;
            if (k == images.size()) {
                images.push_back(img);
                pixels.emplace_back(3 * img->pixels.size());
                std::memcpy(pixels.back().data(), img->pixels.data(), 12 * img->pixels.size());
            }
            d.material.kd_texture = (int32_t)k;
        }
    }
    for (size_t k = 0; k < images.size(); k++) textures.push_back(romis_texture { pixels[k].data(), images[k]->width, images[k]->height });
    check(romis_upload_scene(ctx, descs.data(), (int)descs.size(), textures.data(), (int)textures.size()), "romis_upload_scene");
}

romis_light toPod(const std::variant<PointLight, SegmentLight, ParallelogramLight>& v) {
    romis_light l; std::memset(&l, 0, sizeof l);
    if (std::holds_alternative<PointLight>(v)) {
        const PointLight& p = std::get<PointLight>(v); l.type = ROMIS_LIGHT_POINT;
        std::memcpy(l.p0, &p.position.x, 12); std::memcpy(l.c0, &p.color.x, 12);
    } else if (std::holds_alternative<SegmentLight>(v)) {
        const SegmentLight& s = std::get<SegmentLight>(v); l.type = ROMIS_LIGHT_SEGMENT;
        std::memcpy(l.p0, &s.endpoint0.x, 12); std::memcpy(l.e1, &s.endpoint1.x, 12);
        std::memcpy(l.c0, &s.color0.x, 12); std::memcpy(l.c1, &s.color1.x, 12);
    } else {
        const ParallelogramLight& p = std::get<ParallelogramLight>(v); l.type = ROMIS_LIGHT_PARALLELOGRAM;
        std::memcpy(l.p0, &p.v0.x, 12); std::memcpy(l.e1, &p.edge01.x, 12); std::memcpy(l.e2, &p.edge02.x, 12);
        std::memcpy(l.c0, &p.color0.x, 12); std::memcpy(l.c1, &p.color1.x, 12);
        std::memcpy(l.c2, &p.color2.x, 12); std::memcpy(l.c3, &p.color3.x, 12);
    }
    return l;
}

// scene.lights is read fresh every frame by the reference (light.cpp:46-66) and edited by the UI without notification
// (ui.cpp:172-261).  Unhooked: convert and hand over the whole table, romis_upload_lights finds what changed.  Hooked (the UI
// calls romis_dropin_lights_dirty after an edit): only the marked lights are converted and examined, an untouched table costs
// nothing.  Either way the history keeps its samples as drawn (romis_gpu.h, romis_upload_lights).
void uploadLights(const Scene& scene) {
    const size_t n = scene.lights.size();
    const bool all = !g.lightsHooked || g.lightsAllDirty || g.lights.size() != n;
    if (all) {
        g.lights.resize(n);
        for (size_t i = 0; i < n; i++) g.lights[i] = toPod(scene.lights[i]);
        check(romis_upload_lights(g.ctx, g.lights.data(), (int)n), "romis_upload_lights");
    } else if (g.dirtyEnd > g.dirtyFirst) {
        const int a = std::max(0, g.dirtyFirst), b = std::min((int)n, g.dirtyEnd);
        for (int i = a; i < b; i++) g.lights[i] = toPod(scene.lights[i]);
        check(romis_upload_lights_range(g.ctx, g.lights.data(), (int)n, a, std::max(0, b - a)), "romis_upload_lights_range");
    }
    g.lightsAllDirty = false; g.dirtyFirst = g.dirtyEnd = 0;
}

// Screen::pixels() is a std::vector<glm::vec3> (pageable): registered once as page-locked memory, the device-to-host copy of
// the image becomes an asynchronous DMA that overlaps the shading of the next rows (romis_frame_end)
float* screenStorage(Screen& screen) {
    std::vector<glm::vec3>& px = screen.pixels();
    void* p = px.data(); const size_t bytes = px.size() * sizeof(glm::vec3);
    if (!g.pinScreen) return &px[0].x;
    if (p != g.pinnedPtr || bytes != g.pinnedBytes) {
        if (g.pinnedPtr) romis_host_unregister(g.pinnedPtr);
        g.pinnedPtr = nullptr; g.pinnedBytes = 0;
        if (romis_host_register(p, bytes) == ROMIS_OK) { g.pinnedPtr = p; g.pinnedBytes = bytes; }     // failure: pageable still works
    }
    return &px[0].x;
}

romis_features toPod(const Features& f) {
    romis_features o;
    o.enableShading = f.enableShading; o.enableTextureMapping = f.enableTextureMapping;
    o.initialSamplesVisibilityCheck = f.initialSamplesVisibilityCheck;
    o.numSamplesInReservoir = f.numSamplesInReservoir; o.initialLightSamples = f.initialLightSamples;
    o.numNeighboursToSample = f.numNeighboursToSample; o.spatialResampleRadius = f.spatialResampleRadius;
    o.unbiasedCombination = f.unbiasedCombination; o.spatialReuse = f.spatialReuse;
    o.spatialReuseVisibilityCheck = f.spatialReuseVisibilityCheck; o.temporalReuse = f.temporalReuse;
    o.spatialResamplingPasses = f.spatialResamplingPasses; o.temporalClampM = f.temporalClampM;
    o.enableToneMapping = f.enableToneMapping; o.gamma = f.gamma; o.exposure = f.exposure;
    return o;
}

}  // namespace

// Random stream of the next frame (parity runs pin it; the interactive renderer can leave the defaults).
extern "C" void romis_dropin_set_rng(uint64_t seed, uint32_t frame) { g.seed = seed; g.frame = frame; }

// Optional hooks for the maintainer.  romis_dropin_invalidate_scene: next to EmbreeInterface::changeScene (ui.cpp:104-106).
// romis_dropin_lights_dirty: after the UI edited lights [first, first + count) (ui.cpp:172-261; count < 0 = all of them,
// e.g. after adding / removing one); once called, frames no longer compare the whole table.
extern "C" void romis_dropin_invalidate_scene(void) { g.sceneDirty = true; }
extern "C" void romis_dropin_lights_dirty(int first, int count) {
    g.lightsHooked = true;
    if (count < 0) { g.lightsAllDirty = true; return; }
    if (g.dirtyEnd <= g.dirtyFirst) { g.dirtyFirst = first; g.dirtyEnd = first + count; }
    else { g.dirtyFirst = std::min(g.dirtyFirst, first); g.dirtyEnd = std::max(g.dirtyEnd, first + count); }
}

// The Screen's pixel storage stays page-locked between frames; call this before a Screen whose frames went through the GPU path
// is destroyed or resized (the application's one Screen lives as long as its window, main.cpp:56-65: at exit).
extern "C" void romis_dropin_release_screen(void) {
    if (g.pinnedPtr) romis_host_unregister(g.pinnedPtr);
    g.pinnedPtr = nullptr; g.pinnedBytes = 0;
}

// measurement knob: leave Screen::pixels() pageable (what the read-back costs without the registration)
extern "C" void romis_dropin_set_pin_screen(int on) { if (!on) romis_dropin_release_screen(); g.pinScreen = on != 0; }

// half extents of the image plane: Trackball keeps them private (trackball.h:55-56).  The maintainer either adds two
// accessors or, as here, the caller provides them; they are tan(fovy/2) and aspect*tan(fovy/2) (trackball.cpp:26-27).
static thread_local float g_halfW = 0.0f, g_halfH = 0.0f;
extern "C" void romis_dropin_set_half_extents(float halfWidth, float halfHeight) { g_halfW = halfWidth; g_halfH = halfHeight; }

// Context, scene / light upload and camera, common to the three entry points
static romis_camera prepare(const Scene& scene, const Trackball& camera) {
    if (!g.ctx) {
        // ROMIS_DEVICES=0,1,2,3 in the environment: one row band per listed GPU, all driven from this thread; default: GPU 0
        std::vector<int> devs;
        if (const char* e = std::getenv("ROMIS_DEVICES")) {
            for (const char* p = e; *p;) { char* end; long v = std::strtol(p, &end, 10); if (end == p) break; devs.push_back((int)v); p = *end ? end + 1 : end; }
        }
        if (devs.empty()) devs.push_back(0);
        if (romis_create(devs.data(), (int)devs.size(), &g.ctx) != ROMIS_OK) throw std::runtime_error(std::string("romis_create: ") + romis_last_error(nullptr));
    }
    // EmbreeInterface::changeScene (embree_interface.cpp:53-56) has no notification we could hook: detect geometry changes
    const uint64_t sig = geometrySignature(scene);
    if (g.sceneDirty || g.sceneKey != scene.meshes.data() || g.sceneSig != sig) {
        uploadScene(g.ctx, scene); g.sceneKey = scene.meshes.data(); g.sceneSig = sig; g.sceneDirty = false;
    }
    uploadLights(scene);
    romis_camera cam;
    const glm::vec3 pos = camera.position();                                // trackball.cpp:75-78
    const glm::quat q = glm::quat(camera.rotationEulerAngles());            // same expression generateRay uses (trackball.cpp:111)
    cam.origin[0] = pos.x; cam.origin[1] = pos.y; cam.origin[2] = pos.z;
    cam.quat[0] = q.w; cam.quat[1] = q.x; cam.quat[2] = q.y; cam.quat[3] = q.z;
    cam.half_width = g_halfW; cam.half_height = g_halfH;
    if (g_halfH == 0.0f) throw std::runtime_error("romis drop-in: image-plane half extents not set (romis_dropin_set_half_extents)");
    return cam;
}

ReservoirGrid ROMIS_DROPIN_NAME(std::shared_ptr<ReservoirGrid> previousFrameGrid,
                                const Scene& scene, const Trackball& camera,
                                const EmbreeInterface& /*embreeInterface: the GPU path owns its own BVH*/, Screen& screen,
                                const Features& features) {
    const romis_camera cam = prepare(scene, camera);
    const glm::ivec2 res = screen.resolution();
    const romis_features f = toPod(features);
    romis_rng rng { g.seed, g.frame++, 0 };
    // Screen::pixels() is the row-flipped float RGB framebuffer setPixel writes (screen.cpp:37-43,110-118)
    float* out = screenStorage(screen);
    check(romis_render_frame(g.ctx, &f, &cam, res.x, res.y, previousFrameGrid ? 1 : 0, &rng, out), "romis_render_frame");

    // The reservoir grid lives on the device.  The caller only tests the returned grid for presence and hands a copy back
    // next frame (main.cpp:165), so a 1x1 token is enough to carry "history exists".
    return ReservoirGrid(1, std::vector<Reservoir>(1, Reservoir(features.numSamplesInReservoir)));
}

// ---- the other two estimators behind renderRayTraced (render.cpp:273-276): replacement bodies for
//      void renderRMIS (render.h:29-30, render.cpp:64-119) and void renderROMIS (render.h:31-32, render.cpp:121-265) ----
#ifndef ROMIS_DROPIN_RMIS_NAME
#define ROMIS_DROPIN_RMIS_NAME renderRMIS
#endif
#ifndef ROMIS_DROPIN_ROMIS_NAME
#define ROMIS_DROPIN_ROMIS_NAME renderROMIS
#endif

static romis_rmis_params toMisPod(const Features& f) {
    romis_rmis_params p;
    p.maxIterationsMIS = f.maxIterationsMIS;
    p.misWeightRMIS = (uint32_t)f.misWeightRMIS;                            // Equal = 0, Balance = 1 (common.h:31-34)
    p.neighbourSelectionStrategy = (uint32_t)f.neighbourSelectionStrategy;  // Random, Similar, Dissimilar, EqualSimilarDissimilar (common.h:36-41)
    p.neighbourSameGeometry = f.neighbourSameGeometry;
    p.neighbourMaxDepthDifferenceFraction = f.neighbourMaxDepthDifferenceFraction;
    p.neighbourMaxNormalAngleDifferenceRadians = f.neighbourMaxNormalAngleDifferenceRadians;
    p.useProgressiveROMIS = f.useProgressiveROMIS;
    p.progressiveUpdateMod = f.progressiveUpdateMod;
    return p;
}

void ROMIS_DROPIN_RMIS_NAME(const Scene& scene, const Trackball& camera, const EmbreeInterface&, Screen& screen, const Features& features) {
    const romis_camera cam = prepare(scene, camera);
    const glm::ivec2 res = screen.resolution();
    const romis_features f = toPod(features);
    const romis_rmis_params p = toMisPod(features);
    romis_rng rng { g.seed, g.frame++, 0 };
    romis_ctx* ctx = g.ctx;             // one device, or one row band per device under ROMIS_DEVICES: same call
    if (romis_render_frame_rmis(ctx, &f, &p, &cam, res.x, res.y, &rng, screenStorage(screen)) != ROMIS_OK)
        throw std::runtime_error(std::string("romis_render_frame_rmis: ") + romis_last_error(ctx));
}

// saveAlphasVisualisation (render.cpp:227-229, BMP dumps of the per-technique alphas) is a debugging aid of the CPU path
// and is not produced here.
void ROMIS_DROPIN_ROMIS_NAME(const Scene& scene, const Trackball& camera, const EmbreeInterface&, Screen& screen, const Features& features) {
    const romis_camera cam = prepare(scene, camera);
    const glm::ivec2 res = screen.resolution();
    const romis_features f = toPod(features);
    const romis_rmis_params p = toMisPod(features);
    romis_rng rng { g.seed, g.frame++, 0 };
    romis_ctx* ctx = g.ctx;
    if (romis_render_frame_romis(ctx, &f, &p, &cam, res.x, res.y, &rng, screenStorage(screen)) != ROMIS_OK)
        throw std::runtime_error(std::string("romis_render_frame_romis: ") + romis_last_error(ctx));
}
