/*
 * shims.cpp -- everything the reference's hot-path translation units need at link time that is
 * not in /root/reference or cannot run headless:
 *   (1) EmbreeInterface member definitions (replaces src/ray_tracing/embree_interface.cpp, which
 *       calls Intel Embree 4.3.1 -- absent) backed by oracle/tracer.c;
 *   (2) a headless Window (replaces framework/src/window.cpp, needs GLFW + a display);
 *   (3) no-op debug drawing (replaces src/ui/draw.cpp, needs legacy GL);
 *   (4) rand(), powf(), expf() and the engine hooks of rng_shim.h: the injected counter-based
 *       random stream (include/romis_rng.h) and the shared deterministic pow/exp
 *       (include/romis_detmath.h).
 * TEST ORACLE ONLY.  Built by oracle/Makefile into oracle/_ref/libromis_ref.so.
 */
#include <ray_tracing/embree_interface.h>
#include <framework/window.h>
#include <ui/draw.h>
#include <utils/utils.h>

#include <atomic>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../tracer.h"
#include "romis_rng.h"
#include "romis_detmath.h"
#include "shim_state.h"

// ------------------------------------------------------------------------------------------------
// (4) random stream state machine
// ------------------------------------------------------------------------------------------------
ShimState g_shim;

static thread_local uint64_t tl_fast_state = 0;
static std::atomic<uint64_t> g_fast_seed{0x9e3779b97f4a7c15ull};
static inline uint32_t fast_next() {   // timing mode: thread-local splitmix64, no parity
    if (tl_fast_state == 0) tl_fast_state = g_fast_seed.fetch_add(0x632be59bd9b4e019ull) | 1ull;
    uint64_t z = (tl_fast_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return uint32_t((z ^ (z >> 31)) >> 32);
}

static void shim_fail(const char* what) {
    std::fprintf(stderr, "[ref_harness] random-stream tracking error: %s (stage %d pixel %ld)\n", what, g_shim.stage, g_shim.pixel);
    std::abort();
}

extern "C" void romis_shim_stage_begin(int rows) {
    ShimState& s = g_shim;
    (void)rows;
    if (s.mode != SHIM_PARITY) return;
    if (s.stage_pos >= s.stage_queue.size()) shim_fail("unexpected stage (progressbar constructed with empty stage queue)");
    s.stage = s.stage_queue[s.stage_pos++];
    s.pixel = -1; s.engine_ctr = 0; s.rand_ctr = 0; s.engine_total = 0; s.rand_total = 0; s.rows_done = 0;
}

extern "C" void romis_shim_row_done(void) {
    ShimState& s = g_shim;
    if (s.mode != SHIM_PARITY) return;
    s.rows_done++;
}

// stages in which the reference constructs one std::mt19937 per pixel (genCanonicalSamples, light.cpp:49-50; the R-MIS
// neighbour selection, neighbour_selection.cpp:28-29,72-73): the constructor marks the pixel boundary
static bool per_pixel_engine_stage(int stage) {
    return stage == ROMIS_STAGE_INITIAL || stage == ROMIS_STAGE_RMIS_NEIGH ||
           (stage >= ROMIS_STAGE_RMIS_INITIAL0 && stage < ROMIS_STAGE_RMIS_INITIAL0 + 64);
}

extern "C" void romis_shim_engine_ctor(void) {
    ShimState& s = g_shim;
    if (s.mode != SHIM_PARITY) return;
    // renderRMIS builds its neighbour index grid right after the primary rays without a progress bar of its own
    // (render.cpp:68-69): the first engine constructed in that gap opens the neighbour-selection stage
    if (s.stage == SHIM_STAGE_PRIMARY_THEN_NEIGH) { s.stage = ROMIS_STAGE_RMIS_NEIGH; s.pixel = -1; }
    if (per_pixel_engine_stage(s.stage)) { s.pixel++; s.engine_ctr = 0; s.rand_ctr = 0; }
}

extern "C" uint32_t romis_shim_engine_next(void) {
    ShimState& s = g_shim;
    if (s.mode != SHIM_PARITY) return fast_next();
    if (per_pixel_engine_stage(s.stage)) {
        if (s.pixel < 0 || s.pixel >= (long)s.W * s.H) shim_fail("engine draw outside a pixel");
        romis_stream_key k = romis_rng_stream(s.seed, s.frame, (uint32_t)s.stage, (uint32_t)s.pixel, ROMIS_STREAM_ENGINE);
        return romis_rng_bits(k, s.engine_ctr++);
    }
    if (s.stage >= ROMIS_STAGE_SPATIAL0 && s.stage < ROMIS_STAGE_RMIS_NEIGH) {
        if (s.k <= 0) shim_fail("engine draw with k == 0");
        long pix = s.engine_total / (2L * s.k);
        uint32_t c = (uint32_t)(s.engine_total % (2L * s.k));
        if (c == 0) { s.pixel = pix; s.rand_ctr = 0; }
        s.engine_total++;
        if (pix >= (long)s.W * s.H) shim_fail("more neighbour draws than pixels");
        romis_stream_key k = romis_rng_stream(s.seed, s.frame, (uint32_t)s.stage, (uint32_t)pix, ROMIS_STREAM_ENGINE);
        return romis_rng_bits(k, c);
    }
    shim_fail("engine draw in a stage that has none");
    return 0;
}

extern "C" int romis_shim_asis(void) { return g_shim.mode == SHIM_ASIS; }

extern "C" int rand(void) {
    ShimState& s = g_shim;
    if (s.mode == SHIM_ASIS) {      // glibc's rand(): one process-wide state behind a lock (what the reference calls)
        static int (*libc_rand)(void) = reinterpret_cast<int (*)(void)>(dlsym(RTLD_NEXT, "rand"));
        return libc_rand();
    }
    if (s.mode != SHIM_PARITY) return int(fast_next() >> 1);
    long pix; uint32_t c;
    if (per_pixel_engine_stage(s.stage)) {
        pix = s.pixel; c = s.rand_ctr++;
    } else if (s.stage == ROMIS_STAGE_TEMPORAL) {
        pix = s.rand_total / (2L * s.N); c = (uint32_t)(s.rand_total % (2L * s.N)); s.rand_total++;
    } else if (s.stage >= ROMIS_STAGE_SPATIAL0 && s.stage < ROMIS_STAGE_RMIS_NEIGH) {
        if (s.k > 0) { pix = s.pixel; c = s.rand_ctr++; }
        else { pix = s.rand_total / s.N; c = (uint32_t)(s.rand_total % s.N); s.rand_total++; }
    } else { shim_fail("rand() in a stage that has none"); return 0; }
    if (pix < 0 || pix >= (long)s.W * s.H) shim_fail("rand() outside a pixel");
    romis_stream_key k = romis_rng_stream(s.seed, s.frame, (uint32_t)s.stage, (uint32_t)pix, ROMIS_STREAM_RAND);
    return romis_rng_rand(k, c);
}

// shared deterministic pow/exp instead of libm's (romis_detmath.h); -Bsymbolic binds the
// reference's std::pow(float,float) / std::exp(float) calls to these.
std::atomic<long> g_powf_calls{0};
extern "C" float powf(float x, float y) { g_powf_calls.fetch_add(1, std::memory_order_relaxed); return romis_powf(x, y); }
extern "C" float expf(float x) { return romis_expf(x); }

// ------------------------------------------------------------------------------------------------
// (1) EmbreeInterface on top of oracle/tracer.c
// ------------------------------------------------------------------------------------------------
struct ShimTracerScene {
    otr_tracer* tracer = nullptr;
    std::vector<uint32_t> triMesh;          // global triangle -> mesh (geomID)
    std::vector<Vertex> triVerts;           // 3 per global triangle
};
int g_tracer_mode = 1;                      // 0 brute force, 1 BVH

static ShimTracerScene* asShim(RTCScene s) { return reinterpret_cast<ShimTracerScene*>(s); }

EmbreeInterface::EmbreeInterface(const Scene& scene) { initDevice(); initScene(scene); }
EmbreeInterface::~EmbreeInterface() {
    ShimTracerScene* s = asShim(m_scene);
    if (s) { otr_free(s->tracer); delete s; }
}
void EmbreeInterface::initDevice() { m_device = nullptr; }
void EmbreeInterface::initScene(const Scene& scene) {
    ShimTracerScene* s = new ShimTracerScene();
    std::vector<float> verts;
    uint32_t geomId = 0;
    m_meshToMaterial.clear();
    for (const Mesh& mesh : scene.meshes) {             // one geometry per Mesh, geomID = attach order
        for (const glm::uvec3& tri : mesh.triangles) {
            for (int c = 0; c < 3; c++) {
                const Vertex& v = mesh.vertices[tri[c]];
                verts.push_back(v.position.x); verts.push_back(v.position.y); verts.push_back(v.position.z);
                s->triVerts.push_back(v);
            }
            s->triMesh.push_back(geomId);
        }
        m_meshToMaterial[geomId] = mesh.material;
        geomId++;
    }
    s->tracer = otr_build(verts.data(), (int)s->triMesh.size(), g_tracer_mode);
    m_scene = reinterpret_cast<RTCScene>(s);
}
void EmbreeInterface::changeScene(const Scene& scene) {
    ShimTracerScene* s = asShim(m_scene);
    if (s) { otr_free(s->tracer); delete s; }
    initScene(scene);
}
bool EmbreeInterface::anyHit(Ray& ray) const {          // rtcOccluded1 with tnear = 0, tfar = ray.t
    const ShimTracerScene* s = asShim(m_scene);
    const float o[3] = {ray.origin.x, ray.origin.y, ray.origin.z};
    const float d[3] = {ray.direction.x, ray.direction.y, ray.direction.z};
    return otr_any(s->tracer, o, d, ray.t) != 0;
}
bool EmbreeInterface::closestHit(Ray& ray, HitInfo& hitInfo) const {   // rtcIntersect1 + 3x rtcInterpolate0
    const ShimTracerScene* s = asShim(m_scene);
    const float o[3] = {ray.origin.x, ray.origin.y, ray.origin.z};
    const float d[3] = {ray.direction.x, ray.direction.y, ray.direction.z};
    float t, u, v; uint32_t tri;
    if (!otr_closest(s->tracer, o, d, ray.t, &t, &u, &v, &tri)) return false;   // hitInfo untouched, ray.t unchanged
    const Vertex& a = s->triVerts[3 * tri + 0];
    const Vertex& b = s->triVerts[3 * tri + 1];
    const Vertex& c = s->triVerts[3 * tri + 2];
    const float w = (1.0f - u) - v;
    // attribute interpolation, defined as (w*a + u*b) + v*c per component
    hitInfo.normal           = (w * a.normal + u * b.normal) + v * c.normal;
    hitInfo.barycentricCoord = (w * a.position + u * b.position) + v * c.position;
    hitInfo.texCoord         = (w * a.texCoord + u * b.texCoord) + v * c.texCoord;
    const uint32_t geomId    = s->triMesh[tri];
    hitInfo.material         = m_meshToMaterial.at(geomId);
    hitInfo.geometryId       = geomId;
    ray.t                    = t;
    return true;
}
void EmbreeInterface::populateVertexDataBuffers(glm::vec3*, glm::vec3*, glm::vec2*, const std::vector<Vertex>&) {}
void EmbreeInterface::populateIndexBuffer(glm::uvec3*, const std::vector<glm::uvec3>&) {}
RTCRayHit EmbreeInterface::constructEmbreeRay(const Ray&) const { return RTCRayHit{}; }

// ------------------------------------------------------------------------------------------------
// (2) headless Window: only what Trackball touches (framework/src/trackball.cpp:21-48,105-114)
// ------------------------------------------------------------------------------------------------
Window::Window(std::string_view, const glm::ivec2& windowSize, OpenGLVersion glVersion, bool presentable)
    : m_pWindow(nullptr), m_windowSize(windowSize), m_glVersion(glVersion), m_presentable(presentable) {}
Window::~Window() {}
void Window::registerMouseButtonCallback(MouseButtonCallback&& cb) { m_mouseButtonCallbacks.push_back(std::move(cb)); }
void Window::registerMouseMoveCallback(MouseMoveCallback&& cb) { m_mouseMoveCallbacks.push_back(std::move(cb)); }
void Window::registerScrollCallback(ScrollCallback&& cb) { m_scrollCallbacks.push_back(std::move(cb)); }
void Window::registerWindowResizeCallback(WindowResizeCallback&& cb) { m_windowResizeCallbacks.push_back(std::move(cb)); }
bool Window::isMouseButtonPressed(int) const { return false; }
glm::vec2 Window::getCursorPos() const { return glm::vec2(0.0f); }
float Window::getAspectRatio() const {                  // framework/src/window.cpp:380-385
    if (m_windowSize.x == 0 || m_windowSize.y == 0) return 1.0f;
    return float(m_windowSize.x) / float(m_windowSize.y);
}

// ------------------------------------------------------------------------------------------------
// (3) debug drawing: off
// ------------------------------------------------------------------------------------------------
bool enableDebugDraw = false;
void drawRay(const Ray&, const glm::vec3&) {}
void drawSphere(const Sphere&) {}
void drawSphere(const glm::vec3&, float, const glm::vec3&) {}
void drawScene(const Scene&) {}
void drawMesh(const Mesh&) {}
