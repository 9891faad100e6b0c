// romis_gpu.cu -- implementation of the C-ABI declared in include/romis_gpu.h.
//
// Host side of the B200 ReSTIR path: context, scene/light upload (host BVH build, flattening into
// device tables), the per-frame launch sequence (one kernel per pass, kernels.cuh), row-band state for
// multi-GPU sharding, parity read-backs and CUDA-event timing.  No CPU rendering path exists here: if
// CUDA is unusable romis_create fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "romis_gpu.h"
#include "launch.hpp"
#define ROMIS_COD_MAX_DIM 11     // include/romis_cod.h ROMIS_COD_MAX

using namespace romis;

namespace {
std::mutex g_err_mutex;
std::string g_create_err;

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    cudaError_t ensure(size_t n) {
        if (n <= bytes && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, n ? n : 16);
        if (e == cudaSuccess) bytes = n ? n : 16;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};
}  // namespace

#define ROMIS_FINE_STAGES 65

struct romis_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // scene
    bool has_scene = false;
    DevBuf nodes, tri_geom, tri_attr, materials, tex_pixels, tex_desc, lights;
    SceneDev sc{};
    int n_tris = 0;
    int n_sms = 148;

    // frame geometry
    int W = 0, H = 0, N = 0;
    int band_y0 = 0, band_y1 = 0;       // requested band (0,0 = whole frame)
    int y0 = 0, y1 = 0, ey0 = 0, ey1 = 0, halo = 0;
    size_t row_stride = 0;
    DevBuf gb_tn, gb_mesh, gb_uv, rgb;
    DevBuf res[3];
    int hist = 2, cur = 0, spare = 1;   // indices into res[]: previous frame's final state / newest state / free work buffer
    bool history_valid = false;

    // row-group completion counters (FineDev in device_common.cuh): [stage][group of 4 band rows], stage 0 = the temporal
    // pass, 1 + p = spatial pass p; fine_count[s] = producer launches of stage s since the counters were zeroed
    DevBuf fine_ctr; int fine_groups = 0; uint32_t fine_count[ROMIS_FINE_STAGES] = {}; int fine_src = -1;

    // stepwise frame state
    bool in_frame = false;
    FrameDev fr{};
    int next_pass = 0;
    int n_launches = 0;
    bool stage0_pushed = false;         // the temporal pass of this frame stored its boundary rows into the neighbours' halos itself

    // peer-mapped halos (one process per GPU; see romis_peer_attach)
    // multi-device context (romis_create with n_devices > 1): one child context per device, each renders a row band of the
    // frame; the parent only holds the children, the band edges and the error text (see "multi-device context" below)
    std::vector<romis_ctx*> kids;
    std::vector<int> g_edges; int g_active = 0, g_W = 0, g_H = 0, g_N = 0, g_halo = -1; bool g_wired = false;
    std::vector<int> g_mis_edges; int g_mis_W = 0, g_mis_H = 0;     // band edges of R-MIS / R-OMIS frames (no minimum height)

    struct Peer {
        bool on = false;
        bool ipc = false;                                       // mapped through cudaIpcOpenMemHandle (another process)
        unsigned char* res[3] = {nullptr, nullptr, nullptr};   // the neighbour's three reservoir buffers, IPC-mapped
        uint32_t* flags = nullptr;                              // the neighbour's flag words, IPC-mapped
        int y0 = 0, y1 = 0, ey0 = 0, ey1 = 0;
        size_t row_stride = 0;
    } peer[2];                          // [0] = band below (smaller y), [1] = band above
    DevBuf flags;                       // my flag words: {ready_from_low, ready_from_high, done_from_low, done_from_high, error}
    uint32_t epoch = 0;                 // stage tokens: a running count of exchange stages, the same on every band
    bool exported = false;

    // R-MIS (romis_render_frame_rmis): neighbour grid and accumulator of the last frame
    DevBuf rmis_pv, rmis_nb, rmis_acc, romis_wsum, romis_chosen, romis_tech, romis_contrib, romis_alpha;
    int rmis_W = 0, rmis_H = 0, rmis_K1 = 0;

    // parity capture
    bool capture = false;
    std::map<int, DevBuf> captured;

    // timing
    bool stage_timing = false;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // lights (see "light table maintenance")
    float4* light_mirror = nullptr; size_t light_cap = 0;          // pinned mirror of the device table, capacity in lights
    cudaEvent_t ev_lights = nullptr; bool light_copy_pending = false;
    std::vector<uint32_t> lights_changed, lights_gone;
    DevBuf lights_arch, light_remap, arch_mark, arch_orig_dev, dirty_dev;
    size_t arch_cap = 0;                                            // archive slots allocated on the device
    std::vector<uint32_t> arch_orig;                                // per slot: the light it was archived from, 0xffffffff = free
    std::vector<uint32_t> arch_free;
    uint8_t* arch_mark_host = nullptr; uint32_t marks_slots = 0; bool marks_pending = false; cudaEvent_t ev_marks = nullptr;
    uint32_t* dirty_stage = nullptr; size_t dirty_stage_cap = 0;
    bool arch_auto = true;
    cudaStream_t copy_stream = nullptr;     // image read-back, overlapped with shading (romis_frame_end)
    cudaEvent_t ev_chunk[8] = {};
    std::vector<cudaEvent_t> ev_stage;  // begin, after primary, after initial, after temporal, after spatial p..., after shade
    struct Mark { int kind; int idx; };
    std::vector<Mark> marks;
    romis_timings last{};
    bool timings_pending = false;
};

#define RCHECK(ctx, call)                                                                         \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                     \
            return ROMIS_ERR_CUDA;                                                                \
        }                                                                                         \
    } while (0)

static int fail(romis_ctx* ctx, int code, const std::string& msg) { ctx->err = msg; return code; }
// multi-device context (defined after the frame functions)
static int group_create(const int* device_ids, int n, romis_ctx** out, std::string& err);
static int group_upload_scene(romis_ctx* g, const romis_mesh_desc* meshes, int n_meshes, const romis_texture* textures, int n_textures);
static int group_upload_lights(romis_ctx* g, const romis_light* lights, int n, int first, int count);
static int group_render_frame(romis_ctx* g, const romis_features* f, const romis_camera* cam, int W, int H, int history_valid, const romis_rng* rng, float* out_rgb);
static int group_timings(romis_ctx* g, romis_timings* out);
static int group_fail(romis_ctx* g, romis_ctx* kid, int rc);
#define ROMIS_GROUP_EACH(g, CALL) do { for (romis_ctx* k : (g)->kids) { int rc__ = (CALL); if (rc__) return group_fail((g), k, rc__); } return ROMIS_OK; } while (0)
#define ROMIS_GROUP_ACTIVE(g, CALL) do { for (int i__ = 0; i__ < std::max(1, (g)->g_active); i__++) { romis_ctx* k = (g)->kids[i__]; int rc__ = (CALL); if (rc__) return group_fail((g), k, rc__); } return ROMIS_OK; } while (0)
#define ROMIS_NOT_ON_GROUP(c, what) if ((c) && !(c)->kids.empty()) return fail((c), ROMIS_ERR_INVALID, what ": per-device call, not available on a multi-device context")
static float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

static ResBuf resbuf(const romis_ctx* c, int i) {
    ResBuf b; b.base = (unsigned char*)c->res[i].p; b.row_stride = c->row_stride; b.W = c->W; b.N = c->N; return b;
}
static GBufDev gbuf(const romis_ctx* c) {
    GBufDev g; g.tn = (float4*)c->gb_tn.p; g.mesh = (uint32_t*)c->gb_mesh.p; g.uv = (float2*)c->gb_uv.p; g.pv = nullptr; return g;
}

// ------------------------------------------------------------------------------------------------
extern "C" int romis_abi_version(void) { return ROMIS_ABI_VERSION; }

extern "C" const char* romis_last_error(const romis_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mutex);
    static thread_local std::string copy;
    copy = g_create_err;
    return copy.c_str();
}

extern "C" int romis_create(const int* device_ids, int n_devices, romis_ctx** out) {
    auto set_err = [](const std::string& s) { std::lock_guard<std::mutex> lk(g_err_mutex); g_create_err = s; };
    if (!out) { set_err("out == NULL"); return ROMIS_ERR_INVALID; }
    *out = nullptr;
    if (n_devices > 1 && device_ids) {           // one caller, several GPUs: a parent context with one row band per device
        std::string err;
        int rc = group_create(device_ids, n_devices, out, err);
        if (rc != ROMIS_OK) set_err(err);
        return rc;
    }
    if (n_devices != 1 && !(n_devices == 0 && device_ids == nullptr)) { set_err("romis_create: bad device list"); return ROMIS_ERR_INVALID; }
    int dev = (n_devices == 1 && device_ids) ? device_ids[0] : 0;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_err(std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
        return ROMIS_ERR_CUDA;
    }
    if (dev < 0 || dev >= count) { set_err("device id out of range"); return ROMIS_ERR_INVALID; }
    e = cudaSetDevice(dev);
    if (e != cudaSuccess) { set_err(std::string("cudaSetDevice: ") + cudaGetErrorString(e)); return ROMIS_ERR_CUDA; }
    romis_ctx* c = new (std::nothrow) romis_ctx();
    if (!c) { set_err("out of host memory"); return ROMIS_ERR_NOMEM; }
    c->device = dev;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev_end);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_lights, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_marks, cudaEventDisableTiming);
    for (int k = 0; k < 8 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&c->ev_chunk[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { set_err(std::string("context setup: ") + cudaGetErrorString(e)); delete c; return ROMIS_ERR_CUDA; }
    *out = c;
    return ROMIS_OK;
}

extern "C" int romis_peer_detach(romis_ctx* c);
extern "C" void romis_destroy(romis_ctx* c) {
    if (!c) return;
    if (!c->kids.empty()) {
        for (romis_ctx* k : c->kids) { cudaSetDevice(k->device); if (k->stream) cudaStreamSynchronize(k->stream); k->peer[0] = romis_ctx::Peer(); k->peer[1] = romis_ctx::Peer(); }
        for (romis_ctx* k : c->kids) romis_destroy(k);
        delete c;
        return;
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (DevBuf* b : {&c->nodes, &c->tri_geom, &c->tri_attr, &c->materials, &c->tex_pixels, &c->tex_desc, &c->lights,
                      &c->gb_tn, &c->gb_mesh, &c->gb_uv, &c->rgb, &c->res[0], &c->res[1], &c->res[2], &c->rmis_pv, &c->rmis_nb, &c->rmis_acc, &c->romis_wsum, &c->romis_chosen, &c->romis_tech, &c->romis_contrib, &c->romis_alpha,
                      &c->lights_arch, &c->light_remap, &c->arch_mark, &c->arch_orig_dev, &c->dirty_dev, &c->fine_ctr}) b->release();
    romis_peer_detach(c);
    c->flags.release();
    for (auto& kv : c->captured) kv.second.release();
    for (cudaEvent_t e : c->ev_stage) cudaEventDestroy(e);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    for (int k = 0; k < 8; k++) if (c->ev_chunk[k]) cudaEventDestroy(c->ev_chunk[k]);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_lights) cudaEventDestroy(c->ev_lights);
    if (c->ev_marks) cudaEventDestroy(c->ev_marks);
    if (c->light_mirror) cudaFreeHost(c->light_mirror);
    if (c->arch_mark_host) cudaFreeHost(c->arch_mark_host);
    if (c->dirty_stage) cudaFreeHost(c->dirty_stage);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int romis_stream(romis_ctx* c, void** s) {
    if (!c || !s) return ROMIS_ERR_INVALID;
    ROMIS_NOT_ON_GROUP(c, "romis_stream");
    *s = (void*)c->stream; return ROMIS_OK;
}

extern "C" int romis_synchronize(romis_ctx* c) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) ROMIS_GROUP_EACH(c, romis_synchronize(k));
    RCHECK(c, cudaSetDevice(c->device));
    RCHECK(c, cudaStreamSynchronize(c->stream));
    return ROMIS_OK;
}

// ------------------------------------------------------------------------------------------------
// scene
// ------------------------------------------------------------------------------------------------
// A bound c such that romis_powf(x, shininess) is +-0 (or NaN) for EVERY |x| <= c: computeShading's specular factor
// pow(cosTheta, shininess) (shading.cpp:26) underflows there, e.g. c = 0.659 for Ns 250.  Found by bisection on the
// underflow edge of the same romis_powf the kernels call, then pulled in by 0.1 %: x^s <= 0.999^s * 2^-150 is far enough
// below the rounding boundary for the binary64 evaluation (relative error ~1e-15) to round to zero too.  The device test
// compares dot(Rraw, V)^2 with c^2 (1 - 1e-4) |Rraw|^2 on the UN-normalised reflection vector; its fp32 roundings and those
// of the exact route (normalise, dot) are each below 1.5e-6 absolute on |cosTheta|, which the 1e-4 relative margin covers
// only while c >= 0.05 -- below that (low exponents) the shortcut is off (returns 0).
extern "C" float romis_specular_cutoff(float shininess) {
    if (!(shininess >= 1.0f) || !(shininess < 3.0e38f)) return 0.0f;
    if (romis_powf(0.05f, shininess) != 0.0f) return 0.0f;
    float lo = 0.05f, hi = 1.0f;                    // pow(lo) == 0, pow(hi) == 1
    for (int i = 0; i < 64; i++) {
        const float mid = 0.5f * (lo + hi);
        if (mid <= lo || mid >= hi) break;
        if (romis_powf(mid, shininess) == 0.0f) lo = mid; else hi = mid;
    }
    return lo * 0.999f;
}

extern "C" int romis_upload_scene(romis_ctx* c, const romis_mesh_desc* meshes, int n_meshes,
                                  const romis_texture* textures, int n_textures) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) return group_upload_scene(c, meshes, n_meshes, textures, n_textures);
    if (n_meshes < 0 || n_textures < 0 || (n_meshes > 0 && !meshes) || (n_textures > 0 && !textures))
        return fail(c, ROMIS_ERR_INVALID, "romis_upload_scene: bad arguments");
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_upload_scene: frame in flight");
    RCHECK(c, cudaSetDevice(c->device));
    size_t ntri = 0;
    for (int m = 0; m < n_meshes; m++) {
        if ((meshes[m].n_triangles && !meshes[m].triangles) || (meshes[m].n_vertices && !meshes[m].vertices))
            return fail(c, ROMIS_ERR_INVALID, "romis_upload_scene: mesh with null arrays");
        for (uint32_t t = 0; t < 3 * meshes[m].n_triangles; t++)
            if (meshes[m].triangles[t] >= meshes[m].n_vertices) return fail(c, ROMIS_ERR_INVALID, "romis_upload_scene: vertex index out of range");
        ntri += meshes[m].n_triangles;
    }
    if (ntri > 0x7fffffffu / 16) return fail(c, ROMIS_ERR_INVALID, "romis_upload_scene: too many triangles");
    // global triangle order = mesh order, then the mesh's triangle order (geomID = mesh index, embree_interface.cpp:46-47)
    std::vector<float> verts(9 * ntri + 1);
    std::vector<float4> attr(4 * ntri + 1);
    size_t g = 0;
    for (int m = 0; m < n_meshes; m++) {
        for (uint32_t t = 0; t < meshes[m].n_triangles; t++, g++) {
            const romis_vertex* v[3];
            for (int k = 0; k < 3; k++) { v[k] = &meshes[m].vertices[meshes[m].triangles[3 * t + k]]; std::memcpy(&verts[9 * g + 3 * k], v[k]->position, 12); }
            attr[4 * g + 0] = make_float4(v[0]->normal[0], v[0]->normal[1], v[0]->normal[2], v[0]->texcoord[0]);
            attr[4 * g + 1] = make_float4(v[1]->normal[0], v[1]->normal[1], v[1]->normal[2], v[0]->texcoord[1]);
            attr[4 * g + 2] = make_float4(v[2]->normal[0], v[2]->normal[1], v[2]->normal[2], v[1]->texcoord[0]);
            attr[4 * g + 3] = make_float4(v[1]->texcoord[1], v[2]->texcoord[0], v[2]->texcoord[1], u2f((uint32_t)m));
        }
    }
    Bvh bvh = build_bvh(verts.data(), (int)ntri);
    if (bvh.max_depth >= ROMIS_STACK) return fail(c, ROMIS_ERR_INVALID, "romis_upload_scene: BVH deeper than the traversal stack");

    std::vector<float4> mats(3 * (size_t)(n_meshes + 1));
    bool any_tex = false;
    for (int m = 0; m < n_meshes; m++) {
        const romis_material& mt = meshes[m].material;
        int tex = (mt.kd_texture >= 0 && mt.kd_texture < n_textures) ? mt.kd_texture : -1;
        any_tex |= tex >= 0;
        mats[3 * m] = make_float4(mt.kd[0], mt.kd[1], mt.kd[2], mt.shininess);
        mats[3 * m + 1] = make_float4(mt.ks[0], mt.ks[1], mt.ks[2], u2f((uint32_t)tex));
        // ks = 0: the specular term is +-0 (or NaN, zeroed by shading.cpp:28) for every sample; otherwise it is outside the lobe
        float cut2 = 0.0f;
        if (mt.ks[0] == 0.0f && mt.ks[1] == 0.0f && mt.ks[2] == 0.0f) cut2 = INFINITY;
        else { const float cut = romis_specular_cutoff(mt.shininess); cut2 = cut * cut * (1.0f - 1e-4f); }
        mats[3 * m + 2] = make_float4(cut2, 0, 0, 0);
    }
    // miss pixels carry a value-initialised Material: kd = ks = 0, shininess = 1 (mesh.h:22-34)
    mats[3 * n_meshes] = make_float4(0, 0, 0, 1.0f);
    mats[3 * n_meshes + 1] = make_float4(0, 0, 0, u2f(0xffffffffu));
    mats[3 * n_meshes + 2] = make_float4(0, 0, 0, 0);

    std::vector<float> texpx; std::vector<int4> texdesc((size_t)std::max(1, n_textures));
    for (int t = 0; t < n_textures; t++) {
        if (!textures[t].pixels || textures[t].width <= 0 || textures[t].height <= 0) return fail(c, ROMIS_ERR_INVALID, "romis_upload_scene: bad texture");
        size_t n = (size_t)textures[t].width * textures[t].height * 3;
        texdesc[t] = make_int4((int)texpx.size(), textures[t].width, textures[t].height, 0);
        texpx.insert(texpx.end(), textures[t].pixels, textures[t].pixels + n);
    }

    RCHECK(c, cudaStreamSynchronize(c->stream));
    RCHECK(c, c->nodes.ensure(bvh.nodes.size() * sizeof(BvhNode)));
    RCHECK(c, c->tri_geom.ensure(std::max<size_t>(1, bvh.tris.size()) * sizeof(TriGeom)));
    RCHECK(c, c->tri_attr.ensure(attr.size() * sizeof(float4)));
    RCHECK(c, c->materials.ensure(mats.size() * sizeof(float4)));
    RCHECK(c, c->tex_pixels.ensure(std::max<size_t>(1, texpx.size()) * sizeof(float)));
    RCHECK(c, c->tex_desc.ensure(texdesc.size() * sizeof(int4)));
    RCHECK(c, cudaMemcpy(c->nodes.p, bvh.nodes.data(), bvh.nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));
    if (!bvh.tris.empty()) RCHECK(c, cudaMemcpy(c->tri_geom.p, bvh.tris.data(), bvh.tris.size() * sizeof(TriGeom), cudaMemcpyHostToDevice));
    RCHECK(c, cudaMemcpy(c->tri_attr.p, attr.data(), attr.size() * sizeof(float4), cudaMemcpyHostToDevice));
    RCHECK(c, cudaMemcpy(c->materials.p, mats.data(), mats.size() * sizeof(float4), cudaMemcpyHostToDevice));
    if (!texpx.empty()) RCHECK(c, cudaMemcpy(c->tex_pixels.p, texpx.data(), texpx.size() * sizeof(float), cudaMemcpyHostToDevice));
    RCHECK(c, cudaMemcpy(c->tex_desc.p, texdesc.data(), texdesc.size() * sizeof(int4), cudaMemcpyHostToDevice));

    c->sc.nodes = (const BvhNode*)c->nodes.p;
    c->sc.tri_geom = (const float4*)c->tri_geom.p;
    c->sc.tri_attr = (const float4*)c->tri_attr.p;
    c->sc.materials = (const float4*)c->materials.p;
    c->sc.tex_pixels = (const float*)c->tex_pixels.p;
    c->sc.tex_desc = (const int4*)c->tex_desc.p;
    c->sc.n_meshes = n_meshes;
    c->sc.has_textures = any_tex ? 1 : 0;
    c->n_tris = (int)ntri;
    c->has_scene = true;
    c->history_valid = false;
    c->W = c->H = c->N = 0;             // G-buffer layout depends on has_textures: force re-allocation
    return ROMIS_OK;
}

// Light record: six float4, ordered so that a point light needs the first two and a segment light the first four.
//   {type, p0} {c0, c3.x} {e1, c3.y} {c1, c3.z} {e2, 0} {c2, 0}
static void pack_light(const romis_light& l, float4* r) {
    r[0] = make_float4(u2f((uint32_t)l.type), l.p0[0], l.p0[1], l.p0[2]);
    r[1] = make_float4(l.c0[0], l.c0[1], l.c0[2], l.c3[0]);
    r[2] = make_float4(l.e1[0], l.e1[1], l.e1[2], l.c3[1]);
    r[3] = make_float4(l.c1[0], l.c1[1], l.c1[2], l.c3[2]);
    r[4] = make_float4(l.e2[0], l.e2[1], l.e2[2], 0.0f);
    r[5] = make_float4(l.c2[0], l.c2[1], l.c2[2], 0.0f);
}

// ---- light table maintenance ----
// The packed table lives three times: on the device (sc.lights), as a pinned host mirror (light_mirror: what the device holds
// once the copies in flight have landed; also the staging area of those copies) and in the caller's Scene.  The reference reads
// scene.lights fresh every frame and the UI edits them without notification (light.cpp:46-66, ui.cpp:172-261), so the drop-in
// hands the table over every frame and dirty tracking lives here.
//
// What an edit means for the temporal history: the reference's reservoirs hold LightSample{position, color} by value
// (reservoir.h:18-26) and temporalReuse streams the predecessor's samples as stored (render_utils.cpp:154-170), so a sample
// drawn from a light keeps that light's OLD position / colour however the light is edited or removed afterwards, for as long
// as it survives resampling.  The records here hold (light, u, v); to stay bit-identical, the old record of every edited or
// removed light is copied (device to device) into an archive slot and the history records that hold the light are re-pointed to
// the slot (launch_light_archive), before the new record overwrites the old.  The same pass marks which slots are still held;
// unheld slots are recycled at the next edit (a slot no history record holds can never be held again).
static int grow_light_tables(romis_ctx* c, size_t need) {
    if (need <= c->light_cap) return ROMIS_OK;
    size_t cap = std::max<size_t>(std::max<size_t>(need, 64), c->light_cap + c->light_cap / 2);
    float4* m = nullptr;
    RCHECK(c, cudaStreamSynchronize(c->stream));
    RCHECK(c, cudaMallocHost((void**)&m, cap * 6 * sizeof(float4)));
    if (c->light_mirror) { std::memcpy(m, c->light_mirror, c->light_cap * 6 * sizeof(float4)); cudaFreeHost(c->light_mirror); }
    c->light_mirror = m;
    RCHECK(c, c->lights.ensure(cap * 6 * sizeof(float4)));
    RCHECK(c, c->light_remap.ensure(cap * sizeof(uint32_t)));
    RCHECK(c, cudaMemsetAsync(c->light_remap.p, 0xff, cap * sizeof(uint32_t), c->stream));
    c->light_cap = cap;
    c->sc.lights = (const float4*)c->lights.p;
    return ROMIS_OK;
}

static void archive_free_all(romis_ctx* c) {
    c->arch_free.clear();
    for (size_t s = c->arch_orig.size(); s-- > 0;) { c->arch_orig[s] = 0xffffffffu; c->arch_free.push_back((uint32_t)s); }
    c->marks_pending = false;
}

// recycle the slots the last edit's mark pass found unheld (keep == nullptr: this context's own marks)
static int archive_harvest(romis_ctx* c, const uint8_t* keep, uint32_t n_slots) {
    if (!keep) {
        if (!c->marks_pending) return ROMIS_OK;
        RCHECK(c, cudaEventSynchronize(c->ev_marks));
        keep = c->arch_mark_host; n_slots = c->marks_slots;
    }
    for (uint32_t s = 0; s < n_slots && s < c->arch_orig.size(); s++)
        if (c->arch_orig[s] != 0xffffffffu && !keep[s]) { c->arch_orig[s] = 0xffffffffu; c->arch_free.push_back(s); }
    c->marks_pending = false;
    return ROMIS_OK;
}

static int archive_lights(romis_ctx* c, const std::vector<uint32_t>& gone) {
    if (c->arch_auto) { int rc = archive_harvest(c, nullptr, 0); if (rc) return rc; }
    const size_t n = gone.size();
    // pinned staging of {light, slot} pairs
    if (c->dirty_stage_cap < 2 * n) {
        RCHECK(c, cudaStreamSynchronize(c->stream));
        if (c->dirty_stage) cudaFreeHost(c->dirty_stage);
        c->dirty_stage = nullptr; c->dirty_stage_cap = 0;
        RCHECK(c, cudaMallocHost((void**)&c->dirty_stage, 2 * n * sizeof(uint32_t)));
        c->dirty_stage_cap = 2 * n;
        RCHECK(c, c->dirty_dev.ensure(2 * n * sizeof(uint32_t)));
    }
    for (size_t k = 0; k < n; k++) {
        uint32_t slot;
        if (!c->arch_free.empty()) { slot = c->arch_free.back(); c->arch_free.pop_back(); }
        else { slot = (uint32_t)c->arch_orig.size(); c->arch_orig.push_back(0xffffffffu); }
        c->arch_orig[slot] = gone[k];
        c->dirty_stage[k] = gone[k]; c->dirty_stage[n + k] = slot;
    }
    const size_t slots = c->arch_orig.size();
    if (slots >= 0x7fffffffu) return fail(c, ROMIS_ERR_NOMEM, "light archive full");
    if (slots > c->arch_cap) {                          // grow the device archive, keeping what it holds
        const size_t cap = std::max<size_t>(std::max<size_t>(slots, 256), 2 * c->arch_cap);
        DevBuf bigger, orig, mark;
        RCHECK(c, bigger.ensure(cap * 6 * sizeof(float4)));
        RCHECK(c, orig.ensure(cap * sizeof(uint32_t)));
        RCHECK(c, mark.ensure(cap));
        if (c->arch_cap) RCHECK(c, cudaMemcpyAsync(bigger.p, c->lights_arch.p, c->arch_cap * 6 * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        RCHECK(c, cudaStreamSynchronize(c->stream));
        c->lights_arch.release(); c->arch_orig_dev.release(); c->arch_mark.release();
        c->lights_arch = bigger; c->arch_orig_dev = orig; c->arch_mark = mark;
        if (c->arch_mark_host) cudaFreeHost(c->arch_mark_host);
        c->arch_mark_host = nullptr;
        RCHECK(c, cudaMallocHost((void**)&c->arch_mark_host, cap));
        c->arch_cap = cap;
        c->sc.lights_arch = (const float4*)c->lights_arch.p;
    }
    RCHECK(c, cudaMemcpyAsync(c->dirty_dev.p, c->dirty_stage, 2 * n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    RCHECK(c, cudaMemcpyAsync(c->arch_orig_dev.p, c->arch_orig.data(), slots * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    RCHECK(c, cudaMemsetAsync(c->arch_mark.p, 0, slots, c->stream));
    ResBuf h; h.base = (unsigned char*)c->res[c->hist].p; h.row_stride = c->row_stride; h.W = c->W; h.N = c->N;
    launch_light_archive(c->stream, (const float4*)c->lights.p, (float4*)c->lights_arch.p, (uint32_t*)c->light_remap.p,
                         (const uint32_t*)c->dirty_dev.p, (const uint32_t*)c->dirty_dev.p + n, (int)n,
                         h, c->y0 - c->ey0, c->y1 - c->y0, c->N, (uint32_t)c->light_cap, (uint8_t*)c->arch_mark.p, (uint32_t)slots);
    RCHECK(c, cudaGetLastError());
    RCHECK(c, cudaMemcpyAsync(c->arch_mark_host, c->arch_mark.p, slots, cudaMemcpyDeviceToHost, c->stream));
    RCHECK(c, cudaEventRecord(c->ev_marks, c->stream));
    c->marks_pending = true; c->marks_slots = (uint32_t)slots;
    // the pageable arch_orig copy above and the pinned pair list must not be rewritten before they are read
    RCHECK(c, cudaEventRecord(c->ev_lights, c->stream));
    c->light_copy_pending = true;
    return ROMIS_OK;
}

// lights[first .. first + count) may differ from what the device holds; everything else is known to be unchanged
static int upload_lights_impl(romis_ctx* c, const romis_light* lights, int n, int first, int count) {
    if (!c) return ROMIS_ERR_INVALID;
    if (n < 0 || (n > 0 && !lights) || n >= 0x7fffffff) return fail(c, ROMIS_ERR_INVALID, "romis_upload_lights: bad arguments");
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_upload_lights: frame in flight");
    const int old_n = c->sc.n_lights;
    if (n != old_n || first < 0 || count < 0 || first > n || count > n - first) { first = 0; count = n; }
    for (int i = first; i < first + count; i++)
        if (lights[i].type > ROMIS_LIGHT_PARALLELOGRAM) return fail(c, ROMIS_ERR_INVALID, "romis_upload_lights: unknown light type");
    RCHECK(c, cudaSetDevice(c->device));
    if (c->light_copy_pending) { RCHECK(c, cudaEventSynchronize(c->ev_lights)); c->light_copy_pending = false; }   // copies out of the mirror: long done

    // 1. which lights changed (bitwise, on the packed record)
    std::vector<uint32_t>& changed = c->lights_changed; changed.clear();
    std::vector<uint32_t>& gone = c->lights_gone; gone.clear();      // edited or removed: their old records matter to the history
    const int common = std::min(n, old_n);
    float4 tmp[6];
    for (int i = first; i < std::min(first + count, common); i++) {
        pack_light(lights[i], tmp);
        if (std::memcmp(c->light_mirror + 6 * (size_t)i, tmp, sizeof tmp) != 0) { changed.push_back((uint32_t)i); gone.push_back((uint32_t)i); }
    }
    for (int i = n; i < old_n; i++) gone.push_back((uint32_t)i);
    if (changed.empty() && n == old_n && c->sc.lights) return ROMIS_OK;

    // 2. the history keeps the old records (before anything overwrites them on the device)
    if (!c->history_valid || !c->W) archive_free_all(c);
    else if (!gone.empty()) { int rc = archive_lights(c, gone); if (rc) return rc; }

    // 3. new records: into the mirror, then the changed runs (or the whole table after a re-allocation) to the device
    const size_t cap_before = c->light_cap;
    int rc = grow_light_tables(c, (size_t)std::max(n, 1));
    if (rc) return rc;
    const bool all = c->light_cap != cap_before;
    for (uint32_t i : changed) pack_light(lights[i], c->light_mirror + 6 * (size_t)i);
    for (int i = old_n; i < n; i++) { pack_light(lights[i], c->light_mirror + 6 * (size_t)i); changed.push_back((uint32_t)i); }
    auto send = [&](size_t a, size_t b) {       // lights [a, b)
        return cudaMemcpyAsync((float4*)c->lights.p + 6 * a, c->light_mirror + 6 * a, (b - a) * 6 * sizeof(float4), cudaMemcpyHostToDevice, c->stream);
    };
    if (all) { if (n > 0) RCHECK(c, send(0, (size_t)n)); }
    else if (!changed.empty()) {
        size_t runs = 1;
        for (size_t k = 1; k < changed.size(); k++) runs += changed[k] != changed[k - 1] + 1;
        if (runs > 16) RCHECK(c, send(changed.front(), (size_t)changed.back() + 1));
        else for (size_t k = 0; k < changed.size();) {
            size_t e = k + 1;
            while (e < changed.size() && changed[e] == changed[e - 1] + 1) e++;
            RCHECK(c, send(changed[k], (size_t)changed[e - 1] + 1));
            k = e;
        }
    }
    RCHECK(c, cudaEventRecord(c->ev_lights, c->stream));
    c->light_copy_pending = true;
    c->sc.lights = (const float4*)c->lights.p;
    c->sc.lights_arch = (const float4*)c->lights_arch.p;
    c->sc.n_lights = n;
    return ROMIS_OK;
}

extern "C" int romis_upload_lights(romis_ctx* c, const romis_light* lights, int n) {
    if (c && !c->kids.empty()) return group_upload_lights(c, lights, n, 0, n);
    return upload_lights_impl(c, lights, n, 0, n);
}

extern "C" int romis_upload_lights_range(romis_ctx* c, const romis_light* lights, int n, int first_dirty, int n_dirty) {
    if (c && !c->kids.empty()) return group_upload_lights(c, lights, n, first_dirty, n_dirty);
    return upload_lights_impl(c, lights, n, first_dirty, n_dirty);
}

extern "C" int romis_set_light_archive_auto(romis_ctx* c, int on) {
    if (!c) return ROMIS_ERR_INVALID;
    ROMIS_NOT_ON_GROUP(c, "romis_set_light_archive_auto");
    c->arch_auto = on != 0; return ROMIS_OK;
}

extern "C" int romis_light_archive_marks(romis_ctx* c, uint8_t* marks, int capacity, int* n_slots) {
    if (!c || !n_slots || capacity < 0 || (capacity > 0 && !marks)) return ROMIS_ERR_INVALID;
    ROMIS_NOT_ON_GROUP(c, "romis_light_archive_marks");
    *n_slots = 0;
    if (!c->marks_pending) return ROMIS_OK;
    RCHECK(c, cudaSetDevice(c->device));
    RCHECK(c, cudaEventSynchronize(c->ev_marks));
    if ((uint32_t)capacity < c->marks_slots) return fail(c, ROMIS_ERR_INVALID, "romis_light_archive_marks: buffer too small");
    std::memcpy(marks, c->arch_mark_host, c->marks_slots);
    *n_slots = (int)c->marks_slots;
    return ROMIS_OK;
}

extern "C" int romis_light_archive_release(romis_ctx* c, const uint8_t* keep, int n_slots) {
    if (!c || n_slots < 0 || (n_slots > 0 && !keep)) return ROMIS_ERR_INVALID;
    ROMIS_NOT_ON_GROUP(c, "romis_light_archive_release");
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_light_archive_release: frame in flight");
    return archive_harvest(c, keep, (uint32_t)n_slots);
}

extern "C" int romis_light_archive_size(romis_ctx* c, int* n_slots, int* n_held) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) c = c->kids[0];
    if (n_slots) *n_slots = (int)c->arch_orig.size();
    if (n_held) *n_held = (int)(c->arch_orig.size() - c->arch_free.size());
    return ROMIS_OK;
}

// ------------------------------------------------------------------------------------------------
// frame
// ------------------------------------------------------------------------------------------------
extern "C" int romis_set_band(romis_ctx* c, int y0, int y1) {
    if (!c) return ROMIS_ERR_INVALID;
    ROMIS_NOT_ON_GROUP(c, "romis_set_band");
    if (y0 < 0 || y1 < y0) return fail(c, ROMIS_ERR_INVALID, "romis_set_band: need 0 <= y0 <= y1");
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_set_band: frame in flight");
    if (y0 != c->band_y0 || y1 != c->band_y1) { c->band_y0 = y0; c->band_y1 = y1; c->W = c->H = c->N = 0; c->history_valid = false; }
    return ROMIS_OK;
}

extern "C" int romis_reset_history(romis_ctx* c) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) ROMIS_GROUP_EACH(c, romis_reset_history(k));
    c->history_valid = false; return ROMIS_OK;
}
extern "C" int romis_set_capture(romis_ctx* c, int on) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) ROMIS_GROUP_EACH(c, romis_set_capture(k, on));
    c->capture = on != 0; return ROMIS_OK;
}
extern "C" int romis_set_stage_timing(romis_ctx* c, int on) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) ROMIS_GROUP_EACH(c, romis_set_stage_timing(k, on));
    c->stage_timing = on != 0; return ROMIS_OK;
}

static int validate(romis_ctx* c, const romis_features* f, const romis_camera* cam, int W, int H, const romis_rng* rng) {
    if (!f || !cam || !rng) return fail(c, ROMIS_ERR_INVALID, "null features / camera / rng");
    if (W < 1 || H < 1 || W > 65535 || H > 65535) return fail(c, ROMIS_ERR_INVALID, "resolution must be 1..65535 in each dimension");
    if (f->numSamplesInReservoir < 1 || f->numSamplesInReservoir > 32) return fail(c, ROMIS_ERR_INVALID, "numSamplesInReservoir must be 1..32 (ui.cpp:305)");
    if (f->initialLightSamples < 1) return fail(c, ROMIS_ERR_INVALID, "initialLightSamples must be >= 1");
    if (f->numNeighboursToSample > ROMIS_MAX_K) return fail(c, ROMIS_ERR_INVALID, "numNeighboursToSample must be <= 32");
    if (f->spatialResampleRadius > 4096) return fail(c, ROMIS_ERR_INVALID, "spatialResampleRadius must be <= 4096");
    // random-stream stages ROMIS_STAGE_SPATIAL0 + pass must stay below ROMIS_STAGE_RMIS_NEIGH (include/romis_rng.h); the UI allows 1..5
    if (f->spatialReuse && f->spatialResamplingPasses > 61) return fail(c, ROMIS_ERR_INVALID, "spatialResamplingPasses must be <= 61");
    if (!c->has_scene) return fail(c, ROMIS_ERR_STATE, "no scene uploaded");
    return ROMIS_OK;
}

static cudaError_t mark(romis_ctx* c, int kind, int idx) {
    if (!c->stage_timing) return cudaSuccess;
    size_t i = c->marks.size();
    if (i >= c->ev_stage.size()) { cudaEvent_t e; cudaError_t r = cudaEventCreate(&e); if (r != cudaSuccess) return r; c->ev_stage.push_back(e); }
    c->marks.push_back({kind, idx});
    return cudaEventRecord(c->ev_stage[i], c->stream);
}

static int capture_stage(romis_ctx* c, int pass_id, int buf) {
    if (!c->capture) return ROMIS_OK;
    DevBuf& d = c->captured[pass_id];
    RCHECK(c, d.ensure(c->res[buf].bytes));
    RCHECK(c, cudaMemcpyAsync(d.p, c->res[buf].p, c->res[buf].bytes, cudaMemcpyDeviceToDevice, c->stream));
    return ROMIS_OK;
}

// Block shapes of the pass kernels (launch bounds are for 256 threads).  The kernels that gather a +-r window (spatial pass, R-MIS /
// R-OMIS) run 32x8 blocks, the per-pixel streaming ones (primary, initial, temporal, shade) 32x4: measured on B200 at C2, initial
// 0.995 -> 0.976 ms and shade 0.274 -> 0.266 ms with the smaller block (finer refill of an SM whose warps run for > 100 us), the
// spatial pass 0.381 -> 0.389 ms (less overlap between the windows of a block).  ROMIS_BLOCK_Y / ROMIS_BLOCK_YS=<n> in the
// environment override them for tuning runs.
static dim3 make_block(const char* env, int by) {
    if (const char* e = std::getenv(env)) { int v = std::atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8) by = v; }
    return dim3(32, by);
}
static const dim3 kBlock = make_block("ROMIS_BLOCK_Y", 8);
static const dim3 kBlockS = make_block("ROMIS_BLOCK_YS", 4);
static dim3 grid_for(int W, int rows, const dim3& b = kBlock) { return dim3((W + b.x - 1) / b.x, (rows + b.y - 1) / b.y); }

// Row-group completion counters between the passes of a frame (FineDev): `produce` >= 0: this launch counts its finished blocks
// into stage `produce`; `consume` >= 0: it waits per row group for stage `consume` (reading `reach` rows beyond its own) instead
// of for the whole previous kernel.  Needs blocks of 4 or 8 rows (a block then covers whole groups); ROMIS_FINE=0 turns it off.
// The counting costs a barrier in the middle of the consumer and a barrier + fence + atomic at the end of every producer block
// (measured on B200, C2: temporal +12 %, spatial pass +20 % on the full 1080p frame), the gain is the drained tail of each pass --
// a fixed ~25 us per frame.  It pays on thin bands only (135 rows: 0.573 -> 0.548 ms; full frame: 2.50 -> 2.73 ms), so by default
// it is on when the spatial grid is fewer than 4 waves of resident blocks.  ROMIS_FINE=0 / 1 forces it off / on.
// (Re-measured with the one-directional fences: 135 rows 0.574 -> 0.539 ms, 270 rows 0.943 -> 0.940 ms, 540 rows 1.588 -> 1.668 ms,
// full frame 2.468 -> 2.650 ms: the threshold stands.)
static bool fine_enabled(const romis_ctx* c) {
    static const int mode = [] { const char* e = std::getenv("ROMIS_FINE"); return e ? (std::atoi(e) != 0 ? 1 : 0) : -1; }();
    if (mode == 0 || kBlock.y % 4 != 0 || kBlockS.y % 4 != 0 || !c->W) return false;
    if (mode == 1) return true;
    const long long blocks = (long long)((c->W + 31) / 32) * ((c->y1 - c->y0 + (int)kBlock.y - 1) / (int)kBlock.y);
    return blocks < 4LL * c->n_sms * ROMIS_MINB_SPATIAL;
}
static FineDev fine_for(romis_ctx* c, int consume, int produce, int reach) {
    FineDev fd; std::memset(&fd, 0, sizeof fd);
    fd.y0 = c->y0; fd.y1 = c->y1; fd.reach = reach;
    unsigned int* base = (unsigned int*)c->fine_ctr.p;
    fd.err = (uint32_t*)(base + (size_t)ROMIS_FINE_STAGES * c->fine_groups);
    const unsigned int per_group = (unsigned int)((c->W + 31) / 32);        // blocks are 32 pixels wide
    if (!fine_enabled(c)) return fd;
    if (consume >= 0) { fd.wait_ctr = base + (size_t)consume * c->fine_groups; fd.wait_target = per_group * c->fine_count[consume]; }
    if (produce >= 0) { fd.sig_ctr = base + (size_t)produce * c->fine_groups; c->fine_count[produce]++; }
    return fd;
}

// (Re)allocates the per-frame buffers for (W, H, N, band, halo).  A change of resolution, band or N drops the temporal
// history (the reference would read out of bounds).
static int ensure_frame_buffers(romis_ctx* c, const romis_features* f, int W, int H, int min_halo = 0) {
    const int N = (int)f->numSamplesInReservoir;
    int y0 = 0, y1 = H;
    if (c->band_y1 > c->band_y0) { y0 = c->band_y0; y1 = c->band_y1; if (y1 > H) return fail(c, ROMIS_ERR_INVALID, "band exceeds the image height"); }
    const int halo = std::max(min_halo, f->spatialReuse ? (int)f->spatialResampleRadius : 0);
    if (W != c->W || H != c->H || N != c->N || halo > c->halo || y0 != c->y0 || y1 != c->y1) {
        if (c->peer[0].on || c->peer[1].on || c->exported)
            return fail(c, ROMIS_ERR_STATE, "frame geometry changed after romis_peer_export: detach, prepare, export and attach again");
        RCHECK(c, cudaStreamSynchronize(c->stream));
        c->W = W; c->H = H; c->N = N; c->halo = halo; c->y0 = y0; c->y1 = y1;
        c->ey0 = std::max(0, y0 - halo); c->ey1 = std::min(H, y1 + halo);
        const size_t rows = (size_t)(c->ey1 - c->ey0), px = rows * W;
        c->row_stride = (((size_t)W * ROMIS_RES_BYTES * N) + 15) & ~(size_t)15;
        RCHECK(c, c->gb_tn.ensure(px * sizeof(float4)));
        RCHECK(c, c->gb_mesh.ensure(px * sizeof(uint32_t)));
        RCHECK(c, c->gb_uv.ensure(c->sc.has_textures ? px * sizeof(float2) : 16));
        RCHECK(c, c->rgb.ensure((size_t)W * H * 3 * sizeof(float)));
        for (int i = 0; i < 3; i++) {
            RCHECK(c, c->res[i].ensure(rows * c->row_stride));
            RCHECK(c, cudaMemsetAsync(c->res[i].p, 0, c->res[i].bytes, c->stream));
        }
        RCHECK(c, cudaMemsetAsync(c->rgb.p, 0, c->rgb.bytes, c->stream));
        c->fine_groups = (y1 - y0 + 3) / 4;
        RCHECK(c, c->fine_ctr.ensure(((size_t)ROMIS_FINE_STAGES * c->fine_groups + 1) * sizeof(unsigned int)));    // + the error word
        RCHECK(c, cudaMemsetAsync(c->fine_ctr.p, 0, c->fine_ctr.bytes, c->stream));
        std::memset(c->fine_count, 0, sizeof c->fine_count);
        c->history_valid = false;
        for (auto& kv : c->captured) kv.second.release();
        c->captured.clear();
    }
    return ROMIS_OK;
}

static HaloDev halo_for_stage0(romis_ctx* c, int out, const dim3& grid, const dim3& block);

extern "C" int romis_frame_begin(romis_ctx* c, const romis_features* f, const romis_camera* cam, int W, int H,
                                 int history_valid, const romis_rng* rng) {
    if (!c) return ROMIS_ERR_INVALID;
    ROMIS_NOT_ON_GROUP(c, "romis_frame_begin");
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_frame_begin: previous frame not ended");
    int rc = validate(c, f, cam, W, H, rng);
    if (rc) return rc;
    RCHECK(c, cudaSetDevice(c->device));
    if ((rc = ensure_frame_buffers(c, f, W, H))) return rc;
    if (!history_valid) c->history_valid = false;

    FrameDev& fr = c->fr;
    fr.cam.origin.x = cam->origin[0]; fr.cam.origin.y = cam->origin[1]; fr.cam.origin.z = cam->origin[2];
    fr.cam.qw = cam->quat[0]; fr.cam.qx = cam->quat[1]; fr.cam.qy = cam->quat[2]; fr.cam.qz = cam->quat[3];
    fr.cam.half_w = cam->half_width; fr.cam.half_h = cam->half_height;
    fr.f = *f; fr.seed = rng->seed; fr.frame = rng->frame;
    fr.W = W; fr.H = H; fr.y0 = c->y0; fr.y1 = c->y1; fr.ey0 = c->ey0; fr.ey1 = c->ey1;
    fr.initial_stage = ROMIS_STAGE_INITIAL;
    // halo rows are only meaningful up to this frame's radius
    const int r_now = f->spatialReuse ? (int)f->spatialResampleRadius : 0;
    const int pey0 = std::max(0, c->y0 - r_now), pey1 = std::min(H, c->y1 + r_now);

    c->marks.clear(); c->n_launches = 0; c->timings_pending = true;
    std::memset(&c->last, 0, sizeof c->last);
    RCHECK(c, cudaEventRecord(c->ev_begin, c->stream));
    RCHECK(c, mark(c, 0, 0));

    // work buffers: the two that are not the history
    const int w0 = (c->hist + 1) % 3;
    c->spare = (c->hist + 2) % 3;
    const dim3 gOwn = grid_for(W, c->y1 - c->y0, kBlockS);

    // 1. primary rays for band + halo rows (the halo G-buffer is re-traced locally instead of exchanged)
    launch_primary(c->stream, grid_for(W, pey1 - pey0, kBlockS), kBlockS, c->sc, fr, gbuf(c), pey0, pey1);
    c->n_launches++;
    RCHECK(c, cudaGetLastError());
    RCHECK(c, mark(c, 1, 0));

    c->stage0_pushed = false;
    c->fine_src = -1;           // which stage's counters the next pass may wait on (-1: wait for the whole previous kernel)
    // 2. initial RIS (+ visibility reuse)
    const bool with_temporal = f->temporalReuse && c->history_valid;
    launch_initial(c->stream, gOwn, kBlockS, c->N, c->sc, c->fr, gbuf(c), resbuf(c, w0));
    c->n_launches++;
    RCHECK(c, cudaGetLastError());
    RCHECK(c, mark(c, 2, 0));
    if ((rc = capture_stage(c, ROMIS_PASS_INITIAL, w0))) return rc;

    // 3. temporal reuse (in place on w0; reads the history)
    if (with_temporal) {
        HaloDev hd; std::memset(&hd, 0, sizeof hd);
        if ((c->peer[0].on || c->peer[1].on) && f->spatialReuse && f->spatialResamplingPasses > 0) { hd = halo_for_stage0(c, w0, gOwn, kBlockS); c->stage0_pushed = true; }
        launch_temporal(c->stream, gOwn, kBlockS, c->N, c->sc, c->fr, gbuf(c), resbuf(c, w0), resbuf(c, c->hist), resbuf(c, w0), fine_for(c, -1, 0, 0), hd);
        c->fine_src = fine_enabled(c) ? 0 : -1;
        c->n_launches++;
        RCHECK(c, cudaGetLastError());
        RCHECK(c, mark(c, 3, 0));
        if ((rc = capture_stage(c, ROMIS_PASS_TEMPORAL, w0))) return rc;
    }
    c->cur = w0;
    c->next_pass = 0;
    c->in_frame = true;
    return ROMIS_OK;
}

// ------------------------------------------------------------------------------------------------
// peer-mapped halos ("fused halo exchange"): the boundary rows a spatial pass writes are stored twice by the pass itself --
// into this band's buffer and, over NVLink, into the neighbouring bands' halo rows of the same buffer (IPC- or peer-mapped
// device memory) -- and the passes are ordered by stage tokens in device memory: no NCCL call, no host synchronisation, no
// separate copy kernel except for the rows of stage 0 (the temporal pass's output).
//   stage    0 = the push of the first pass's input, k = spatial pass k-1; a running count over the frames (`epoch`)
//   token    band A publishes token(s) into both neighbours' flag words once all its row groups next to that edge have
//            finished stage s
//   RAW      the row groups of stage s next to an edge wait for the neighbour's token(s-1): its rows are in my halo
//   WAR      stage s writes into the neighbour's halo of buffer B_s = B_(s-2); the neighbour last read that halo in ITS
//            stage s-1, whose edge row groups had finished when it published token(s-1) -- the same wait.  Stage 0 of a frame
//            waits for the last token of the previous frame for the same reason.
// Only row groups within `radius` rows of an edge ever wait; they run first, the interior of the band follows without a wait.
// ------------------------------------------------------------------------------------------------
struct PeerBlob {
    uint32_t magic;
    int32_t W, H, N, y0, y1, ey0, ey1;
    uint64_t row_stride;
    cudaIpcMemHandle_t res[3];
    cudaIpcMemHandle_t flags;
};
static_assert(sizeof(PeerBlob) <= ROMIS_PEER_BLOB_BYTES, "peer blob fits the ABI buffer");

// Stage 0 of a frame's exchange: the rows the first spatial pass needs come from the temporal (or initial) pass, which knows
// nothing of neighbours, so ONE small copy kernel pushes them; it waits only for the neighbours' last token of the previous
// frame (their halo rows of this buffer are free) and publishes this stage's token -- it does not wait for the neighbours'
// rows, the boundary blocks of the pass do.
static int peer_push_stage0(romis_ctx* c, int in) {
    const int r = (int)c->fr.f.spatialResampleRadius;
    uint32_t* my = (uint32_t*)c->flags.p;      // {token_from_low, token_from_high, -, -, error, ticket, edge counter low, edge counter high}
    const uint32_t wait = c->epoch, token = ++c->epoch;
    const unsigned char* src = (const unsigned char*)c->res[in].p;
    const romis_ctx::Peer& lo = c->peer[0]; const romis_ctx::Peer& hi = c->peer[1];
    const size_t bytes = (size_t)r * c->row_stride;     // row_stride is a multiple of 16
    // my lowest r rows -> the upper halo of the band below; my highest r rows -> the lower halo of the band above
    launch_halo_push(c->stream,
                     lo.on ? src + (size_t)(c->y0 - c->ey0) * c->row_stride : nullptr, lo.on ? lo.res[in] + (size_t)(c->y0 - lo.ey0) * lo.row_stride : nullptr, lo.on ? bytes : 0,
                     hi.on ? src + (size_t)(c->y1 - r - c->ey0) * c->row_stride : nullptr, hi.on ? hi.res[in] + (size_t)(c->y1 - r - hi.ey0) * hi.row_stride : nullptr, hi.on ? bytes : 0,
                     lo.on ? my + 0 : nullptr, hi.on ? my + 1 : nullptr, wait,
                     lo.on ? lo.flags + 1 : nullptr, hi.on ? hi.flags + 0 : nullptr, token,
                     nullptr, nullptr, my + 4, (unsigned int*)(my + 5));
    RCHECK(c, cudaGetLastError());
    return ROMIS_OK;
}

// The same stage 0 done by the temporal pass itself (k_temporal.cu): its blocks within `radius` rows of an edge store their rows
// into the neighbours' halo rows of `out` as well, after the neighbours' last token of the previous frame, and publish the stage
// token.  Used whenever the temporal pass runs; the copy kernel above remains for frames without one.
static HaloDev halo_for_stage0(romis_ctx* c, int out, const dim3& grid, const dim3& block) {
    HaloDev hd; std::memset(&hd, 0, sizeof hd);
    uint32_t* my = (uint32_t*)c->flags.p;
    const int r = (int)c->fr.f.spatialResampleRadius, rows = c->y1 - c->y0, bh = (int)block.y, nby = (int)grid.y;
    const int nlo = std::min(nby, (r + bh - 1) / bh), gh0 = std::max(0, (rows - r) / bh);
    for (int e = 0; e < 2; e++) {
        const romis_ctx::Peer& p = c->peer[e];
        if (!p.on) continue;
        hd.peer_out[e] = p.res[out];
        hd.peer_stride[e] = p.row_stride; hd.peer_ey0[e] = p.ey0;
        hd.wait_flag[e] = my + e;
        hd.sig_flag[e] = p.flags + (e == 0 ? 1 : 0);
        hd.edge_blocks[e] = (unsigned int)((e == 0 ? nlo : nby - gh0) * (int)grid.x);
    }
    hd.counter = (unsigned int*)(my + 6);
    hd.err = my + 4;
    hd.wait_token = c->epoch; hd.token = ++c->epoch;
    hd.push = 1;
    hd.r = r;
    return hd;
}

// Fused halo exchange of spatial pass `pass` (HaloDev, k_spatial.cu spatial_halo_kernel): stage tokens are a running count
// over the frames, identical on every band because all bands issue the same sequence of stages.
static HaloDev halo_for_pass(romis_ctx* c, int out, int pass, const dim3& grid, const dim3& block) {
    HaloDev hd; std::memset(&hd, 0, sizeof hd);
    uint32_t* my = (uint32_t*)c->flags.p;
    const int r = (int)c->fr.f.spatialResampleRadius, rows = c->y1 - c->y0, bh = (int)block.y, nby = (int)grid.y;
    const bool last = pass + 1 == (int)c->fr.f.spatialResamplingPasses;
    const int nlo = std::min(nby, (r + bh - 1) / bh), gh0 = std::max(0, (rows - r) / bh);
    for (int e = 0; e < 2; e++) {
        const romis_ctx::Peer& p = c->peer[e];
        if (!p.on) continue;
        hd.peer_out[e] = last ? nullptr : p.res[out];
        hd.peer_stride[e] = p.row_stride; hd.peer_ey0[e] = p.ey0;
        hd.wait_flag[e] = my + e;
        hd.sig_flag[e] = p.flags + (e == 0 ? 1 : 0);
        hd.edge_blocks[e] = (unsigned int)((e == 0 ? nlo : nby - gh0) * (int)grid.x);
    }
    hd.counter = (unsigned int*)(my + 6);
    hd.err = my + 4;
    hd.wait_token = c->epoch; hd.token = ++c->epoch;
    hd.nl = c->peer[0].on ? nlo : 0;
    hd.gh0 = std::max(gh0, hd.nl);
    hd.nh = c->peer[1].on ? nby - hd.gh0 : 0;
    hd.push = last ? 0 : 1;
    hd.r = r;
    return hd;
}

// Per-row count of pixels whose primary ray hits geometry, for the whole frame: most of the work of every pass sits in
// hit pixels (miss pixels short-circuit), so hosts use this profile as a first cut of the frame into equal-cost row bands.
// Every rank computes the same profile from the same scene and camera, so the band edges agree without communication.
extern "C" int romis_row_hit_counts(romis_ctx* c, const romis_camera* cam, int W, int H, uint32_t* hits_per_row) {
    if (c && !c->kids.empty()) c = c->kids[0];
    if (!c || !cam || !hits_per_row || W < 1 || H < 1) return ROMIS_ERR_INVALID;
    if (!c->has_scene) return fail(c, ROMIS_ERR_STATE, "no scene uploaded");
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_row_hit_counts: frame in flight");
    RCHECK(c, cudaSetDevice(c->device));
    DevBuf tn, mesh, uv, rows;
    auto done = [&](int code) { tn.release(); mesh.release(); uv.release(); rows.release(); return code; };
    const size_t px = (size_t)W * H;
    cudaError_t e = tn.ensure(px * sizeof(float4));
    if (e == cudaSuccess) e = mesh.ensure(px * sizeof(uint32_t));
    if (e == cudaSuccess) e = uv.ensure(c->sc.has_textures ? px * sizeof(float2) : 16);
    if (e == cudaSuccess) e = rows.ensure((size_t)H * sizeof(uint32_t));
    if (e != cudaSuccess) return done(fail(c, ROMIS_ERR_NOMEM, std::string("romis_row_hit_counts: ") + cudaGetErrorString(e)));
    FrameDev fr; std::memset(&fr, 0, sizeof fr);
    fr.cam.origin.x = cam->origin[0]; fr.cam.origin.y = cam->origin[1]; fr.cam.origin.z = cam->origin[2];
    fr.cam.qw = cam->quat[0]; fr.cam.qx = cam->quat[1]; fr.cam.qy = cam->quat[2]; fr.cam.qz = cam->quat[3];
    fr.cam.half_w = cam->half_width; fr.cam.half_h = cam->half_height;
    fr.W = W; fr.H = H; fr.y0 = 0; fr.y1 = H; fr.ey0 = 0; fr.ey1 = H;
    GBufDev g; g.tn = (float4*)tn.p; g.mesh = (uint32_t*)mesh.p; g.uv = (float2*)uv.p;
    launch_primary(c->stream, grid_for(W, H, kBlockS), kBlockS, c->sc, fr, g, 0, H);
    launch_row_hits(c->stream, g, W, H, c->sc.n_meshes, (uint32_t*)rows.p);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(hits_per_row, rows.p, (size_t)H * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return done(fail(c, ROMIS_ERR_CUDA, std::string("romis_row_hit_counts: ") + cudaGetErrorString(e)));
    return done(ROMIS_OK);
}

extern "C" int romis_band_prepare(romis_ctx* c, const romis_features* f, int W, int H) {
    ROMIS_NOT_ON_GROUP(c, "romis_band_prepare");
    if (!c || !f) return ROMIS_ERR_INVALID;
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_band_prepare: frame in flight");
    if (!c->has_scene) return fail(c, ROMIS_ERR_STATE, "no scene uploaded");
    if (f->numSamplesInReservoir < 1 || f->numSamplesInReservoir > 32 || W < 1 || H < 1) return fail(c, ROMIS_ERR_INVALID, "romis_band_prepare: bad parameters");
    RCHECK(c, cudaSetDevice(c->device));
    int rc = ensure_frame_buffers(c, f, W, H);
    if (rc) return rc;
    RCHECK(c, c->flags.ensure(8 * sizeof(uint32_t)));
    RCHECK(c, cudaMemsetAsync(c->flags.p, 0, 8 * sizeof(uint32_t), c->stream));
    RCHECK(c, cudaStreamSynchronize(c->stream));
    c->epoch = 0;
    return ROMIS_OK;
}

extern "C" int romis_peer_export(romis_ctx* c, void* blob) {
    ROMIS_NOT_ON_GROUP(c, "romis_peer_export");
    if (!c || !blob) return ROMIS_ERR_INVALID;
    if (!c->W || !c->flags.p) return fail(c, ROMIS_ERR_STATE, "romis_peer_export: call romis_band_prepare first");
    RCHECK(c, cudaSetDevice(c->device));
    PeerBlob b; std::memset(&b, 0, sizeof b);
    b.magic = 0x524d5042u; b.W = c->W; b.H = c->H; b.N = c->N; b.y0 = c->y0; b.y1 = c->y1; b.ey0 = c->ey0; b.ey1 = c->ey1;
    b.row_stride = c->row_stride;
    for (int i = 0; i < 3; i++) RCHECK(c, cudaIpcGetMemHandle(&b.res[i], c->res[i].p));
    RCHECK(c, cudaIpcGetMemHandle(&b.flags, c->flags.p));
    std::memset(blob, 0, ROMIS_PEER_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof b);
    c->exported = true;
    return ROMIS_OK;
}

static int attach_one(romis_ctx* c, int side, const void* blob) {
    PeerBlob b; std::memcpy(&b, blob, sizeof b);
    if (b.magic != 0x524d5042u) return fail(c, ROMIS_ERR_INVALID, "romis_peer_attach: not a peer blob");
    if (b.W != c->W || b.H != c->H || b.N != c->N || b.row_stride != c->row_stride) return fail(c, ROMIS_ERR_INVALID, "romis_peer_attach: neighbour renders a different frame geometry");
    if ((side == 0 && b.y1 != c->y0) || (side == 1 && b.y0 != c->y1)) return fail(c, ROMIS_ERR_INVALID, "romis_peer_attach: bands are not adjacent");
    romis_ctx::Peer& p = c->peer[side];
    p.ipc = true;
    for (int i = 0; i < 3; i++) RCHECK(c, cudaIpcOpenMemHandle((void**)&p.res[i], b.res[i], cudaIpcMemLazyEnablePeerAccess));
    RCHECK(c, cudaIpcOpenMemHandle((void**)&p.flags, b.flags, cudaIpcMemLazyEnablePeerAccess));
    p.y0 = b.y0; p.y1 = b.y1; p.ey0 = b.ey0; p.ey1 = b.ey1; p.row_stride = (size_t)b.row_stride;
    p.on = true;
    return ROMIS_OK;
}

extern "C" int romis_peer_detach(romis_ctx* c) {
    ROMIS_NOT_ON_GROUP(c, "romis_peer_detach");
    if (!c) return ROMIS_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int s = 0; s < 2; s++) {
        romis_ctx::Peer& p = c->peer[s];
        if (!p.on && !p.ipc) continue;
        if (p.ipc) {
            for (int i = 0; i < 3; i++) if (p.res[i]) cudaIpcCloseMemHandle(p.res[i]);
            if (p.flags) cudaIpcCloseMemHandle(p.flags);
        }
        p = romis_ctx::Peer();
    }
    c->exported = false;
    return ROMIS_OK;
}

extern "C" int romis_peer_attach(romis_ctx* c, const void* low_blob, const void* high_blob) {
    ROMIS_NOT_ON_GROUP(c, "romis_peer_attach");
    if (!c) return ROMIS_ERR_INVALID;
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_peer_attach: frame in flight");
    if (!c->W || !c->flags.p) return fail(c, ROMIS_ERR_STATE, "romis_peer_attach: call romis_band_prepare first");
    if ((low_blob != nullptr) != (c->y0 > 0) || (high_blob != nullptr) != (c->y1 < c->H))
        return fail(c, ROMIS_ERR_INVALID, "romis_peer_attach: need exactly the neighbours this band has");
    if (c->y1 - c->y0 < c->halo) return fail(c, ROMIS_ERR_INVALID, "band has fewer rows than the spatial radius");
    RCHECK(c, cudaSetDevice(c->device));
    int rc = ROMIS_OK;
    if (low_blob) rc = attach_one(c, 0, low_blob);
    if (rc == ROMIS_OK && high_blob) rc = attach_one(c, 1, high_blob);
    if (rc != ROMIS_OK) { std::string keep = c->err; romis_peer_detach(c); c->err = keep; }
    return rc;
}

extern "C" int romis_peer_error(romis_ctx* c, int* timed_out) {
    ROMIS_NOT_ON_GROUP(c, "romis_peer_error");
    if (!c || !timed_out) return ROMIS_ERR_INVALID;
    *timed_out = 0;
    if (!c->flags.p) return ROMIS_OK;
    RCHECK(c, cudaSetDevice(c->device));
    uint32_t e = 0;
    RCHECK(c, cudaMemcpy(&e, (uint32_t*)c->flags.p + 4, 4, cudaMemcpyDeviceToHost));
    *timed_out = (int)e;
    return ROMIS_OK;
}

extern "C" int romis_frame_spatial_pass(romis_ctx* c, int pass) {
    if (!c) return ROMIS_ERR_INVALID;
    ROMIS_NOT_ON_GROUP(c, "romis_frame_spatial_pass");
    if (!c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_frame_spatial_pass: no frame in flight");
    if (!c->fr.f.spatialReuse || pass != c->next_pass || pass >= (int)c->fr.f.spatialResamplingPasses)
        return fail(c, ROMIS_ERR_STATE, "romis_frame_spatial_pass: unexpected pass index");
    RCHECK(c, cudaSetDevice(c->device));
    // ping-pong between the two work buffers; the history buffer is never written during a frame
    const int in = c->cur, out = c->spare;
    const dim3 gOwn = grid_for(c->W, c->y1 - c->y0);
    const bool fine_out = fine_enabled(c) && 1 + pass < ROMIS_FINE_STAGES;
    const FineDev fd = fine_for(c, c->fine_src, fine_out ? 1 + pass : -1, (int)c->fr.f.spatialResampleRadius);
    c->fine_src = fine_out ? 1 + pass : -1;
    if (c->peer[0].on || c->peer[1].on) {
        if (pass == 0 && !c->stage0_pushed) { int prc = peer_push_stage0(c, in); if (prc) return prc; c->n_launches++; RCHECK(c, mark(c, 6, pass)); }
        const HaloDev hd = halo_for_pass(c, out, pass, gOwn, kBlock);
        launch_spatial_halo(c->stream, gOwn, kBlock, c->N, c->fr.f.unbiasedCombination != 0, c->sc, c->fr, gbuf(c), resbuf(c, in), resbuf(c, out), pass, hd, fd);
    } else
        launch_spatial(c->stream, gOwn, kBlock, c->N, c->fr.f.unbiasedCombination != 0, c->sc, c->fr, gbuf(c), resbuf(c, in), resbuf(c, out), pass, fd);
    c->n_launches++;
    RCHECK(c, cudaGetLastError());
    RCHECK(c, mark(c, 4, pass));
    int rc = capture_stage(c, ROMIS_PASS_SPATIAL0 + pass, out);
    if (rc) return rc;
    c->spare = in;
    c->cur = out;
    c->next_pass++;
    return ROMIS_OK;
}

static int frame_end_enqueue(romis_ctx* c, float* out_rgb);
static int frame_end_wait(romis_ctx* c, float* out_rgb) {
    if (out_rgb) {
        RCHECK(c, cudaSetDevice(c->device));
        RCHECK(c, cudaStreamSynchronize(c->copy_stream));
        RCHECK(c, cudaStreamSynchronize(c->stream));
    }
    return ROMIS_OK;
}
extern "C" int romis_frame_end(romis_ctx* c, float* out_rgb) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) return fail(c, ROMIS_ERR_INVALID, "stepwise frames are per-device calls: not available on a multi-device context");
    int rc = frame_end_enqueue(c, out_rgb);
    return rc ? rc : frame_end_wait(c, out_rgb);
}
static int frame_end_enqueue(romis_ctx* c, float* out_rgb) {
    if (!c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_frame_end: no frame in flight");
    if (c->fr.f.spatialReuse && c->next_pass != (int)c->fr.f.spatialResamplingPasses)
        return fail(c, ROMIS_ERR_STATE, "romis_frame_end: spatial passes missing");
    RCHECK(c, cudaSetDevice(c->device));
    // Shade + read-back.  With a host destination the band is shaded in a few row chunks and every chunk's rows start
    // their device-to-host copy (second stream) while the next chunk is being shaded, so most of the PCIe time of the
    // 12 B/pixel image hides behind the shade kernel.  Band rows [a, b) live at image rows [H - b, H - a) of the flipped
    // Screen layout (screen.cpp:37-43): one contiguous range per chunk.
    const int rows = c->y1 - c->y0;
    const int chunks = out_rgb ? std::max(1, std::min(8, rows / 96)) : 1;
    for (int k = 0; k < chunks; k++) {
        const int a = c->y0 + (int)((long long)rows * k / chunks), b = c->y0 + (int)((long long)rows * (k + 1) / chunks);
        FrameDev fr = c->fr; fr.y0 = a; fr.y1 = b;
        launch_shade(c->stream, grid_for(c->W, b - a, kBlockS), kBlockS, c->N, c->sc, fr, gbuf(c), resbuf(c, c->cur), (float*)c->rgb.p,
                     fine_for(c, c->fine_src, -1, 0));
        c->n_launches++;
        RCHECK(c, cudaGetLastError());
        if (out_rgb) {
            RCHECK(c, cudaEventRecord(c->ev_chunk[k], c->stream));
            RCHECK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_chunk[k], 0));
            const size_t off = (size_t)(c->H - b) * c->W * 3, cnt = (size_t)(b - a) * c->W * 3;
            RCHECK(c, cudaMemcpyAsync(out_rgb + off, (const float*)c->rgb.p + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
        }
    }
    RCHECK(c, mark(c, 5, 0));
    RCHECK(c, cudaEventRecord(c->ev_end, c->stream));
    int rc = capture_stage(c, ROMIS_PASS_FINAL, c->cur);
    if (rc) return rc;
    // the returned grid becomes next frame's previousFrameGrid (main.cpp:165)
    c->hist = c->cur;
    c->history_valid = true;
    c->in_frame = false;
    return ROMIS_OK;
}

static int render_common(romis_ctx* c, const romis_features* f, const romis_camera* cam, int W, int H, int history_valid,
                         const romis_rng* rng, float* out_rgb) {
    int rc = romis_frame_begin(c, f, cam, W, H, history_valid, rng);
    if (rc) return rc;
    if (f->spatialReuse)
        for (int p = 0; p < (int)f->spatialResamplingPasses; p++)
            if ((rc = romis_frame_spatial_pass(c, p))) { c->in_frame = false; return rc; }
    rc = romis_frame_end(c, out_rgb);
    if (rc) c->in_frame = false;
    return rc;
}


// ------------------------------------------------------------------------------------------------
// multi-device context: ONE caller, ONE process, several GPUs (romis_create with n_devices > 1)
// ------------------------------------------------------------------------------------------------
// The reference calls its frame from one thread of one process (main.cpp:164, ui.cpp:161), so the drop-in must be able to use
// every GPU of the box from there.  The parent context owns one child context per device; each child renders a row band of
// the frame with the same kernels and the same fused halo exchange as the one-process-per-GPU mode, only the neighbours'
// buffers are reached through cudaDeviceEnablePeerAccess instead of CUDA IPC.  All launches are asynchronous, so a single host
// thread keeps every device busy; the bands' rows land in the caller's single out_rgb.  Band edges: equal COST from the
// per-row hit profile (romis_row_hit_counts), fixed until the resolution, N or radius changes (moving an edge would drop the
// rows' temporal history).
static int group_fail(romis_ctx* g, romis_ctx* kid, int rc) { g->err = kid->err; return rc; }

static int group_create(const int* device_ids, int n, romis_ctx** out, std::string& err) {
    romis_ctx* g = new (std::nothrow) romis_ctx();
    if (!g) { err = "out of host memory"; return ROMIS_ERR_NOMEM; }
    g->device = device_ids[0];
    for (int i = 0; i < n; i++) {
        romis_ctx* k = nullptr;
        int rc = romis_create(&device_ids[i], 1, &k);
        if (rc != ROMIS_OK) { err = romis_last_error(nullptr); for (romis_ctx* q : g->kids) romis_destroy(q); delete g; return rc; }
        k->arch_auto = false;           // the bands recycle light-archive slots together (group_upload_lights)
        g->kids.push_back(k);
    }
    for (int i = 0; i < n; i++)         // neighbouring bands store into each other's halo rows
        for (int j : {i - 1, i + 1}) {
            if (j < 0 || j >= n || device_ids[j] == device_ids[i]) continue;
            int can = 0;
            cudaSetDevice(device_ids[i]);
            cudaError_t e = cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[j]);
            if (e == cudaSuccess && can) { e = cudaDeviceEnablePeerAccess(device_ids[j], 0); if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; } }
            if (e != cudaSuccess || !can) {
                err = "device " + std::to_string(device_ids[i]) + " cannot map the memory of device " + std::to_string(device_ids[j]) + " (peer access)";
                for (romis_ctx* q : g->kids) romis_destroy(q);
                delete g; return ROMIS_ERR_CUDA;
            }
        }
    *out = g;
    return ROMIS_OK;
}

static void group_unwire(romis_ctx* g) {
    for (romis_ctx* k : g->kids) { cudaSetDevice(k->device); if (k->stream) cudaStreamSynchronize(k->stream); k->peer[0] = romis_ctx::Peer(); k->peer[1] = romis_ctx::Peer(); k->exported = false; }
    g->g_wired = false;
}

static std::vector<int> equal_cost_edges(const std::vector<double>& cost, int bands, int min_rows) {
    const int H = (int)cost.size();
    std::vector<double> prefix(H + 1, 0.0);
    for (int y = 0; y < H; y++) prefix[y + 1] = prefix[y] + cost[y];
    std::vector<int> e(1, 0);
    for (int b = 1; b < bands; b++) {
        const double target = prefix[H] * b / bands;
        int cut = (int)(std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin());
        cut = std::max(cut, e.back() + min_rows);
        cut = std::min(cut, H - (bands - b) * min_rows);
        e.push_back(cut);
    }
    e.push_back(H);
    return e;
}

// bands, buffers and neighbour wiring for (W, H, N, radius)
static int group_prepare(romis_ctx* g, const romis_features* f, const romis_camera* cam, int W, int H) {
    const int N = (int)f->numSamplesInReservoir, halo = f->spatialReuse ? (int)f->spatialResampleRadius : 0;
    if (g->g_wired && W == g->g_W && H == g->g_H && N == g->g_N && halo <= g->g_halo) return ROMIS_OK;
    group_unwire(g);
    const int min_rows = std::max(halo, 1);
    const int active = std::max(1, std::min((int)g->kids.size(), H / min_rows));
    std::vector<uint32_t> hits((size_t)H);
    int rc = romis_row_hit_counts(g->kids[0], cam, W, H, hits.data());
    if (rc) return group_fail(g, g->kids[0], rc);
    std::vector<double> cost((size_t)H);
    for (int y = 0; y < H; y++) cost[y] = hits[y] + 0.04 * (W - (double)hits[y]);      // miss pixels short-circuit every pass (measured ratio)
    g->g_edges = equal_cost_edges(cost, active, min_rows);
    for (int i = 0; i < active; i++) {
        romis_ctx* k = g->kids[i];
        if ((rc = romis_set_band(k, g->g_edges[i], g->g_edges[i + 1]))) return group_fail(g, k, rc);
        if ((rc = romis_band_prepare(k, f, W, H))) return group_fail(g, k, rc);
    }
    for (int i = 0; i < active; i++)
        for (int side = 0; side < 2; side++) {
            const int j = side == 0 ? i - 1 : i + 1;
            if (j < 0 || j >= active) continue;
            romis_ctx* k = g->kids[i]; romis_ctx* nb = g->kids[j];
            romis_ctx::Peer& p = k->peer[side];
            for (int b = 0; b < 3; b++) p.res[b] = (unsigned char*)nb->res[b].p;
            p.flags = (uint32_t*)nb->flags.p;
            p.y0 = nb->y0; p.y1 = nb->y1; p.ey0 = nb->ey0; p.ey1 = nb->ey1; p.row_stride = nb->row_stride;
            p.ipc = false; p.on = true;
        }
    g->g_active = active; g->g_W = W; g->g_H = H; g->g_N = N; g->g_halo = halo; g->g_wired = true;
    return ROMIS_OK;
}

static int group_upload_scene(romis_ctx* g, const romis_mesh_desc* meshes, int n_meshes, const romis_texture* textures, int n_textures) {
    group_unwire(g);
    for (romis_ctx* k : g->kids) { int rc = romis_upload_scene(k, meshes, n_meshes, textures, n_textures); if (rc) return group_fail(g, k, rc); }
    return ROMIS_OK;
}

static int group_upload_lights(romis_ctx* g, const romis_light* lights, int n, int first, int count) {
    // halo rows carry light-archive slots from band to band: every band recycles the slots NO band's history holds
    std::vector<uint8_t> keep;
    for (romis_ctx* k : g->kids) {
        if (!k->marks_pending) continue;
        cudaSetDevice(k->device);
        if (cudaEventSynchronize(k->ev_marks) != cudaSuccess) { g->err = "light archive marks"; return ROMIS_ERR_CUDA; }
        if (keep.size() < k->marks_slots) keep.resize(k->marks_slots, 0);
        for (uint32_t s = 0; s < k->marks_slots; s++) keep[s] |= k->arch_mark_host[s];
    }
    for (romis_ctx* k : g->kids) {
        if (!keep.empty()) { int rc = archive_harvest(k, keep.data(), (uint32_t)keep.size()); if (rc) return group_fail(g, k, rc); }
        int rc = upload_lights_impl(k, lights, n, first, count);
        if (rc) return group_fail(g, k, rc);
    }
    return ROMIS_OK;
}

static int group_render_frame(romis_ctx* g, const romis_features* f, const romis_camera* cam, int W, int H, int history_valid,
                              const romis_rng* rng, float* out_rgb) {
    int rc = validate(g->kids[0], f, cam, W, H, rng);
    if (rc) return group_fail(g, g->kids[0], rc);
    if ((rc = group_prepare(g, f, cam, W, H))) return rc;
    const int n = g->g_active;
    auto abort_frame = [&](romis_ctx* k, int code) { for (int i = 0; i < n; i++) g->kids[i]->in_frame = false; return group_fail(g, k, code); };
    for (int i = 0; i < n; i++) if ((rc = romis_frame_begin(g->kids[i], f, cam, W, H, history_valid, rng))) return abort_frame(g->kids[i], rc);
    if (f->spatialReuse)
        for (int p = 0; p < (int)f->spatialResamplingPasses; p++)
            for (int i = 0; i < n; i++) if ((rc = romis_frame_spatial_pass(g->kids[i], p))) return abort_frame(g->kids[i], rc);
    for (int i = 0; i < n; i++) if ((rc = frame_end_enqueue(g->kids[i], out_rgb))) return abort_frame(g->kids[i], rc);
    for (int i = 0; i < n; i++) {
        romis_ctx* k = g->kids[i];
        if (out_rgb) rc = frame_end_wait(k, out_rgb);
        else { cudaSetDevice(k->device); rc = cudaStreamSynchronize(k->stream) == cudaSuccess ? ROMIS_OK : ROMIS_ERR_CUDA; if (rc) k->err = "cudaStreamSynchronize"; }
        if (rc) return group_fail(g, k, rc);
        uint32_t timed_out = 0;
        if (k->flags.p && (k->peer[0].on || k->peer[1].on)) {
            cudaMemcpy(&timed_out, (uint32_t*)k->flags.p + 4, 4, cudaMemcpyDeviceToHost);
            if (timed_out) { g->err = "a band waited for its neighbour's rows for more than 2 s (device " + std::to_string(k->device) + ")"; return ROMIS_ERR_CUDA; }
        }
    }
    return ROMIS_OK;
}

static int group_timings(romis_ctx* g, romis_timings* out) {
    romis_timings t; std::memset(&t, 0, sizeof t);
    for (int i = 0; i < std::max(1, g->g_active); i++) {
        romis_timings k; int rc = romis_last_frame_timings(g->kids[i], &k);
        if (rc) return group_fail(g, g->kids[i], rc);
        t.primary_ms = std::max(t.primary_ms, k.primary_ms); t.initial_ms = std::max(t.initial_ms, k.initial_ms);
        t.temporal_ms = std::max(t.temporal_ms, k.temporal_ms); t.shade_ms = std::max(t.shade_ms, k.shade_ms);
        t.total_ms = std::max(t.total_ms, k.total_ms); t.n_spatial = std::max(t.n_spatial, k.n_spatial);
        for (int p = 0; p < 8; p++) { t.spatial_ms[p] = std::max(t.spatial_ms[p], k.spatial_ms[p]); t.exchange_ms[p] = std::max(t.exchange_ms[p], k.exchange_ms[p]); }
        t.n_launches += k.n_launches;
    }
    *out = t;
    return ROMIS_OK;
}

extern "C" int romis_render_frame(romis_ctx* c, const romis_features* f, const romis_camera* cam, int W, int H,
                                  int history_valid, const romis_rng* rng, float* out_rgb) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) return group_render_frame(c, f, cam, W, H, history_valid, rng, out_rgb);
    int rc = render_common(c, f, cam, W, H, history_valid, rng, out_rgb);
    if (rc == ROMIS_OK && !out_rgb) RCHECK(c, cudaStreamSynchronize(c->stream));
    return rc;
}

extern "C" int romis_render_frame_device(romis_ctx* c, const romis_features* f, const romis_camera* cam, int W, int H,
                                         int history_valid, const romis_rng* rng, const float** dev_rgb) {
    if (!c) return ROMIS_ERR_INVALID;
    ROMIS_NOT_ON_GROUP(c, "romis_render_frame_device (the image is spread over the devices)");
    int rc = render_common(c, f, cam, W, H, history_valid, rng, nullptr);
    if (rc == ROMIS_OK && dev_rgb) *dev_rgb = (const float*)c->rgb.p;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// R-MIS frame (renderRMIS, reference src/rendering/render.cpp:64-119)
// ------------------------------------------------------------------------------------------------
// mode 0 = R-MIS (renderRMIS, render.cpp:64-119), mode 1 = R-OMIS (renderROMIS, render.cpp:121-265)
// Everything of an R-MIS / R-OMIS frame but the final wait: the kernels and the read-back of the band's rows are in the stream.
static int render_mis_enqueue(romis_ctx* c, int mode, const romis_features* f, const romis_rmis_params* rp, const romis_camera* cam,
                              int W, int H, const romis_rng* rng, float* out_rgb, bool wired_ok = false) {
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "frame in flight");
    if (!rp) return fail(c, ROMIS_ERR_INVALID, "null rmis parameters");
    int rc = validate(c, f, cam, W, H, rng);
    if (rc) return rc;
    if (rp->maxIterationsMIS < 1) return fail(c, ROMIS_ERR_INVALID, "maxIterationsMIS must be >= 1");
    if (mode == 0 && rp->misWeightRMIS > ROMIS_MIS_BALANCE) return fail(c, ROMIS_ERR_INVALID, "unhandled MIS weight type (render.cpp:99)");
    if (rp->neighbourSelectionStrategy == ROMIS_NEIGHBOURS_DISSIMILAR)
        return fail(c, ROMIS_ERR_INVALID, "NeighbourSelectionStrategy::Dissimilar is undefined behaviour in the reference (neighbour_selection.cpp:88-93)");
    if (rp->neighbourSelectionStrategy > ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR) return fail(c, ROMIS_ERR_INVALID, "unknown neighbour selection strategy");
    // k = 0 with EqualSimilarDissimilar: `numNeighboursToSample - similarsSampled` wraps in the reference's unsigned arithmetic
    // (neighbour_selection.cpp:95-98) and std::sample then takes the WHOLE window (up to (2r+1)^2 - 1 pixels per pixel)
    if (rp->neighbourSelectionStrategy == ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR && f->numNeighboursToSample == 0)
        return fail(c, ROMIS_ERR_INVALID, "numNeighboursToSample = 0 with EqualSimilarDissimilar is unsigned wrap-around in the reference (neighbour_selection.cpp:95-98)");
    if (rp->neighbourSelectionStrategy != ROMIS_NEIGHBOURS_RANDOM && f->spatialResampleRadius > ROMIS_RMIS_MAX_R)
        return fail(c, ROMIS_ERR_INVALID, "spatialResampleRadius must be <= 30 for similarity-based neighbour selection (ui.cpp:308)");
    if (rp->maxIterationsMIS > 0x7fffffffu - ROMIS_STAGE_RMIS_INITIAL0) return fail(c, ROMIS_ERR_INVALID, "maxIterationsMIS too large");
    const int K1 = (int)f->numNeighboursToSample + 1;
    if (mode == 1) {
        if (K1 > ROMIS_COD_MAX_DIM) return fail(c, ROMIS_ERR_INVALID, "numNeighboursToSample must be <= 10 in R-OMIS mode (ui.cpp:307)");
        if (rp->useProgressiveROMIS && rp->progressiveUpdateMod == 0)
            return fail(c, ROMIS_ERR_INVALID, "progressiveUpdateMod must be >= 1 (the reference takes iteration % progressiveUpdateMod, render.cpp:160)");
        // renderROMIS indexes neighborhood[0 .. k] whatever its size (render.cpp:165,173): a window with fewer than k other
        // pixels makes the reference read unconstructed Reservoirs
        if (rp->neighbourSelectionStrategy != ROMIS_NEIGHBOURS_RANDOM) {
            const long long r1 = (long long)f->spatialResampleRadius + 1;
            if (std::min<long long>(r1, W) * std::min<long long>(r1, H) - 1 < (long long)f->numNeighboursToSample)
                return fail(c, ROMIS_ERR_INVALID, "the resample window holds fewer than numNeighboursToSample pixels: undefined in the reference (render.cpp:165)");
        }
    }
    RCHECK(c, cudaSetDevice(c->device));
    // Row bands (romis_set_band): a band needs the reservoirs and shading contexts of the pixels within the resample radius of its
    // rows.  Everything those depend on is a function of the pixel alone (the random stream is keyed by the global pixel), so a
    // band simply renders its halo rows itself -- primary rays, contexts and every iteration's initial reservoirs over rows
    // [y0 - r, y1 + r) -- instead of exchanging them: no collective, bit-identical to the undivided frame, (rows + 2 r) / rows of
    // the initial-pass work.  Neighbour grid, gather / accumulation, solve and combine run over the band's own rows.
    const int r_halo = (int)f->spatialResampleRadius;
    if ((rc = ensure_frame_buffers(c, f, W, H, c->band_y1 > c->band_y0 ? r_halo : 0))) return rc;
    // (a multi-device context keeps its ReSTIR wiring: nothing of this frame touches a neighbour, and all devices are idle between frames)
    if (!wired_ok && (c->peer[0].on || c->peer[1].on)) return fail(c, ROMIS_ERR_STATE, "R-MIS / R-OMIS frames need no peer mapping: detach first");
    const int hy0 = std::max(0, c->y0 - r_halo), hy1 = std::min(H, c->y1 + r_halo);       // rows rendered incl. halo (within ey0 .. ey1)

    const size_t px = (size_t)W * H;
    RCHECK(c, c->rmis_nb.ensure(px * K1 * sizeof(uint32_t)));
    RCHECK(c, c->rmis_pv.ensure(px * 2 * sizeof(float4)));
    c->rmis_W = W; c->rmis_H = H; c->rmis_K1 = K1;
    RmisDev rm; std::memset(&rm, 0, sizeof rm);
    rm.p = *rp; rm.nb = (uint32_t*)c->rmis_nb.p; rm.K1 = K1; rm.plane = px;
    if (mode == 0) {
        RCHECK(c, c->rmis_acc.ensure(px * sizeof(float4)));
        rm.acc = (float4*)c->rmis_acc.p;
        RCHECK(c, cudaMemsetAsync(c->rmis_acc.p, 0, px * sizeof(float4), c->stream));
    } else {
        RCHECK(c, c->romis_wsum.ensure(px * c->N * sizeof(float)));
        RCHECK(c, c->romis_chosen.ensure(px * c->N * sizeof(float)));
        RCHECK(c, c->romis_tech.ensure(px * K1 * K1 * sizeof(float)));
        RCHECK(c, c->romis_contrib.ensure(px * 3 * K1 * sizeof(float)));
        rm.wsum = (float*)c->romis_wsum.p; rm.chosen = (float*)c->romis_chosen.p;
        rm.tech = (float*)c->romis_tech.p; rm.contrib = (float*)c->romis_contrib.p;
        RCHECK(c, cudaMemsetAsync(rm.tech, 0, px * K1 * K1 * sizeof(float), c->stream));        // MatrixXf::Zero, VectorXf::Zero (:128-131)
        RCHECK(c, cudaMemsetAsync(rm.contrib, 0, px * 3 * K1 * sizeof(float), c->stream));
        if (rp->useProgressiveROMIS) {                                                          // :133-137
            RCHECK(c, c->romis_alpha.ensure(px * 3 * K1 * sizeof(float)));
            RCHECK(c, c->rmis_acc.ensure(px * sizeof(float4)));
            rm.alpha = (float*)c->romis_alpha.p; rm.acc = (float4*)c->rmis_acc.p;
            RCHECK(c, cudaMemsetAsync(rm.alpha, 0, px * 3 * K1 * sizeof(float), c->stream));
            RCHECK(c, cudaMemsetAsync(rm.acc, 0, px * sizeof(float4), c->stream));
        }
    }

    FrameDev& fr = c->fr;
    fr.cam.origin.x = cam->origin[0]; fr.cam.origin.y = cam->origin[1]; fr.cam.origin.z = cam->origin[2];
    fr.cam.qw = cam->quat[0]; fr.cam.qx = cam->quat[1]; fr.cam.qy = cam->quat[2]; fr.cam.qz = cam->quat[3];
    fr.cam.half_w = cam->half_width; fr.cam.half_h = cam->half_height;
    fr.f = *f; fr.seed = rng->seed; fr.frame = rng->frame;
    fr.W = W; fr.H = H; fr.y0 = c->y0; fr.y1 = c->y1; fr.ey0 = c->ey0; fr.ey1 = c->ey1;

    c->marks.clear(); c->n_launches = 0; c->timings_pending = true;
    std::memset(&c->last, 0, sizeof c->last);
    RCHECK(c, cudaEventRecord(c->ev_begin, c->stream));
    RCHECK(c, mark(c, 0, 0));
    const dim3 grid = grid_for(W, c->y1 - c->y0);           // the band's own rows
    const int work = (c->hist + 1) % 3;                     // a work buffer: the ReSTIR history stays untouched
    GBufDev g = gbuf(c); g.pv = (float4*)c->rmis_pv.p;
    const dim3 gridS = grid_for(W, c->y1 - c->y0, kBlockS), gridH = grid_for(W, hy1 - hy0, kBlockS);       // own rows / with halo rows
    FrameDev frH = fr; frH.y0 = hy0; frH.y1 = hy1;          // the initial pass also covers the halo rows
    launch_primary(c->stream, gridH, kBlockS, c->sc, fr, g, hy0, hy1);                     // render.cpp:68 / :125
    launch_ctx(c->stream, gridH, kBlockS, c->sc, fr, g, hy0, hy1);
    RCHECK(c, mark(c, 1, 0));
    launch_rmis_neighbours(c->stream, grid, kBlock, c->sc, fr, g, rm);                // :69 / :126
    RCHECK(c, mark(c, 7, 0));
    c->n_launches += 3;
    RCHECK(c, cudaGetLastError());
    for (uint32_t it = 0; it < rp->maxIterationsMIS; it++) {                                // :72 / :141
        fr.initial_stage = frH.initial_stage = ROMIS_STAGE_RMIS_INITIAL0 + it;
        launch_initial(c->stream, gridH, kBlockS, c->N, c->sc, frH, g, resbuf(c, work), rm.wsum, rm.chosen);    // :74 / :143
        RCHECK(c, mark(c, 8, (int)it));
        if (mode == 0) launch_rmis_gather(c->stream, gridS, kBlockS, c->N, c->sc, fr, g, resbuf(c, work), rm);
        else if (rp->useProgressiveROMIS && it >= 1u && it % rp->progressiveUpdateMod == 0u) {                  // :160-164
            launch_romis_solve(c->stream, grid, kBlock, fr, rm, nullptr, true);
            c->n_launches++;
        }
        if (mode == 1) launch_romis_accumulate(c->stream, grid, kBlock, c->N, c->sc, fr, g, resbuf(c, work), rm);
        RCHECK(c, mark(c, 9, (int)it));
        c->n_launches += 2;
        RCHECK(c, cudaGetLastError());
    }
    fr.initial_stage = ROMIS_STAGE_INITIAL;
    if (mode == 0 || rp->useProgressiveROMIS) launch_rmis_combine(c->stream, grid, kBlock, fr, rm, (float*)c->rgb.p);   // :118 / :232
    else launch_romis_solve(c->stream, grid, kBlock, fr, rm, (float*)c->rgb.p, false);      // :233-262
    RCHECK(c, mark(c, 10, 0));
    c->n_launches++;
    RCHECK(c, cudaGetLastError());
    RCHECK(c, cudaEventRecord(c->ev_end, c->stream));
    if (out_rgb) {          // band rows [y0, y1) are image rows [H - y1, H - y0) of the flipped Screen layout: one contiguous range
        const size_t off = (size_t)(H - c->y1) * W * 3, cnt = (size_t)(c->y1 - c->y0) * W * 3;
        RCHECK(c, cudaMemcpyAsync(out_rgb + off, (const float*)c->rgb.p + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    return ROMIS_OK;
}

// One caller, several GPUs: the R-MIS / R-OMIS frame as one row band per device (bands render their halo rows themselves, so the
// devices exchange nothing), all enqueued from the caller's thread, then awaited.  When the context is already laid out for ReSTIR
// frames of this geometry, the same bands are used and nothing is re-allocated: the temporal history of a ReSTIR sequence
// survives R-MIS / R-OMIS frames in between, as the reference's previousFrameGrid does (render.cpp:268-280, main.cpp:164-165).
static int group_render_mis(romis_ctx* g, int mode, const romis_features* f, const romis_rmis_params* rp, const romis_camera* cam,
                            int W, int H, const romis_rng* rng, float* out_rgb) {
    if (!f || !cam || !rng || W < 1 || H < 1) return fail(g, ROMIS_ERR_INVALID, "null features / camera / rng or empty frame");
    const bool reuse = g->g_wired && g->g_W == W && g->g_H == H && g->g_N == (int)f->numSamplesInReservoir &&
                       (int)f->spatialResampleRadius <= g->g_halo;
    int rc = ROMIS_OK, active = g->g_active;
    const std::vector<int>* edges = &g->g_edges;
    if (!reuse) {
        group_unwire(g);
        active = std::max(1, std::min((int)g->kids.size(), H));
        if (g->g_mis_edges.empty() || g->g_mis_W != W || g->g_mis_H != H || (int)g->g_mis_edges.size() != active + 1) {
            std::vector<uint32_t> hits((size_t)H);
            for (romis_ctx* k : g->kids) romis_set_band(k, 0, 0);
            if ((rc = romis_row_hit_counts(g->kids[0], cam, W, H, hits.data()))) return group_fail(g, g->kids[0], rc);
            std::vector<double> cost((size_t)H);
            for (int y = 0; y < H; y++) cost[y] = hits[y] + 0.04 * (W - (double)hits[y]);
            g->g_mis_edges = equal_cost_edges(cost, active, 1);
            g->g_mis_W = W; g->g_mis_H = H;
        }
        edges = &g->g_mis_edges;
        for (int i = 0; i < active; i++) if ((rc = romis_set_band(g->kids[i], (*edges)[i], (*edges)[i + 1]))) return group_fail(g, g->kids[i], rc);
        g->g_active = active;
    }
    for (int i = 0; i < active; i++) {
        if ((rc = render_mis_enqueue(g->kids[i], mode, f, rp, cam, W, H, rng, out_rgb, reuse))) {
            // the devices before this one already copy into the caller's image: not behind the caller's back after the call returns
            for (int j = 0; j < i; j++) { cudaSetDevice(g->kids[j]->device); cudaStreamSynchronize(g->kids[j]->stream); }
            return group_fail(g, g->kids[i], rc);
        }
    }
    for (int i = 0; i < active; i++) {
        romis_ctx* k = g->kids[i];
        cudaSetDevice(k->device);
        if (cudaStreamSynchronize(k->stream) != cudaSuccess) { k->err = "cudaStreamSynchronize"; return group_fail(g, k, ROMIS_ERR_CUDA); }
    }
    return ROMIS_OK;
}

static int render_mis_frame(romis_ctx* c, int mode, const romis_features* f, const romis_rmis_params* rp, const romis_camera* cam,
                            int W, int H, const romis_rng* rng, float* out_rgb) {
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) return group_render_mis(c, mode, f, rp, cam, W, H, rng, out_rgb);
    int rc = render_mis_enqueue(c, mode, f, rp, cam, W, H, rng, out_rgb);
    if (rc) return rc;
    RCHECK(c, cudaStreamSynchronize(c->stream));
    return ROMIS_OK;
}

extern "C" int romis_render_frame_rmis(romis_ctx* c, const romis_features* f, const romis_rmis_params* rp, const romis_camera* cam,
                                       int W, int H, const romis_rng* rng, float* out_rgb) {
    return render_mis_frame(c, 0, f, rp, cam, W, H, rng, out_rgb);
}

extern "C" int romis_render_frame_romis(romis_ctx* c, const romis_features* f, const romis_rmis_params* rp, const romis_camera* cam,
                                        int W, int H, const romis_rng* rng, float* out_rgb) {
    return render_mis_frame(c, 1, f, rp, cam, W, H, rng, out_rgb);
}

// Parity read-back of the last R-OMIS frame: matrices[H][W][K1][K1], contributions[H][W][3][K1]
extern "C" int romis_download_romis_system(romis_ctx* c, float* matrices, float* contributions) {
    ROMIS_NOT_ON_GROUP(c, "romis_download_romis_system");
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->rmis_K1 || !c->romis_tech.p) return fail(c, ROMIS_ERR_STATE, "romis_download_romis_system: no R-OMIS frame rendered");
    RCHECK(c, cudaSetDevice(c->device));
    RCHECK(c, cudaStreamSynchronize(c->stream));
    const size_t px = (size_t)c->rmis_W * c->rmis_H; const int K1 = c->rmis_K1;
    std::vector<float> t(px * K1 * K1), v(px * 3 * K1);
    RCHECK(c, cudaMemcpy(t.data(), c->romis_tech.p, t.size() * sizeof(float), cudaMemcpyDeviceToHost));
    RCHECK(c, cudaMemcpy(v.data(), c->romis_contrib.p, v.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t p = 0; p < px; p++) {
        if (matrices) for (int i = 0; i < K1; i++) for (int b = 0; b < K1; b++)       // the device keeps the upper triangle
            matrices[p * K1 * K1 + i * K1 + b] = t[(size_t)(std::min(i, b) * K1 + std::max(i, b)) * px + p];
        if (contributions) for (int i = 0; i < 3 * K1; i++) contributions[p * 3 * K1 + i] = v[(size_t)i * px + p];
    }
    return ROMIS_OK;
}

extern "C" int romis_download_rmis_neighbours(romis_ctx* c, int32_t* xy, uint32_t* count) {
    ROMIS_NOT_ON_GROUP(c, "romis_download_rmis_neighbours");
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->rmis_K1 || !c->rmis_nb.p) return fail(c, ROMIS_ERR_STATE, "romis_download_rmis_neighbours: no R-MIS frame rendered");
    RCHECK(c, cudaSetDevice(c->device));
    RCHECK(c, cudaStreamSynchronize(c->stream));
    const size_t px = (size_t)c->rmis_W * c->rmis_H; const int K1 = c->rmis_K1;
    std::vector<uint32_t> nb(px * K1);
    RCHECK(c, cudaMemcpy(nb.data(), c->rmis_nb.p, nb.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (size_t p = 0; p < px; p++) {
        uint32_t n = 0;
        for (int a = 0; a < K1; a++) {
            const uint32_t e = nb[(size_t)a * px + p];
            if (e != 0xffffffffu) n++;
            if (xy) { xy[(p * K1 + a) * 2] = e == 0xffffffffu ? -1 : (int32_t)(e & 0xffffu); xy[(p * K1 + a) * 2 + 1] = e == 0xffffffffu ? -1 : (int32_t)(e >> 16); }
        }
        if (count) count[p] = n;
    }
    return ROMIS_OK;
}

extern "C" int romis_halo_region(romis_ctx* c, int which, void** dev_ptr, size_t* bytes) {
    ROMIS_NOT_ON_GROUP(c, "romis_halo_region");
    if (!c || !dev_ptr || !bytes) return ROMIS_ERR_INVALID;
    if (!c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_halo_region: no frame in flight");
    const int r = c->fr.f.spatialReuse ? (int)c->fr.f.spatialResampleRadius : 0;
    if (c->y1 - c->y0 < r) return fail(c, ROMIS_ERR_INVALID, "band has fewer rows than the spatial radius");
    int row0 = 0, rows = 0;     // local rows (relative to ey0) of the buffer the next spatial pass reads
    switch (which) {
        case ROMIS_HALO_SEND_LOW:  rows = c->y0 > 0 ? r : 0; row0 = c->y0 - c->ey0; break;
        case ROMIS_HALO_RECV_LOW:  rows = c->y0 - std::max(0, c->y0 - r); row0 = (c->y0 - rows) - c->ey0; break;
        case ROMIS_HALO_SEND_HIGH: rows = c->y1 < c->H ? r : 0; row0 = (c->y1 - rows) - c->ey0; break;
        case ROMIS_HALO_RECV_HIGH: rows = std::min(c->H, c->y1 + r) - c->y1; row0 = c->y1 - c->ey0; break;
        default: return fail(c, ROMIS_ERR_INVALID, "romis_halo_region: bad selector");
    }
    *dev_ptr = (unsigned char*)c->res[c->cur].p + (size_t)row0 * c->row_stride;
    *bytes = (size_t)rows * c->row_stride;
    return ROMIS_OK;
}

// ------------------------------------------------------------------------------------------------
// measurement
// ------------------------------------------------------------------------------------------------
extern "C" int romis_last_frame_timings(romis_ctx* c, romis_timings* out) {
    if (!c || !out) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) return group_timings(c, out);
    RCHECK(c, cudaSetDevice(c->device));
    if (c->timings_pending) {
        RCHECK(c, cudaEventSynchronize(c->ev_end));
        romis_timings t; std::memset(&t, 0, sizeof t);
        RCHECK(c, cudaEventElapsedTime(&t.total_ms, c->ev_begin, c->ev_end));
        for (size_t i = 1; i < c->marks.size(); i++) {
            float ms = 0; RCHECK(c, cudaEventElapsedTime(&ms, c->ev_stage[i - 1], c->ev_stage[i]));
            switch (c->marks[i].kind) {
                case 1: t.primary_ms = ms; break;
                case 2: t.initial_ms = ms; break;
                case 3: t.temporal_ms = ms; break;
                case 4: if (c->marks[i].idx < 8) t.spatial_ms[c->marks[i].idx] = ms; t.n_spatial = std::min(8, std::max(t.n_spatial, c->marks[i].idx + 1)); break;
                case 5: t.shade_ms = ms; break;
                case 6: if (c->marks[i].idx < 8) t.exchange_ms[c->marks[i].idx] = ms; break;
                case 7: t.neighbours_ms = ms; break;
                case 8: t.initial_ms += ms; break;
                case 9: t.gather_ms += ms; break;
                case 10: t.resolve_ms = ms; break;
            }
        }
        t.n_launches = c->n_launches;
        c->last = t; c->timings_pending = false;
    }
    *out = c->last;
    return ROMIS_OK;
}

// ------------------------------------------------------------------------------------------------
// parity read-backs
// ------------------------------------------------------------------------------------------------
extern "C" int romis_download_reservoirs(romis_ctx* c, int pass_id, romis_reservoir_dump* out) {
    if (!c || !out) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) ROMIS_GROUP_ACTIVE(c, romis_download_reservoirs(k, pass_id, out));      // every band writes its own rows
    if (c->in_frame) return fail(c, ROMIS_ERR_STATE, "romis_download_reservoirs: frame in flight");
    if (!c->W) return fail(c, ROMIS_ERR_STATE, "romis_download_reservoirs: no frame rendered");
    RCHECK(c, cudaSetDevice(c->device));
    const unsigned char* src = nullptr;
    if (c->capture) {
        auto it = c->captured.find(pass_id);
        if (it != c->captured.end()) src = (const unsigned char*)it->second.p;
    } else if (pass_id == ROMIS_PASS_FINAL && c->history_valid) src = (const unsigned char*)c->res[c->hist].p;
    if (!src) return fail(c, ROMIS_ERR_STATE, "romis_download_reservoirs: pass not captured (romis_set_capture) or not run this frame");
    const size_t n = (size_t)c->N * c->W * c->H;
    struct Slot { void* host; size_t elem; void* dev; } slots[7] = {
        {out->light_id, 4, nullptr}, {out->u, 4, nullptr}, {out->v, 4, nullptr}, {out->W, 4, nullptr}, {out->M, 4, nullptr},
        {out->position, 12, nullptr}, {out->color, 12, nullptr}};
    int rc = ROMIS_OK;
    for (auto& s : slots) if (s.host) {
        if (cudaMalloc(&s.dev, n * s.elem) != cudaSuccess) { rc = fail(c, ROMIS_ERR_NOMEM, "romis_download_reservoirs: device scratch"); break; }
        cudaMemsetAsync(s.dev, 0, n * s.elem, c->stream);
    }
    if (rc == ROMIS_OK) {
        ResBuf b; b.base = (unsigned char*)src; b.row_stride = c->row_stride; b.W = c->W; b.N = c->N;
        launch_dump(c->stream, grid_for(c->W, c->y1 - c->y0), kBlock, c->sc, c->fr, b, c->N, (const uint32_t*)c->arch_orig_dev.p, (uint32_t*)slots[0].dev, (float*)slots[1].dev,
                    (float*)slots[2].dev, (float*)slots[3].dev, (uint32_t*)slots[4].dev, (float*)slots[5].dev, (float*)slots[6].dev);
        cudaError_t e = cudaGetLastError();
        // only this band's rows are written (every band of a multi-device context fills its own part of the caller's arrays)
        const size_t first = (size_t)c->y0 * c->W, cnt = (size_t)(c->y1 - c->y0) * c->W, plane = (size_t)c->H * c->W;
        for (auto& s : slots) if (s.host)
            for (int j = 0; j < c->N && e == cudaSuccess; j++)
                e = cudaMemcpyAsync((char*)s.host + ((size_t)j * plane + first) * s.elem, (char*)s.dev + ((size_t)j * plane + first) * s.elem,
                                    cnt * s.elem, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(c, ROMIS_ERR_CUDA, std::string("romis_download_reservoirs: ") + cudaGetErrorString(e));
    }
    for (auto& s : slots) if (s.dev) cudaFree(s.dev);
    return rc;
}

extern "C" int romis_download_gbuffer(romis_ctx* c, romis_gbuffer_dump* out) {
    if (!c || !out) return ROMIS_ERR_INVALID;
    if (!c->kids.empty()) ROMIS_GROUP_ACTIVE(c, romis_download_gbuffer(k, out));
    if (!c->W) return fail(c, ROMIS_ERR_STATE, "romis_download_gbuffer: no frame rendered");
    RCHECK(c, cudaSetDevice(c->device));
    RCHECK(c, cudaStreamSynchronize(c->stream));
    const size_t rows = (size_t)(c->y1 - c->y0), px = rows * c->W, off = (size_t)(c->y0 - c->ey0) * c->W;
    std::vector<float4> tn(px); std::vector<uint32_t> mesh(px); std::vector<float2> uv(c->sc.has_textures ? px : 0);
    RCHECK(c, cudaMemcpy(tn.data(), (const float4*)c->gb_tn.p + off, px * sizeof(float4), cudaMemcpyDeviceToHost));
    RCHECK(c, cudaMemcpy(mesh.data(), (const uint32_t*)c->gb_mesh.p + off, px * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (c->sc.has_textures) RCHECK(c, cudaMemcpy(uv.data(), (const float2*)c->gb_uv.p + off, px * sizeof(float2), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < px; i++) {
        size_t p = (size_t)c->y0 * c->W + i;
        if (out->t) out->t[p] = tn[i].x;
        if (out->normal) { out->normal[3 * p] = tn[i].y; out->normal[3 * p + 1] = tn[i].z; out->normal[3 * p + 2] = tn[i].w; }
        if (out->mesh) out->mesh[p] = mesh[i];
        if (out->texcoord) { out->texcoord[2 * p] = c->sc.has_textures ? uv[i].x : 0.0f; out->texcoord[2 * p + 1] = c->sc.has_textures ? uv[i].y : 0.0f; }
    }
    return ROMIS_OK;
}

extern "C" int romis_trace_rays(romis_ctx* c, const float* origins, const float* dirs, const float* tfar, int n, int any_hit,
                                uint8_t* hit, float* t, float* u, float* v, uint32_t* tri) {
    if (c && !c->kids.empty()) c = c->kids[0];
    if (!c) return ROMIS_ERR_INVALID;
    if (!c->has_scene) return fail(c, ROMIS_ERR_STATE, "romis_trace_rays: no scene");
    if (n < 0 || (n > 0 && (!origins || !dirs || !tfar || !hit))) return fail(c, ROMIS_ERR_INVALID, "romis_trace_rays: bad arguments");
    if (n == 0) return ROMIS_OK;
    RCHECK(c, cudaSetDevice(c->device));
    DevBuf d_o, d_d, d_tf, d_hit, d_t, d_u, d_v, d_tri;
    int rc = ROMIS_OK;
    auto done = [&](int code) { for (DevBuf* b : {&d_o, &d_d, &d_tf, &d_hit, &d_t, &d_u, &d_v, &d_tri}) b->release(); return code; };
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    ok(d_o.ensure((size_t)n * 12)); ok(d_d.ensure((size_t)n * 12)); ok(d_tf.ensure((size_t)n * 4)); ok(d_hit.ensure((size_t)n));
    ok(d_t.ensure((size_t)n * 4)); ok(d_u.ensure((size_t)n * 4)); ok(d_v.ensure((size_t)n * 4)); ok(d_tri.ensure((size_t)n * 4));
    if (e == cudaSuccess) {
        ok(cudaMemcpyAsync(d_o.p, origins, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
        ok(cudaMemcpyAsync(d_d.p, dirs, (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
        ok(cudaMemcpyAsync(d_tf.p, tfar, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        ok(cudaMemsetAsync(d_t.p, 0, (size_t)n * 4, c->stream)); ok(cudaMemsetAsync(d_u.p, 0, (size_t)n * 4, c->stream));
        ok(cudaMemsetAsync(d_v.p, 0, (size_t)n * 4, c->stream)); ok(cudaMemsetAsync(d_tri.p, 0xff, (size_t)n * 4, c->stream));
        launch_trace(c->stream, n, c->sc, (const float*)d_o.p, (const float*)d_d.p, (const float*)d_tf.p, any_hit,
                     (uint8_t*)d_hit.p, (float*)d_t.p, (float*)d_u.p, (float*)d_v.p, (uint32_t*)d_tri.p);
        ok(cudaGetLastError());
        ok(cudaMemcpyAsync(hit, d_hit.p, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        if (t) ok(cudaMemcpyAsync(t, d_t.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (u) ok(cudaMemcpyAsync(u, d_u.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (v) ok(cudaMemcpyAsync(v, d_v.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        if (tri) ok(cudaMemcpyAsync(tri, d_tri.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        ok(cudaStreamSynchronize(c->stream));
    }
    if (e != cudaSuccess) rc = fail(c, ROMIS_ERR_CUDA, std::string("romis_trace_rays: ") + cudaGetErrorString(e));
    return done(rc);
}

// ------------------------------------------------------------------------------------------------
extern "C" int romis_selftest_division(romis_ctx* c, const float* num, const float* den, int n, float* out_fast, float* out_ref) {
    if (c && !c->kids.empty()) c = c->kids[0];
    if (!c || !num || !den || !out_fast || !out_ref || n < 1) return ROMIS_ERR_INVALID;
    RCHECK(c, cudaSetDevice(c->device));
    DevBuf dn, dd, df, dr;
    auto done = [&](int code) { dn.release(); dd.release(); df.release(); dr.release(); return code; };
    const size_t b3 = (size_t)n * 3 * sizeof(float), b1 = (size_t)n * sizeof(float);
    cudaError_t e = dn.ensure(b3);
    if (e == cudaSuccess) e = dd.ensure(b1);
    if (e == cudaSuccess) e = df.ensure(b3);
    if (e == cudaSuccess) e = dr.ensure(b3);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dn.p, num, b3, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dd.p, den, b1, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) { launch_division_selftest(c->stream, (const float*)dn.p, (const float*)dd.p, n, (float*)df.p, (float*)dr.p); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_fast, df.p, b3, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_ref, dr.p, b3, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return done(fail(c, ROMIS_ERR_CUDA, std::string("romis_selftest_division: ") + cudaGetErrorString(e)));
    return done(ROMIS_OK);
}

extern "C" void* romis_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void romis_host_free(void* p) { if (p) cudaFreeHost(p); }
// Page-locks memory the caller already owns (the reference's Screen keeps its pixels in a std::vector, screen.h): the image
// read-back of romis_render_frame into it then runs as asynchronous DMA.
extern "C" int romis_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return ROMIS_ERR_INVALID;
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? ROMIS_OK : (cudaGetLastError(), ROMIS_ERR_CUDA);
}
extern "C" int romis_host_unregister(void* p) {
    if (!p) return ROMIS_ERR_INVALID;
    return cudaHostUnregister(p) == cudaSuccess ? ROMIS_OK : (cudaGetLastError(), ROMIS_ERR_CUDA);
}
