/*
 * restir_oracle.c -- CPU restatement of the reference's ReSTIR frame (renderReSTIR).
 *
 * TEST ORACLE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs; never by the product path (romis_b200/, libromis_gpu.so).
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's own translation units compiled here with an
 * injected random stream (oracle/_ref/libromis_ref.so, built by oracle/Makefile `ref`):
 * tests/test_oracle_vs_ref.py compares G-buffer, per-stage reservoirs and image bit for bit, and
 * tests/golden/ holds vectors generated from that library (tests/golden/gen_golden.py).
 * Intersection results are defined by oracle/tracer.c (Embree is absent and unpinned).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 * Arithmetic is fp32 in GLM 0.9.9.9's scalar operation order, compiled with
 * -ffp-contract=off; pow/exp are romis_powf/romis_expf (include/romis_detmath.h), random draws come
 * from include/romis_rng.h with the counter layout documented there.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "romis_gpu.h"
#include "romis_rng.h"
#include "romis_detmath.h"
#include "romis_cod.h"
#include "tracer.h"

#define ORC_MAX_N 32
#define ORC_MAX_STAGES 12      /* initial, temporal, spatial 0..7, final */
#define ORC_NO_LIGHT 0xffffffffu

typedef struct { float x, y, z; } v3;
static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scale3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 div3(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }      /* glm type_vec3.inl:717-723 */
/* glm func_geometric.inl:48-55: products first, then (x + y) + z */
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* glm func_geometric.inl:68-79 */
static inline v3 cross3(v3 a, v3 b) { return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
/* glm func_geometric.inl:8-14 */
static inline float length3(v3 a) { return sqrtf(dot3(a, a)); }
/* glm func_geometric.inl:82-90 + func_exponential.inl:136-139: v * (1 / sqrt(dot)) */
static inline v3 normalize3(v3 a) { float s = 1.0f / sqrtf(dot3(a, a)); return scale3(a, s); }
/* glm func_common.inl:104-112: x*(1-a) + y*a */
static inline v3 mix3(v3 x, v3 y, float a) { return add3(scale3(x, 1.0f - a), scale3(y, a)); }
static inline int anynan3(v3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }

typedef struct { v3 kd, ks; float shininess; int32_t tex; } orc_material;
typedef struct { float* px; int w, h; } orc_texture;
typedef struct { float t; v3 n; float uv[2]; uint32_t mesh; } orc_hit;      /* Ray::t + HitInfo (common.h:43-54) */
typedef struct {                                                             /* one sub-reservoir (reservoir.h:18-42) */
    uint32_t light; float u, v;     /* which light / where on it: bookkeeping the reference does not keep */
    v3 pos, col;                    /* LightSample */
    float W;                        /* outputWeight */
    uint64_t M;                     /* sampleNums */
    float wSum, chosen;             /* wSums, chosenSampleWeights */
} orc_sub;

typedef struct orc_ctx {
    otr_tracer* tracer;
    int tracer_mode;
    int ntri, nmesh;
    uint32_t* tri_mesh;
    romis_vertex* tri_verts;        /* 3 per global triangle */
    orc_material* mats;             /* nmesh + 1 (last = value-initialised Material of a miss pixel) */
    orc_texture* tex; int ntex;
    romis_light* lights; int nlights;
    int W, H, N;
    orc_hit* gbuf;
    orc_sub* cur; orc_sub* prev; orc_sub* tmp;
    int have_prev;
    orc_sub* stage_dump[ORC_MAX_STAGES]; int stage_valid[ORC_MAX_STAGES]; int capture;
    char err[256];
} orc_ctx;

/* ------------------------------------------------------------------------------------------- */
typedef struct { const romis_features* f; const orc_ctx* c; v3 origin; } orc_env;

/* Trackball::generateRay (framework/src/trackball.cpp:105-114) with the pixel -> NDC map of
 * genPrimaryRayHits (src/rendering/render_utils.cpp:24-25); quat * vec3 per glm type_quat.inl:347-354. */
static v3 gen_ray_dir(const romis_camera* cam, int x, int y, int W, int H) {
    float px = (float)x / (float)W * 2.0f - 1.0f;
    float py = (float)y / (float)H * 2.0f - 1.0f;
    v3 cs = normalize3(V3(-px * cam->half_width, py * cam->half_height, 1.0f));
    v3 q = V3(cam->quat[1], cam->quat[2], cam->quat[3]);
    float qw = cam->quat[0];
    v3 uv = cross3(q, cs);
    v3 uuv = cross3(q, uv);
    return add3(cs, scale3(add3(scale3(uv, qw), uuv), 2.0f));
}

/* acquireTexel (src/scene/texture.cpp:4-9): nearest texel, no wrap; the reference indexes out of
 * bounds for uv outside [0,1] -- the index is clamped here instead of reading foreign memory. */
static v3 acquire_texel(const orc_texture* im, const float uv[2]) {
    float fx = uv[0] * (float)(im->w - 1), fy = uv[1] * (float)(im->h - 1);
    int64_t dx = (int64_t)fx, dy = (int64_t)fy;
    int64_t loc = dy * im->w + dx, n = (int64_t)im->w * im->h;
    if (loc < 0) loc = 0;
    if (loc >= n) loc = n - 1;
    return V3(im->px[3 * loc], im->px[3 * loc + 1], im->px[3 * loc + 2]);
}

/* diffuseAlbedo (src/utils/utils.cpp:33-37) */
static v3 diffuse_albedo(const orc_env* e, const orc_hit* h) {
    const orc_material* m = &e->c->mats[h->mesh];
    if (e->f->enableTextureMapping && m->tex >= 0) return acquire_texel(&e->c->tex[m->tex], h->uv);
    return m->kd;
}

/* computeShading (src/rendering/shading.cpp:7-34) */
static v3 compute_shading(const orc_env* e, v3 lightPos, v3 lightCol, v3 dir, const orc_hit* h) {
    const orc_material* m = &e->c->mats[h->mesh];
    if (!e->f->enableShading) return m->kd;                                     /* :8 */
    v3 albedo = diffuse_albedo(e, h);                                           /* :11 */
    v3 P = add3(e->origin, scale3(dir, h->t));                                  /* :12 */
    v3 L = normalize3(sub3(lightPos, P));                                       /* :13 */
    float NL = dot3(h->n, L);                                                   /* :14 */
    if (NL < 0.0f) return V3(0, 0, 0);                                          /* :17 */
    v3 Vv = normalize3(sub3(e->origin, P));                                     /* :20 */
    v3 R = normalize3(sub3(scale3(h->n, 2.0f * NL), L));                        /* :21 */
    float cosTheta = dot3(R, Vv);                                               /* :22 */
    v3 diffuse = scale3(mul3(lightCol, albedo), NL);                            /* :25 */
    v3 specular = scale3(mul3(lightCol, m->ks), romis_powf(cosTheta, m->shininess));   /* :26 */
    if (anynan3(diffuse)) diffuse = V3(0, 0, 0);                                /* :27 */
    if (anynan3(specular)) specular = V3(0, 0, 0);                              /* :28 */
    float dist = length3(sub3(lightPos, P));                                    /* :31 glm::distance = length(p1 - p0) */
    if (fabsf(dist) < 1e-5f) dist = 1.0f;                                       /* :32, utils.cpp:24, ZERO_EPSILON utils.h:19 */
    return div3(add3(diffuse, specular), dist * dist);                          /* :33 */
}

/* targetPDF (src/rendering/reservoir.cpp:106-109) */
static float target_pdf(const orc_env* e, v3 pos, v3 col, v3 dir, const orc_hit* h) {
    return length3(compute_shading(e, pos, col, dir, h));
}

/* testVisibilityLightSample (src/utils/utils.cpp:41-56) -> EmbreeInterface::anyHit (embree_interface.cpp:58-62) */
static int visible(const orc_env* e, v3 samplePos, v3 dir, const orc_hit* h) {
    v3 P = add3(e->origin, scale3(dir, h->t));
    v3 toS = normalize3(sub3(samplePos, P));
    P = add3(P, scale3(toS, 1e-3f));                                            /* SHADOW_RAY_EPSILON utils.h:16 */
    float tfar = length3(sub3(samplePos, P));
    float o[3] = {P.x, P.y, P.z}, d[3] = {toS.x, toS.y, toS.z};
    return !otr_any(e->c->tracer, o, d, tfar);
}

/* Reservoir::Reservoir (src/rendering/reservoir.h:29-32) */
static void reservoir_init(orc_sub* r, int N) {
    for (int j = 0; j < N; j++) {
        r[j].light = ORC_NO_LIGHT; r[j].u = r[j].v = 0.0f;
        r[j].pos = V3(0, 0, 0); r[j].col = V3(0, 0, 0); r[j].W = 0.0f;
        r[j].M = 1; r[j].wSum = FLT_MIN; r[j].chosen = 0.0f;
    }
}

/* Reservoir::update (src/rendering/reservoir.cpp:10-32): exactly one rand() per call */
static int reservoir_update(orc_sub* r, int N, const orc_sub* sample, float weight, romis_stream_key rk, uint32_t* rand_ctr) {
    int idx = 0; float smallest = FLT_MAX;
    for (int j = 0; j < N; j++) if (r[j].wSum < smallest) { idx = j; smallest = r[j].wSum; }   /* :12-19 */
    r[idx].M += 1;                                                                              /* :22 */
    r[idx].wSum += weight;                                                                      /* :23 */
    float u = romis_rand_to_unit(romis_rng_rand(rk, (*rand_ctr)++));                            /* :24, utils.cpp:26-31 */
    if (u < (weight / r[idx].wSum)) {                                                           /* :25 */
        r[idx].light = sample->light; r[idx].u = sample->u; r[idx].v = sample->v;
        r[idx].pos = sample->pos; r[idx].col = sample->col; r[idx].chosen = weight;
    }
    return idx;
}

static uint64_t total_m(const orc_sub* r, int N) { uint64_t s = 0; for (int j = 0; j < N; j++) s += r[j].M; return s; }   /* reservoir.cpp:34-38 */

static v3 lv(const float* p) { return V3(p[0], p[1], p[2]); }

/* genCanonicalSamples (src/scene/light.cpp:39-99) with the samplers of light.cpp:19-34 */
static void gen_canonical(const orc_env* e, const romis_rng* rng, uint32_t stage, uint32_t pixel, v3 dir, const orc_hit* h, orc_sub* r) {
    const orc_ctx* c = e->c; const romis_features* f = e->f; const int N = c->N;
    reservoir_init(r, N);                                                       /* :41 */
    if (c->nlights == 0) return;                                                /* :46 */
    romis_stream_key ek = romis_rng_stream(rng->seed, rng->frame, stage, pixel, ROMIS_STREAM_ENGINE);
    romis_stream_key rk = romis_rng_stream(rng->seed, rng->frame, stage, pixel, ROMIS_STREAM_RAND);
    uint32_t rc = 0;
    for (int j = 0; j < N; j++) r[j].M = 0;                                     /* :58-60 */
    for (uint32_t i = 0; i < f->initialLightSamples; i++) {                     /* :63 */
        orc_sub s; memset(&s, 0, sizeof s);
        uint32_t li = (uint32_t)romis_rng_uniform_int(romis_rng_bits(ek, i), 0, c->nlights - 1);   /* :51,66 */
        const romis_light* l = &c->lights[li];
        s.light = li;
        if (l->type == ROMIS_LIGHT_POINT) {                                     /* :67-70 */
            s.pos = lv(l->p0); s.col = lv(l->c0);
        } else if (l->type == ROMIS_LIGHT_SEGMENT) {                            /* :71-73 -> :19-23 */
            s.u = romis_rand_to_unit(romis_rng_rand(rk, rc++));
            s.pos = mix3(lv(l->p0), lv(l->e1), s.u);
            s.col = mix3(lv(l->c0), lv(l->c1), s.u);
        } else {                                                                /* :74-77 -> :27-34 */
            s.u = romis_rand_to_unit(romis_rng_rand(rk, rc++));
            s.v = romis_rand_to_unit(romis_rng_rand(rk, rc++));
            s.pos = add3(add3(lv(l->p0), scale3(lv(l->e1), s.u)), scale3(lv(l->e2), s.v));
            v3 l01 = mix3(lv(l->c0), lv(l->c1), s.u);
            v3 l23 = mix3(lv(l->c2), lv(l->c3), s.u);
            s.col = mix3(l01, l23, s.v);
        }
        float w = target_pdf(e, s.pos, s.col, dir, h) / (1.0f / (float)c->nlights);   /* :80 */
        reservoir_update(r, N, &s, w, rk, &rc);                                 /* :81 */
    }
    for (int j = 0; j < N; j++) {                                               /* :85-95 */
        if (f->initialSamplesVisibilityCheck && !visible(e, r[j].pos, dir, h)) { r[j].W = 0.0f; continue; }
        float pdf = target_pdf(e, r[j].pos, r[j].col, dir, h);
        if (pdf == 0.0f) r[j].W = 0.0f;
        else r[j].W = (1.0f / pdf) * (1.0f / (float)r[j].M) * r[j].wSum;
    }
}

/* The streaming part shared by combineBiased / combineUnbiased (src/rendering/reservoir.cpp:42-54, 70-82) */
static void combine_stream(const orc_env* e, const orc_sub* const* stream, int ns, v3 dir, const orc_hit* h,
                           orc_sub* out, romis_stream_key rk, uint32_t* rc) {
    const int N = e->c->N;
    uint64_t cnt[ORC_MAX_N];
    for (int j = 0; j < N; j++) cnt[j] = 0;
    reservoir_init(out, N);
    for (int s = 0; s < ns; s++) for (int i = 0; i < N; i++) {
        const orc_sub* smp = &stream[s][i];
        float pdf = target_pdf(e, smp->pos, smp->col, dir, h);
        int j = reservoir_update(out, N, smp, pdf * smp->W * (float)smp->M, rk, rc);
        cnt[j] += smp->M;
    }
    for (int j = 0; j < N; j++) out[j].M = cnt[j];
}

/* Reservoir::combineBiased (src/rendering/reservoir.cpp:40-66) */
static void combine_biased(const orc_env* e, const orc_sub* const* stream, int ns, v3 dir, const orc_hit* h,
                           orc_sub* out, romis_stream_key rk, uint32_t* rc) {
    const int N = e->c->N;
    combine_stream(e, stream, ns, dir, h, out, rk, rc);
    for (int j = 0; j < N; j++) {                                               /* :57-65 */
        float pdf = target_pdf(e, out[j].pos, out[j].col, dir, h);
        if (pdf == 0.0f) out[j].W = 0.0f;
        else out[j].W = (1.0f / pdf) * (1.0f / (float)out[j].M) * out[j].wSum;
    }
}

/* Reservoir::combineUnbiased (src/rendering/reservoir.cpp:68-104); sdir/sh = each stream reservoir's own ray / hit */
static void combine_unbiased(const orc_env* e, const orc_sub* const* stream, const v3* sdir, const orc_hit* const* sh, int ns,
                             v3 dir, const orc_hit* h, orc_sub* out, romis_stream_key rk, uint32_t* rc) {
    const int N = e->c->N;
    combine_stream(e, stream, ns, dir, h, out, rk, rc);
    uint64_t Z[ORC_MAX_N];
    for (int j = 0; j < N; j++) Z[j] = 0;
    for (int s = 0; s < ns; s++) for (int j = 0; j < N; j++) {                  /* :85-93 */
        float pdf = target_pdf(e, out[j].pos, out[j].col, sdir[s], sh[s]);
        if (e->f->spatialReuseVisibilityCheck) pdf *= (float)visible(e, out[j].pos, sdir[s], sh[s]);
        if (pdf > 0.0f) Z[j] += total_m(stream[s], N);
    }
    for (int j = 0; j < N; j++) {                                               /* :96-103 */
        float pdf = target_pdf(e, out[j].pos, out[j].col, dir, h);
        if (pdf == 0.0f || Z[j] == 0) out[j].W = 0.0f;
        else out[j].W = (1.0f / pdf) * (1.0f / (float)Z[j]) * out[j].wSum;
    }
}

static void snapshot(orc_ctx* c, int slot) {
    if (!c->capture) return;
    size_t n = (size_t)c->W * c->H * c->N;
    c->stage_dump[slot] = (orc_sub*)realloc(c->stage_dump[slot], n * sizeof(orc_sub));
    memcpy(c->stage_dump[slot], c->cur, n * sizeof(orc_sub));
    c->stage_valid[slot] = 1;
}

/* ------------------------------------------------------------------------------------------- */
/* C API (mirrors include/romis_gpu.h with an orc_ prefix)                                       */
/* ------------------------------------------------------------------------------------------- */
orc_ctx* orc_create(void) { orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx)); c->tracer_mode = 1; c->capture = 1; return c; }

static void free_scene(orc_ctx* c) {
    otr_free(c->tracer); c->tracer = NULL;
    free(c->tri_mesh); free(c->tri_verts); free(c->mats); c->tri_mesh = NULL; c->tri_verts = NULL; c->mats = NULL;
    for (int i = 0; i < c->ntex; i++) free(c->tex[i].px);
    free(c->tex); c->tex = NULL; c->ntex = 0;
}
void orc_destroy(orc_ctx* c) {
    if (!c) return;
    free_scene(c); free(c->lights); free(c->gbuf); free(c->cur); free(c->prev); free(c->tmp);
    for (int i = 0; i < ORC_MAX_STAGES; i++) free(c->stage_dump[i]);
    free(c);
}
const char* orc_last_error(const orc_ctx* c) { return c ? c->err : "null ctx"; }
int orc_set_capture(orc_ctx* c, int on) { c->capture = on; return 0; }
int orc_set_tracer_mode(orc_ctx* c, int mode) { c->tracer_mode = mode; return 0; }

/* EmbreeInterface::initScene (src/ray_tracing/embree_interface.cpp:30-51): one geometry per Mesh, geomID = mesh index */
int orc_upload_scene(orc_ctx* c, const romis_mesh_desc* meshes, int n_meshes, const romis_texture* textures, int n_textures) {
    free_scene(c);
    int ntri = 0;
    for (int m = 0; m < n_meshes; m++) ntri += (int)meshes[m].n_triangles;
    c->ntri = ntri; c->nmesh = n_meshes;
    c->tri_mesh = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(ntri + 1));
    c->tri_verts = (romis_vertex*)malloc(sizeof(romis_vertex) * (size_t)(3 * ntri + 1));
    c->mats = (orc_material*)calloc((size_t)n_meshes + 1, sizeof(orc_material));
    float* verts = (float*)malloc(sizeof(float) * (size_t)(9 * ntri + 1));
    int g = 0;
    for (int m = 0; m < n_meshes; m++) {
        for (uint32_t t = 0; t < meshes[m].n_triangles; t++, g++) {
            for (int k = 0; k < 3; k++) {
                const romis_vertex* v = &meshes[m].vertices[meshes[m].triangles[3 * t + k]];
                c->tri_verts[3 * g + k] = *v;
                memcpy(verts + 9 * g + 3 * k, v->position, 12);
            }
            c->tri_mesh[g] = (uint32_t)m;
        }
        c->mats[m].kd = lv(meshes[m].material.kd); c->mats[m].ks = lv(meshes[m].material.ks);
        c->mats[m].shininess = meshes[m].material.shininess; c->mats[m].tex = meshes[m].material.kd_texture;
        if (c->mats[m].tex >= n_textures) c->mats[m].tex = -1;
    }
    /* miss pixels keep a value-initialised Material: kd = ks = 0, shininess = 1 (mesh.h:22-34) */
    c->mats[n_meshes].kd = V3(0, 0, 0); c->mats[n_meshes].ks = V3(0, 0, 0);
    c->mats[n_meshes].shininess = 1.0f; c->mats[n_meshes].tex = -1;
    c->tracer = otr_build(verts, ntri, c->tracer_mode);
    free(verts);
    c->ntex = n_textures;
    c->tex = (orc_texture*)calloc((size_t)n_textures + 1, sizeof(orc_texture));
    for (int i = 0; i < n_textures; i++) {
        size_t n = (size_t)textures[i].width * textures[i].height * 3;
        c->tex[i].w = textures[i].width; c->tex[i].h = textures[i].height;
        c->tex[i].px = (float*)malloc(n * sizeof(float));
        memcpy(c->tex[i].px, textures[i].pixels, n * sizeof(float));
    }
    c->have_prev = 0;
    return 0;
}

int orc_upload_lights(orc_ctx* c, const romis_light* lights, int n) {
    free(c->lights);
    c->lights = (romis_light*)malloc(sizeof(romis_light) * (size_t)(n + 1));
    memcpy(c->lights, lights, sizeof(romis_light) * (size_t)n);
    c->nlights = n;
    return 0;
}

int orc_reset_history(orc_ctx* c) { c->have_prev = 0; return 0; }

/* genPrimaryRayHits (src/rendering/render_utils.cpp:13-34) -> closestHit (embree_interface.cpp:64-90) */
static void primary_hits(orc_ctx* c, const romis_camera* cam, v3 origin, int W, int H) {
    #pragma omp parallel for schedule(guided)
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        orc_hit* h = &c->gbuf[(size_t)y * W + x];
        v3 d = gen_ray_dir(cam, x, y, W, H);
        float o[3] = {origin.x, origin.y, origin.z}, dd[3] = {d.x, d.y, d.z};
        float t, u, v; uint32_t tri;
        if (otr_closest(c->tracer, o, dd, FLT_MAX, &t, &u, &v, &tri)) {
            const romis_vertex* a = &c->tri_verts[3 * tri];
            float w = (1.0f - u) - v;           /* attribute interpolation: (w*a + u*b) + v*c (tracer.h) */
            h->t = t;
            h->n = add3(add3(scale3(lv(a[0].normal), w), scale3(lv(a[1].normal), u)), scale3(lv(a[2].normal), v));
            h->uv[0] = (w * a[0].texcoord[0] + u * a[1].texcoord[0]) + v * a[2].texcoord[0];
            h->uv[1] = (w * a[0].texcoord[1] + u * a[1].texcoord[1]) + v * a[2].texcoord[1];
            h->mesh = c->tri_mesh[tri];
        } else {                                /* hitInfo untouched, ray.t = FLT_MAX */
            h->t = FLT_MAX; h->n = V3(0, 0, 0); h->uv[0] = h->uv[1] = 0.0f; h->mesh = (uint32_t)c->nmesh;
        }
    }
}

/* renderReSTIR (src/rendering/render.cpp:28-62) */
int orc_render_frame(orc_ctx* c, const romis_features* f, const romis_camera* cam, int W, int H, int history_valid,
                     const romis_rng* rng, float* out_rgb) {
    if (!c->tracer) { strcpy(c->err, "no scene"); return ROMIS_ERR_STATE; }
    const int N = (int)f->numSamplesInReservoir;
    if (N < 1 || N > ORC_MAX_N || W < 1 || H < 1) { strcpy(c->err, "bad size"); return ROMIS_ERR_INVALID; }
    if (c->W != W || c->H != H || c->N != N) {
        size_t n = (size_t)W * H;
        c->gbuf = (orc_hit*)realloc(c->gbuf, n * sizeof(orc_hit));
        c->cur = (orc_sub*)realloc(c->cur, n * N * sizeof(orc_sub));
        c->prev = (orc_sub*)realloc(c->prev, n * N * sizeof(orc_sub));
        c->tmp = (orc_sub*)realloc(c->tmp, n * N * sizeof(orc_sub));
        c->W = W; c->H = H; c->N = N; c->have_prev = 0;
    }
    for (int i = 0; i < ORC_MAX_STAGES; i++) c->stage_valid[i] = 0;
    if (!history_valid) c->have_prev = 0;
    orc_env env; env.f = f; env.c = c; env.origin = lv(cam->origin);
    const orc_env* e = &env;
    const int k = (int)f->numNeighboursToSample, r = (int)f->spatialResampleRadius;
    if (k > 64) { strcpy(c->err, "numNeighboursToSample > 64"); return ROMIS_ERR_INVALID; }

    /* 1. genPrimaryRayHits */
    primary_hits(c, cam, env.origin, W, H);

    /* 2. genInitialSamples (render_utils.cpp:36-52) */
    #pragma omp parallel for schedule(guided)
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        size_t p = (size_t)y * W + x;
        gen_canonical(e, rng, ROMIS_STAGE_INITIAL, (uint32_t)p, gen_ray_dir(cam, x, y, W, H), &c->gbuf[p], &c->cur[p * N]);
    }
    snapshot(c, 0);

    /* 3. temporalReuse (render_utils.cpp:142-177), only with a predecessor (render.cpp:35) */
    if (f->temporalReuse && c->have_prev) {
        #pragma omp parallel for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            size_t p = (size_t)y * W + x;
            orc_sub* cur = &c->cur[p * N]; orc_sub* prv = &c->prev[p * N];
            uint64_t cap = (uint64_t)f->temporalClampM * total_m(cur, N) + 1ull;          /* :156 */
            if (total_m(prv, N) > cap) {                                                  /* :157 */
                for (int j = 0; j < N; j++) {
                    if (prv[j].M == 0) continue;                                          /* :159 */
                    prv[j].wSum *= (float)(cap / prv[j].M);                               /* :160 integer division; dead value */
                    prv[j].M = cap;                                                       /* :161 */
                }
            }
            orc_sub out[ORC_MAX_N];
            const orc_sub* stream[2] = {cur, prv};                                        /* :169 */
            romis_stream_key rk = romis_rng_stream(rng->seed, rng->frame, ROMIS_STAGE_TEMPORAL, (uint32_t)p, ROMIS_STREAM_RAND);
            uint32_t rc = 0;
            combine_biased(e, stream, 2, gen_ray_dir(cam, x, y, W, H), &c->gbuf[p], out, rk, &rc);   /* :170 */
            memcpy(cur, out, sizeof(orc_sub) * (size_t)N);                                /* :171 */
        }
        snapshot(c, 1);
    }

    /* 4. spatialReuse (render_utils.cpp:87-140) */
    if (f->spatialReuse) {
        for (uint32_t pass = 0; pass < f->spatialResamplingPasses; pass++) {
            /* prevIteration = reservoirGrid (:95,138): neighbours AND self are read from the state before the pass */
            memcpy(c->tmp, c->cur, sizeof(orc_sub) * (size_t)W * H * N);
            #pragma omp parallel for schedule(guided)
            for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
                size_t p = (size_t)y * W + x;
                const orc_hit* hc = &c->gbuf[p];
                romis_stream_key ek = romis_rng_stream(rng->seed, rng->frame, ROMIS_STAGE_SPATIAL0 + pass, (uint32_t)p, ROMIS_STREAM_ENGINE);
                romis_stream_key rk = romis_rng_stream(rng->seed, rng->frame, ROMIS_STAGE_SPATIAL0 + pass, (uint32_t)p, ROMIS_STREAM_RAND);
                const orc_sub* stream[65]; v3 sdir[65]; const orc_hit* sh[65]; int ns = 0;
                for (int nb = 0; nb < k; nb++) {                                          /* :108 */
                    int dx = romis_rng_uniform_int(romis_rng_bits(ek, 2u * nb), -r, r);   /* :109 x first */
                    int dy = romis_rng_uniform_int(romis_rng_bits(ek, 2u * nb + 1u), -r, r);   /* :110 */
                    int nx = x + dx; nx = nx < 0 ? 0 : (nx > W - 1 ? W - 1 : nx);
                    int ny = y + dy; ny = ny < 0 ? 0 : (ny > H - 1 ? H - 1 : ny);
                    size_t q = (size_t)ny * W + nx;
                    const orc_hit* hn = &c->gbuf[q];
                    if (!f->unbiasedCombination) {                                        /* :114-118 */
                        float depthFracDiff = fabsf(1.0f - (hn->t / hc->t));
                        float normalsDot = dot3(hn->n, hc->n);
                        if (depthFracDiff > 0.1f || normalsDot < 0.90630778703f) continue;
                    }
                    stream[ns] = &c->tmp[q * N]; sdir[ns] = gen_ray_dir(cam, nx, ny, W, H); sh[ns] = hn; ns++;   /* :120 */
                }
                v3 dir = gen_ray_dir(cam, x, y, W, H);
                stream[ns] = &c->tmp[p * N]; sdir[ns] = dir; sh[ns] = hc; ns++;           /* :124 self last */
                orc_sub out[ORC_MAX_N]; uint32_t rc = 0;
                if (f->unbiasedCombination) combine_unbiased(e, stream, sdir, sh, ns, dir, hc, out, rk, &rc);   /* :130 */
                else combine_biased(e, stream, ns, dir, hc, out, rk, &rc);                /* :131 */
                memcpy(&c->cur[p * N], out, sizeof(orc_sub) * (size_t)N);                 /* :132 */
            }
            if (pass < 8) snapshot(c, 2 + (int)pass);
        }
    }
    snapshot(c, ORC_MAX_STAGES - 1);

    /* 5. final shading loop (render.cpp:45-57): finalShading (render_utils.cpp:54-65), exposureToneMapping
     *    (src/post_processing/tone_mapping.cpp:8-11), Screen::setPixel (src/rendering/screen.cpp:37-43) */
    if (out_rgb) {
        #pragma omp parallel for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            size_t p = (size_t)y * W + x;
            const orc_hit* h = &c->gbuf[p]; const orc_sub* rs = &c->cur[p * N];
            v3 dir = gen_ray_dir(cam, x, y, W, H);
            v3 color = V3(0, 0, 0);
            for (int j = 0; j < N; j++) {
                v3 sc = visible(e, rs[j].pos, dir, h) ? compute_shading(e, rs[j].pos, rs[j].col, dir, h) : V3(0, 0, 0);
                sc = scale3(sc, rs[j].W);
                color = add3(color, sc);
            }
            color = div3(color, (float)N);
            if (f->enableToneMapping) {
                v3 mapped = V3(1.0f - romis_expf(f->exposure * -color.x), 1.0f - romis_expf(f->exposure * -color.y), 1.0f - romis_expf(f->exposure * -color.z));
                float ig = 1.0f / f->gamma;
                color = V3(romis_powf(mapped.x, ig), romis_powf(mapped.y, ig), romis_powf(mapped.z, ig));
            }
            size_t i = (size_t)(H - 1 - y) * W + x;
            out_rgb[3 * i] = color.x; out_rgb[3 * i + 1] = color.y; out_rgb[3 * i + 2] = color.z;
        }
    }

    /* the returned grid becomes next frame's previousFrameGrid (main.cpp:165) */
    orc_sub* t = c->prev; c->prev = c->cur; c->cur = t;
    c->have_prev = 1;
    return 0;
}


/* =============================================================================================== */
/* R-MIS (renderRMIS, src/rendering/render.cpp:64-119)                                              */
/* =============================================================================================== */
#define ORC_MAX_K 32

/* libstdc++ 13 uniform_int_distribution on a 32-bit engine: Lemire's method WITH its rejection step
 * (/usr/include/c++/13/bits/uniform_int_dist.h:255-281), as std::sample instantiates it inside the reference
 * (neighbour_selection.cpp:79-103).  Returns a value in [0, range). */
static uint32_t lemire32(romis_stream_key ek, uint32_t* ec, uint32_t range) {
    uint64_t product = (uint64_t)romis_rng_bits(ek, (*ec)++) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        uint32_t threshold = (0u - range) % range;
        while (low < threshold) { product = (uint64_t)romis_rng_bits(ek, (*ec)++) * (uint64_t)range; low = (uint32_t)product; }
    }
    return (uint32_t)(product >> 32);
}

/* std::sample = libstdc++ selection sampling (/usr/include/c++/13/bits/stl_algo.h:5841-5905) of n out of list[0..size),
 * order preserving; _Size is ptrdiff_t, two decisions per engine call while unsampled^2 fits the engine range. */
static int std_sample(const int* list, int64_t size, int64_t n, romis_stream_key ek, uint32_t* ec, int* out) {
    int no = 0; int64_t first = 0, unsampled = size;
    if (size == 0) return 0;
    if (n > unsampled) n = unsampled;
    if (0xffffffffull / (uint64_t)unsampled >= (uint64_t)unsampled) {
        while (n != 0 && unsampled >= 2) {
            int64_t b1 = unsampled - 1;
            int64_t x = (int64_t)lemire32(ek, ec, (uint32_t)(unsampled * b1));        /* __gen_two_uniform_ints (:3717-3725) */
            int64_t p0 = x / b1, p1 = x % b1;
            --unsampled;
            if (p0 < n) { out[no++] = list[first]; --n; }
            ++first;
            if (n == 0) break;
            --unsampled;
            if (p1 < n) { out[no++] = list[first]; --n; }
            ++first;
        }
    }
    for (; n != 0; ++first) {
        --unsampled;
        if ((int64_t)lemire32(ek, ec, (uint32_t)(unsampled + 1)) < n) { out[no++] = list[first]; --n; }
    }
    return no;
}

/* areSimilar (src/rendering/neighbour_selection.cpp:7-22); lhs = the canonical pixel.  A miss pixel keeps the
 * value-initialised geometryId 0.  The normal test compares the dot product with the ANGLE in radians (:18), as written. */
static int are_similar(const orc_ctx* c, const romis_rmis_params* rp, const orc_hit* l, const orc_hit* r) {
    uint32_t gl = l->mesh == (uint32_t)c->nmesh ? 0u : l->mesh, gr = r->mesh == (uint32_t)c->nmesh ? 0u : r->mesh;
    if (rp->neighbourSameGeometry && gl != gr) return 0;
    float depthFracDiff = fabsf(1.0f - (l->t / r->t));
    if (depthFracDiff > rp->neighbourMaxDepthDifferenceFraction) return 0;
    float normalsDot = dot3(l->n, r->n);
    if (normalsDot < rp->neighbourMaxNormalAngleDifferenceRadians) return 0;
    return 1;
}

/* indicesRandom / indicesSimilarity (neighbour_selection.cpp:24-105): out[0] = the pixel itself; returns the count */
static int rmis_indices(const orc_ctx* c, const romis_features* f, const romis_rmis_params* rp, const romis_rng* rng,
                        int x, int y, int W, int H, int* sim, int* dis, int* out) {
    const int k = (int)f->numNeighboursToSample, r = (int)f->spatialResampleRadius;
    romis_stream_key ek = romis_rng_stream(rng->seed, rng->frame, ROMIS_STAGE_RMIS_NEIGH, (uint32_t)(y * W + x), ROMIS_STREAM_ENGINE);
    uint32_t ec = 0; int no = 0;
    int x0 = x - r < 0 ? 0 : x - r, x1 = x + r > W - 1 ? W - 1 : x + r;
    int y0 = y - r < 0 ? 0 : y - r, y1 = y + r > H - 1 ? H - 1 : y + r;
    out[no++] = y * W + x;                                                              /* :40 / :71 */
    if (rp->neighbourSelectionStrategy == ROMIS_NEIGHBOURS_RANDOM) {                    /* :24-45 */
        for (int i = 0; i < k; i++) {
            /* `glm::ivec2(distrX(gen), distrY(gen))` (:42): the order of the two draws is unspecified in C++; g++, which
             * builds the compiled reference this oracle is pinned against, evaluates the arguments right to left: y first */
            int ny = romis_rng_uniform_int(romis_rng_bits(ek, ec++), y0, y1);
            int nx = romis_rng_uniform_int(romis_rng_bits(ek, ec++), x0, x1);
            out[no++] = ny * W + nx;
        }
        return no;
    }
    int ns = 0, nd = 0;
    const orc_hit* canon = &c->gbuf[(size_t)y * W + x];
    for (int ny = y0; ny <= y1; ny++) for (int nx = x0; nx <= x1; nx++) {               /* :59-72 */
        if (ny == y && nx == x) continue;
        if (are_similar(c, rp, canon, &c->gbuf[(size_t)ny * W + nx])) sim[ns++] = ny * W + nx; else dis[nd++] = ny * W + nx;
    }
    if (rp->neighbourSelectionStrategy == ROMIS_NEIGHBOURS_SIMILAR) {                   /* :79-85 */
        if (ns < k) {
            for (int i = 0; i < ns; i++) out[no++] = sim[i];
            no += std_sample(dis, nd, (int64_t)k - ns, ek, &ec, out + no);
        } else no += std_sample(sim, ns, k, ek, &ec, out + no);
    } else {                                                                            /* EqualSimilarDissimilar :94-103 */
        uint32_t similarsSampled = (uint32_t)k / 2u + 1u; if ((uint64_t)similarsSampled > (uint64_t)ns) similarsSampled = (uint32_t)ns;
        uint32_t desiredDissimilars = (uint32_t)k - similarsSampled;
        if ((uint64_t)desiredDissimilars > (uint64_t)nd) similarsSampled += (uint32_t)((uint64_t)(uint32_t)k - (uint64_t)nd - (uint64_t)similarsSampled);
        no += std_sample(sim, ns, (int64_t)similarsSampled, ek, &ec, out + no);
        no += std_sample(dis, nd, (int64_t)(uint32_t)((uint32_t)k - similarsSampled), ek, &ec, out + no);
    }
    return no;
}

int orc_render_frame_rmis(orc_ctx* c, const romis_features* f, const romis_rmis_params* rp, const romis_camera* cam, int W, int H,
                          const romis_rng* rng, float* out_rgb, int32_t* neigh_xy, uint32_t* neigh_count) {
    if (!c->tracer) { strcpy(c->err, "no scene"); return ROMIS_ERR_STATE; }
    const int N = (int)f->numSamplesInReservoir, k = (int)f->numNeighboursToSample, r = (int)f->spatialResampleRadius;
    if (N < 1 || N > ORC_MAX_N || W < 1 || H < 1 || k > ORC_MAX_K) { strcpy(c->err, "bad size"); return ROMIS_ERR_INVALID; }
    if (rp->neighbourSelectionStrategy == ROMIS_NEIGHBOURS_DISSIMILAR) { strcpy(c->err, "Dissimilar: undefined in the reference"); return ROMIS_ERR_INVALID; }
    if (c->W != W || c->H != H || c->N != N) {
        size_t n = (size_t)W * H;
        c->gbuf = (orc_hit*)realloc(c->gbuf, n * sizeof(orc_hit));
        c->cur = (orc_sub*)realloc(c->cur, n * N * sizeof(orc_sub));
        c->prev = (orc_sub*)realloc(c->prev, n * N * sizeof(orc_sub));
        c->tmp = (orc_sub*)realloc(c->tmp, n * N * sizeof(orc_sub));
        c->W = W; c->H = H; c->N = N;
    }
    c->have_prev = 0;
    orc_env env; env.f = f; env.c = c; env.origin = lv(cam->origin);
    const orc_env* e = &env;
    const int K1 = k + 1;
    primary_hits(c, cam, env.origin, W, H);                                             /* render.cpp:68 */
    /* generateResampleIndicesGrid (neighbour_selection.cpp:107-122) */
    int* idx = (int*)malloc(sizeof(int) * (size_t)W * H * K1);
    int* cnt = (int*)malloc(sizeof(int) * (size_t)W * H);
    const int win = (2 * r + 1) * (2 * r + 1) + 1;
    #pragma omp parallel
    {
        int* sim = (int*)malloc(sizeof(int) * (size_t)win); int* dis = (int*)malloc(sizeof(int) * (size_t)win);
        #pragma omp for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            size_t p = (size_t)y * W + x;
            int tmp[ORC_MAX_K + 2];
            int n = rmis_indices(c, f, rp, rng, x, y, W, H, sim, dis, tmp);
            cnt[p] = n;
            for (int i = 0; i < K1; i++) idx[p * K1 + i] = i < n ? tmp[i] : -1;
        }
        free(sim); free(dis);
    }
    if (neigh_count) for (size_t p = 0; p < (size_t)W * H; p++) neigh_count[p] = (uint32_t)cnt[p];
    if (neigh_xy) for (size_t i = 0; i < (size_t)W * H * K1; i++) {
        neigh_xy[2 * i] = idx[i] < 0 ? -1 : idx[i] % W; neigh_xy[2 * i + 1] = idx[i] < 0 ? -1 : idx[i] / W;
    }
    v3* acc = (v3*)calloc((size_t)W * H, sizeof(v3));
    for (uint32_t it = 0; it < rp->maxIterationsMIS; it++) {                            /* render.cpp:72 */
        #pragma omp parallel for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {                       /* genInitialSamples :74 */
            size_t p = (size_t)y * W + x;
            gen_canonical(e, rng, ROMIS_STAGE_RMIS_INITIAL0 + it, (uint32_t)p, gen_ray_dir(cam, x, y, W, H), &c->gbuf[p], &c->cur[p * N]);
        }
        #pragma omp parallel for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {                       /* :79-112 */
            size_t p = (size_t)y * W + x;
            const orc_hit* h = &c->gbuf[p]; v3 dir = gen_ray_dir(cam, x, y, W, H);
            const int n = cnt[p];
            v3 finalColor = V3(0, 0, 0);
            for (int a = 0; a < n; a++) {
                const orc_sub* px = &c->cur[(size_t)idx[p * K1 + a] * N];
                for (int j = 0; j < N; j++) {
                    float misWeight;
                    if (rp->misWeightRMIS == ROMIS_MIS_EQUAL) misWeight = 1.0f / (float)n;                         /* :97 */
                    else {                                                              /* generalisedBalanceHeuristic, render_utils.cpp:179-187 */
                        float numerator = target_pdf(e, px[j].pos, px[j].col, dir, h);
                        float denominator = FLT_MIN;
                        for (int b = 0; b < n; b++) {
                            int q = idx[p * K1 + b];
                            denominator += target_pdf(e, px[j].pos, px[j].col, gen_ray_dir(cam, q % W, q / W, W, H), &c->gbuf[q]);
                        }
                        misWeight = numerator / denominator;
                    }
                    v3 sampleColor = visible(e, px[j].pos, dir, h) ? compute_shading(e, px[j].pos, px[j].col, dir, h) : V3(0, 0, 0);   /* :103-105 */
                    finalColor = add3(finalColor, div3(scale3(scale3(sampleColor, misWeight), px[j].W), (float)N));                  /* :106 */
                }
            }
            acc[p] = add3(acc[p], finalColor);                                          /* :111 */
        }
    }
    if (out_rgb) {                                                                      /* combineToScreen, render_utils.cpp:68-85 */
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            v3 color = div3(acc[(size_t)y * W + x], (float)rp->maxIterationsMIS);
            if (f->enableToneMapping) {
                v3 mapped = V3(1.0f - romis_expf(f->exposure * -color.x), 1.0f - romis_expf(f->exposure * -color.y), 1.0f - romis_expf(f->exposure * -color.z));
                float ig = 1.0f / f->gamma;
                color = V3(romis_powf(mapped.x, ig), romis_powf(mapped.y, ig), romis_powf(mapped.z, ig));
            }
            size_t i = (size_t)(H - 1 - y) * W + x;
            out_rgb[3 * i] = color.x; out_rgb[3 * i + 1] = color.y; out_rgb[3 * i + 2] = color.z;
        }
    }
    free(idx); free(cnt); free(acc);
    return 0;
}

/* =============================================================================================== */
/* R-OMIS (renderROMIS, src/rendering/render.cpp:121-265)                                           */
/* =============================================================================================== */
/* arbitraryUnbiasedContributionWeightReciprocal (src/rendering/render_utils.cpp:245-257) */
static float aucw_reciprocal(const orc_env* e, v3 pos, v3 col, v3 dir, const orc_hit* h, const orc_sub* pixel, int sampleIdx) {
    float targetPdfValue = target_pdf(e, pos, col, dir, h);
    if (targetPdfValue == 0.0f) return 0.0f;
    float mockSampleWeight = targetPdfValue / (1.0f / (float)e->c->nlights);
    float arbitraryWeight = (1.0f / targetPdfValue) * (1.0f / (float)pixel[sampleIdx].M) *
                            (pixel[sampleIdx].wSum - pixel[sampleIdx].chosen + mockSampleWeight);
    return 1.0f / arbitraryWeight;
}

/* matrices: [H][W][K1][K1], contributions: [H][W][3][K1] after the last iteration (either may be NULL) */
int orc_render_frame_romis(orc_ctx* c, const romis_features* f, const romis_rmis_params* rp, const romis_camera* cam, int W, int H,
                           const romis_rng* rng, float* out_rgb, float* matrices, float* contributions) {
    if (!c->tracer) { strcpy(c->err, "no scene"); return ROMIS_ERR_STATE; }
    const int N = (int)f->numSamplesInReservoir, k = (int)f->numNeighboursToSample, r = (int)f->spatialResampleRadius;
    const int K1 = k + 1;
    if (N < 1 || N > ORC_MAX_N || W < 1 || H < 1 || K1 > ROMIS_COD_MAX) { strcpy(c->err, "bad size"); return ROMIS_ERR_INVALID; }
    if (rp->neighbourSelectionStrategy == ROMIS_NEIGHBOURS_DISSIMILAR) { strcpy(c->err, "Dissimilar: undefined in the reference"); return ROMIS_ERR_INVALID; }
    if (rp->useProgressiveROMIS && rp->progressiveUpdateMod == 0) { strcpy(c->err, "progressiveUpdateMod == 0: modulo by zero in the reference (render.cpp:160)"); return ROMIS_ERR_INVALID; }
    /* renderROMIS indexes neighborhood[0 .. k] whatever its size (render.cpp:165,173): every pixel needs k other pixels in
     * its window, or the reference reads unconstructed Reservoirs */
    if (rp->neighbourSelectionStrategy != ROMIS_NEIGHBOURS_RANDOM) {
        long wx = (r + 1 < W ? r + 1 : W), wy = (r + 1 < H ? r + 1 : H);
        if (wx * wy - 1 < k) { strcpy(c->err, "window smaller than numNeighboursToSample: undefined in the reference"); return ROMIS_ERR_INVALID; }
    }
    if (c->W != W || c->H != H || c->N != N) {
        size_t n = (size_t)W * H;
        c->gbuf = (orc_hit*)realloc(c->gbuf, n * sizeof(orc_hit));
        c->cur = (orc_sub*)realloc(c->cur, n * N * sizeof(orc_sub));
        c->prev = (orc_sub*)realloc(c->prev, n * N * sizeof(orc_sub));
        c->tmp = (orc_sub*)realloc(c->tmp, n * N * sizeof(orc_sub));
        c->W = W; c->H = H; c->N = N;
    }
    c->have_prev = 0;
    orc_env env; env.f = f; env.c = c; env.origin = lv(cam->origin);
    const orc_env* e = &env;
    primary_hits(c, cam, env.origin, W, H);                                             /* render.cpp:125 */
    int* idx = (int*)malloc(sizeof(int) * (size_t)W * H * K1);
    const int win = (2 * r + 1) * (2 * r + 1) + 1;
    int bad = 0;
    #pragma omp parallel
    {
        int* sim = (int*)malloc(sizeof(int) * (size_t)win); int* dis = (int*)malloc(sizeof(int) * (size_t)win);
        #pragma omp for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {                       /* :126 */
            size_t p = (size_t)y * W + x;
            int tmp[ORC_MAX_K + 2];
            int n = rmis_indices(c, f, rp, rng, x, y, W, H, sim, dis, tmp);
            if (n != K1) { bad = 1; n = n < K1 ? n : K1; }
            for (int i = 0; i < K1; i++) idx[p * K1 + i] = i < n ? tmp[i] : 0;
        }
        free(sim); free(dis);
    }
    if (bad) { free(idx); strcpy(c->err, "a pixel has fewer than k neighbours"); return ROMIS_ERR_INVALID; }
    float* A = (float*)calloc((size_t)W * H * K1 * K1, sizeof(float));                  /* techniqueMatrices :128 */
    float* B = (float*)calloc((size_t)W * H * 3 * K1, sizeof(float));                   /* contributionVectors{Red,Green,Blue} :129-131 */
    /* progressive estimator only (:133-139) */
    const int progressive = rp->useProgressiveROMIS != 0;
    float* alpha = (float*)calloc((size_t)W * H * 3 * K1, sizeof(float));               /* alphaVectors{Red,Green,Blue} */
    v3* fin = (v3*)calloc((size_t)W * H, sizeof(v3));                                   /* finalPixelColors */
    const int32_t totalSamples = (int32_t)((uint32_t)K1 * f->numSamplesInReservoir);
    const int32_t fractionOfTotalSamples = (int32_t)(f->numSamplesInReservoir / (uint32_t)K1);     /* integer division, as written (:139) */
    for (uint32_t it = 0; it < rp->maxIterationsMIS; it++) {                            /* :141 */
        #pragma omp parallel for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {                       /* genInitialSamples :143 */
            size_t p = (size_t)y * W + x;
            gen_canonical(e, rng, ROMIS_STAGE_RMIS_INITIAL0 + it, (uint32_t)p, gen_ray_dir(cam, x, y, W, H), &c->gbuf[p], &c->cur[p * N]);
        }
        #pragma omp parallel for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {                       /* :148-222 */
            size_t p = (size_t)y * W + x;
            const orc_hit* h = &c->gbuf[p]; v3 dir = gen_ray_dir(cam, x, y, W, H);
            float* Ap = A + p * K1 * K1; float* Bp = B + p * 3 * K1; float* Al = alpha + p * 3 * K1;
            if (progressive && it >= 1u && it % rp->progressiveUpdateMod == 0u) {       /* :160-164 alpha estimates from what has been gathered */
                romis_cod cod; romis_cod_compute(&cod, Ap, K1);
                for (int ch = 0; ch < 3; ch++) romis_cod_solve(&cod, Bp + ch * K1, Al + ch * K1);
            }
            for (int a = 0; a < K1; a++) {                                              /* :165 */
                if (progressive) fin[p] = add3(fin[p], V3(Al[0 * K1 + a], Al[1 * K1 + a], Al[2 * K1 + a]));     /* :168-170 */
                const orc_sub* px = &c->cur[(size_t)idx[p * K1 + a] * N];
                for (int j = 0; j < N; j++) {                                           /* :173 */
                    float colVecW[ROMIS_COD_MAX];
                    for (int b = 0; b < K1; b++) {                                      /* :178-181 */
                        int q = idx[p * K1 + b];
                        colVecW[b] = aucw_reciprocal(e, px[j].pos, px[j].col, gen_ray_dir(cam, q % W, q / W, W, H), &c->gbuf[q],
                                                     &c->cur[(size_t)q * N], j);
                    }
                    v3 sampleColor = visible(e, px[j].pos, dir, h) ? compute_shading(e, px[j].pos, px[j].col, dir, h) : V3(0, 0, 0);   /* :184-186 */
                    if (progressive) {                                                  /* :190-200 */
                        v3 sumAlphaProducts = V3(0, 0, 0); float sumSampleFractionProducts = FLT_MIN;
                        for (int b = 0; b < K1; b++) {
                            sumAlphaProducts = add3(sumAlphaProducts, scale3(V3(Al[0 * K1 + b], Al[1 * K1 + b], Al[2 * K1 + b]), colVecW[b]));
                            sumSampleFractionProducts += (float)fractionOfTotalSamples * colVecW[b];
                        }
                        fin[p] = add3(fin[p], scale3(sub3(div3(sampleColor, sumSampleFractionProducts), div3(sumAlphaProducts, sumSampleFractionProducts)),
                                                     1.0f / (float)totalSamples));
                    }
                    float scaleFactor = FLT_MIN;                                        /* :203-205 */
                    for (int b = 0; b < K1; b++) scaleFactor += (float)f->numSamplesInReservoir * colVecW[b];
                    scaleFactor = 1.0f / scaleFactor;
                    for (int b = 0; b < K1; b++) colVecW[b] *= scaleFactor;             /* :208 */
                    for (int i = 0; i < K1; i++) for (int b = 0; b < K1; b++) Ap[i * K1 + b] += colVecW[i] * colVecW[b];   /* :209 */
                    for (int row = 0; row < K1; row++) {                                /* :210-215 */
                        float scaleColVecConst = scaleFactor * colVecW[row];
                        Bp[0 * K1 + row] += sampleColor.x * scaleColVecConst;
                        Bp[1 * K1 + row] += sampleColor.y * scaleColVecConst;
                        Bp[2 * K1 + row] += sampleColor.z * scaleColVecConst;
                    }
                }
            }
        }
    }
    if (matrices) memcpy(matrices, A, sizeof(float) * (size_t)W * H * K1 * K1);
    if (contributions) memcpy(contributions, B, sizeof(float) * (size_t)W * H * 3 * K1);
    if (out_rgb) {                                                                      /* :233-262 */
        #pragma omp parallel for schedule(guided)
        for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            size_t p = (size_t)y * W + x;
            v3 color = V3(0, 0, 0);
            if (progressive) color = div3(fin[p], (float)rp->maxIterationsMIS);         /* combineToScreen (:232, render_utils.cpp:68-85) */
            else {
                romis_cod cod; float xs[3][ROMIS_COD_MAX];
                romis_cod_compute(&cod, A + p * K1 * K1, K1);                           /* solveSystem, render_utils.h:52 */
                for (int ch = 0; ch < 3; ch++) romis_cod_solve(&cod, B + (p * 3 + ch) * K1, xs[ch]);
                for (int row = 0; row < K1; row++) { color.x += xs[0][row]; color.y += xs[1][row]; color.z += xs[2][row]; }   /* :247-252 */
            }
            if (f->enableToneMapping) {
                v3 mapped = V3(1.0f - romis_expf(f->exposure * -color.x), 1.0f - romis_expf(f->exposure * -color.y), 1.0f - romis_expf(f->exposure * -color.z));
                float ig = 1.0f / f->gamma;
                color = V3(romis_powf(mapped.x, ig), romis_powf(mapped.y, ig), romis_powf(mapped.z, ig));
            }
            size_t i = (size_t)(H - 1 - y) * W + x;
            out_rgb[3 * i] = color.x; out_rgb[3 * i + 1] = color.y; out_rgb[3 * i + 2] = color.z;
        }
    }
    free(idx); free(A); free(B); free(alpha); free(fin);
    return 0;
}

/* exported for the solver tests: one system */
int orc_cod_solve(const float* A, const float* b, int n, float* x, int* rank) {
    if (n < 1 || n > ROMIS_COD_MAX) return ROMIS_ERR_INVALID;
    romis_cod cod; romis_cod_compute(&cod, A, n); romis_cod_solve(&cod, b, x);
    if (rank) *rank = cod.rank;
    return 0;
}

static const orc_sub* stage_ptr(orc_ctx* c, int pass_id) {
    int slot;
    if (pass_id == ROMIS_PASS_FINAL) slot = ORC_MAX_STAGES - 1;
    else if (pass_id >= 0 && pass_id < ORC_MAX_STAGES - 1) slot = pass_id;
    else return NULL;
    return c->stage_valid[slot] ? c->stage_dump[slot] : NULL;
}

/* dump layout = romis_reservoir_dump plus the two running values the reference keeps (wSums, chosenSampleWeights) */
int orc_download_reservoirs(orc_ctx* c, int pass_id, romis_reservoir_dump* out, float* wSum, float* chosen) {
    const orc_sub* s = stage_ptr(c, pass_id);
    if (!s) { strcpy(c->err, "stage not captured"); return ROMIS_ERR_STATE; }
    const int W = c->W, H = c->H, N = c->N;
    for (int j = 0; j < N; j++) for (size_t p = 0; p < (size_t)W * H; p++) {
        const orc_sub* r = &s[p * N + j]; size_t i = (size_t)j * W * H + p;
        if (out->light_id) out->light_id[i] = r->light;
        if (out->u) out->u[i] = r->u;
        if (out->v) out->v[i] = r->v;
        if (out->W) out->W[i] = r->W;
        if (out->M) out->M[i] = (uint32_t)r->M;
        if (out->position) { out->position[3 * i] = r->pos.x; out->position[3 * i + 1] = r->pos.y; out->position[3 * i + 2] = r->pos.z; }
        if (out->color) { out->color[3 * i] = r->col.x; out->color[3 * i + 1] = r->col.y; out->color[3 * i + 2] = r->col.z; }
        if (wSum) wSum[i] = r->wSum;
        if (chosen) chosen[i] = r->chosen;
    }
    return 0;
}

int orc_download_gbuffer(orc_ctx* c, romis_gbuffer_dump* out) {
    if (!c->gbuf) { strcpy(c->err, "no frame"); return ROMIS_ERR_STATE; }
    for (size_t p = 0; p < (size_t)c->W * c->H; p++) {
        const orc_hit* h = &c->gbuf[p];
        if (out->t) out->t[p] = h->t;
        if (out->normal) { out->normal[3 * p] = h->n.x; out->normal[3 * p + 1] = h->n.y; out->normal[3 * p + 2] = h->n.z; }
        if (out->texcoord) { out->texcoord[2 * p] = h->uv[0]; out->texcoord[2 * p + 1] = h->uv[1]; }
        if (out->mesh) out->mesh[p] = h->mesh;
    }
    return 0;
}

int orc_ray_dirs(const romis_camera* cam, int W, int H, float* dirs) {
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        v3 d = gen_ray_dir(cam, x, y, W, H); size_t p = (size_t)y * W + x;
        dirs[3 * p] = d.x; dirs[3 * p + 1] = d.y; dirs[3 * p + 2] = d.z;
    }
    return 0;
}

int orc_trace_rays(orc_ctx* c, const float* origins, const float* dirs, const float* tfar, int n, int any_hit,
                   uint8_t* hit, float* t, float* u, float* v, uint32_t* tri) {
    if (!c->tracer) { strcpy(c->err, "no scene"); return ROMIS_ERR_STATE; }
    #pragma omp parallel for schedule(guided)
    for (int i = 0; i < n; i++) {
        if (any_hit) hit[i] = (uint8_t)otr_any(c->tracer, origins + 3 * i, dirs + 3 * i, tfar[i]);
        else {
            float tt, uu, vv; uint32_t ti;
            hit[i] = (uint8_t)otr_closest(c->tracer, origins + 3 * i, dirs + 3 * i, tfar[i], &tt, &uu, &vv, &ti);
            if (hit[i]) { if (t) t[i] = tt; if (u) u[i] = uu; if (v) v[i] = vv; if (tri) tri[i] = ti; }
        }
    }
    return 0;
}

/* exported for tests/test_detmath.py */
float orc_powf(float x, float y) { return romis_powf(x, y); }
float orc_expf(float x) { return romis_expf(x); }
uint32_t orc_rng_bits(uint64_t seed, uint32_t frame, uint32_t stage, uint32_t pixel, uint32_t stream, uint32_t counter) {
    return romis_rng_bits(romis_rng_stream(seed, frame, stage, pixel, stream), counter);
}
