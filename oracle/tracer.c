/* oracle/tracer.c -- see tracer.h.  TEST ORACLE ONLY: never linked into the product library. */
#include "tracer.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

typedef struct { float v0[3], e1[3], e2[3]; } otr_tri;
typedef struct { float lo[3], hi[3]; int left, right; int first, count; } otr_node;

struct otr_tracer {
    int ntri, mode;
    otr_tri* tris;
    int* order;      /* BVH leaf order -> global triangle index */
    otr_node* nodes;
    int nnodes;
    float pad;
};

static inline float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
static inline void cross3(const float a[3], const float b[3], float r[3]) {
    r[0] = a[1] * b[2] - b[1] * a[2];
    r[1] = a[2] * b[0] - b[2] * a[0];
    r[2] = a[0] * b[1] - b[0] * a[1];
}

static inline int tri_hit(const otr_tri* T, const float o[3], const float d[3], float tfar,
                          float* t, float* u, float* v) {
    float p[3], s[3], q[3];
    cross3(d, T->e2, p);
    float det = dot3(T->e1, p);
    if (det == 0.0f) return 0;
    float inv = 1.0f / det;
    s[0] = o[0] - T->v0[0]; s[1] = o[1] - T->v0[1]; s[2] = o[2] - T->v0[2];
    float uu = dot3(s, p) * inv;
    if (!(uu >= 0.0f) || uu > 1.0f) return 0;
    cross3(s, T->e1, q);
    float vv = dot3(d, q) * inv;
    if (!(vv >= 0.0f) || uu + vv > 1.0f) return 0;
    float tt = dot3(T->e2, q) * inv;
    if (!(tt >= 0.0f) || tt > tfar) return 0;
    *t = tt; *u = uu; *v = vv;
    return 1;
}

/* ---- BVH: median split on the widest centroid axis, leaves <= 4 triangles, padded boxes ---- */
static void tri_bounds(const float* v, float lo[3], float hi[3]) {
    for (int a = 0; a < 3; a++) {
        lo[a] = fminf(v[a], fminf(v[3 + a], v[6 + a]));
        hi[a] = fmaxf(v[a], fmaxf(v[3 + a], v[6 + a]));
    }
}
static const float* g_verts; static int g_axis;
static int cmp_centroid(const void* pa, const void* pb) {
    int a = *(const int*)pa, b = *(const int*)pb;
    float ca = g_verts[a * 9 + g_axis] + g_verts[a * 9 + 3 + g_axis] + g_verts[a * 9 + 6 + g_axis];
    float cb = g_verts[b * 9 + g_axis] + g_verts[b * 9 + 3 + g_axis] + g_verts[b * 9 + 6 + g_axis];
    if (ca < cb) return -1;
    if (ca > cb) return 1;
    return a - b;
}
static int build_rec(otr_tracer* tr, const float* verts, int first, int count) {
    int id = tr->nnodes++;
    otr_node* n = &tr->nodes[id];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = first; i < first + count; i++) {
        float l[3], h[3]; tri_bounds(verts + tr->order[i] * 9, l, h);
        for (int a = 0; a < 3; a++) {
            lo[a] = fminf(lo[a], l[a]); hi[a] = fmaxf(hi[a], h[a]);
            float c = l[a] + h[a]; clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c);
        }
    }
    for (int a = 0; a < 3; a++) { n->lo[a] = lo[a] - tr->pad; n->hi[a] = hi[a] + tr->pad; }
    n->first = first; n->count = count; n->left = n->right = -1;
    if (count > 4) {
        int axis = 0; float ext = chi[0] - clo[0];
        for (int a = 1; a < 3; a++) if (chi[a] - clo[a] > ext) { ext = chi[a] - clo[a]; axis = a; }
        g_verts = verts; g_axis = axis;
        qsort(tr->order + first, (size_t)count, sizeof(int), cmp_centroid);
        int half = count / 2;
        int l = build_rec(tr, verts, first, half);
        int r = build_rec(tr, verts, first + half, count - half);
        n = &tr->nodes[id];
        n->left = l; n->right = r; n->count = 0;
    }
    return id;
}

otr_tracer* otr_build(const float* verts, int ntri, int mode) {
    otr_tracer* tr = (otr_tracer*)calloc(1, sizeof(*tr));
    tr->ntri = ntri; tr->mode = mode;
    tr->tris = (otr_tri*)malloc(sizeof(otr_tri) * (size_t)(ntri > 0 ? ntri : 1));
    float ext = 0.0f;
    for (int i = 0; i < ntri; i++) {
        const float* v = verts + i * 9;
        for (int a = 0; a < 3; a++) {
            tr->tris[i].v0[a] = v[a];
            tr->tris[i].e1[a] = v[3 + a] - v[a];
            tr->tris[i].e2[a] = v[6 + a] - v[a];
            for (int k = 0; k < 3; k++) ext = fmaxf(ext, fabsf(v[3 * k + a]));
        }
    }
    if (mode == 1 && ntri > 0) {
        tr->pad = 2e-5f * ext + 1e-30f;
        tr->order = (int*)malloc(sizeof(int) * (size_t)ntri);
        for (int i = 0; i < ntri; i++) tr->order[i] = i;
        tr->nodes = (otr_node*)malloc(sizeof(otr_node) * (size_t)(2 * ntri));
        tr->nnodes = 0;
        build_rec(tr, verts, 0, ntri);
    }
    return tr;
}

void otr_free(otr_tracer* t) {
    if (!t) return;
    free(t->tris); free(t->order); free(t->nodes); free(t);
}

static inline int box_hit(const otr_node* n, const float o[3], const float inv[3], float tmax) {
    float tn = 0.0f, tf = tmax;
    for (int a = 0; a < 3; a++) {
        float t0 = (n->lo[a] - o[a]) * inv[a], t1 = (n->hi[a] - o[a]) * inv[a];
        tn = fmaxf(tn, fminf(t0, t1));   /* fminf/fmaxf drop NaNs (0 * inf) */
        tf = fminf(tf, fmaxf(t0, t1));
    }
    return tn <= tf * 1.0000004f;
}

int otr_closest(const otr_tracer* tr, const float o[3], const float d[3], float tfar,
                float* t, float* u, float* v, uint32_t* tri) {
    int found = 0; float bt = tfar, bu = 0, bv = 0; uint32_t bi = 0xffffffffu;
    if (tr->mode == 0 || tr->ntri == 0) {
        for (int i = 0; i < tr->ntri; i++) {
            float tt, uu, vv;
            if (tri_hit(&tr->tris[i], o, d, bt, &tt, &uu, &vv)) {
                if (!found || tt < bt) { found = 1; bt = tt; bu = uu; bv = vv; bi = (uint32_t)i; }
            }
        }
    } else {
        float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        int stack[128]; int sp = 0; stack[sp++] = 0;
        while (sp) {
            const otr_node* n = &tr->nodes[stack[--sp]];
            if (!box_hit(n, o, inv, bt)) continue;
            if (n->left < 0) {
                for (int k = n->first; k < n->first + n->count; k++) {
                    int i = tr->order[k]; float tt, uu, vv;
                    if (tri_hit(&tr->tris[i], o, d, bt, &tt, &uu, &vv)) {
                        if (!found || tt < bt || (tt == bt && (uint32_t)i < bi)) {
                            found = 1; bt = tt; bu = uu; bv = vv; bi = (uint32_t)i;
                        }
                    }
                }
            } else { stack[sp++] = n->left; stack[sp++] = n->right; }
        }
    }
    if (found) { *t = bt; *u = bu; *v = bv; *tri = bi; }
    return found;
}

int otr_any(const otr_tracer* tr, const float o[3], const float d[3], float tfar) {
    float tt, uu, vv;
    if (tr->mode == 0 || tr->ntri == 0) {
        for (int i = 0; i < tr->ntri; i++) if (tri_hit(&tr->tris[i], o, d, tfar, &tt, &uu, &vv)) return 1;
        return 0;
    }
    float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const otr_node* n = &tr->nodes[stack[--sp]];
        if (!box_hit(n, o, inv, tfar)) continue;
        if (n->left < 0) {
            for (int k = n->first; k < n->first + n->count; k++)
                if (tri_hit(&tr->tris[tr->order[k]], o, d, tfar, &tt, &uu, &vv)) return 1;
        } else { stack[sp++] = n->left; stack[sp++] = n->right; }
    }
    return 0;
}
