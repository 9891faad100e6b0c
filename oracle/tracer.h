/*
 * oracle/tracer.h -- CPU ray/triangle tracer of the TEST ORACLE (test infrastructure, not product).
 *
 * The reference delegates intersection to Intel Embree 4.3.1 (reference CMakeLists.txt:20-21;
 * call sites src/ray_tracing/embree_interface.cpp:60,67,76-81), an un-vendored dependency that is
 * not in this image and whose results no reference test pins.  This scalar tracer
 * therefore DEFINES the intersection arithmetic for both the oracle and the CUDA path:
 *
 *   per triangle (v0, e1 = v1 - v0, e2 = v2 - v0), ray (o, d), range [0, tfar]:
 *     p = cross(d, e2); det = dot(e1, p); det == 0 -> miss; inv = 1 / det
 *     s = o - v0; u = dot(s, p) * inv; u < 0 || u > 1 -> miss
 *     q = cross(s, e1); v = dot(d, q) * inv; v < 0 || u + v > 1 -> miss
 *     t = dot(e2, q) * inv; hit iff 0 <= t <= tfar
 *   dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z, cross as GLM (func_geometric.inl:68-79), no FMA.
 *   closest hit = smallest t; equal t -> smallest global triangle index (mesh order, then the
 *   mesh's triangle order).  Any hit = some triangle hits within [0, tfar].
 *
 * With these rules the answer is independent of traversal order, so the BVH here, the brute-force
 * mode and the GPU's own BVH must agree bit for bit.
 */
#ifndef ORACLE_TRACER_H
#define ORACLE_TRACER_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct otr_tracer otr_tracer;

/* verts: ntri * 9 floats (v0 v1 v2); mode 0 = brute force, 1 = BVH */
otr_tracer* otr_build(const float* verts, int ntri, int mode);
void otr_free(otr_tracer* t);
/* returns 1 on hit and fills t,u,v,tri */
int otr_closest(const otr_tracer* tr, const float o[3], const float d[3], float tfar,
                float* t, float* u, float* v, uint32_t* tri);
int otr_any(const otr_tracer* tr, const float o[3], const float d[3], float tfar);

#ifdef __cplusplus
}
#endif
#endif
