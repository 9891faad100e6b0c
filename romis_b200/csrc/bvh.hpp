// bvh.hpp -- host-side BVH builder of libromis_gpu.so.
//
// Replaces what the reference gets from Intel Embree (rtcCommitScene with RTC_BUILD_QUALITY_HIGH,
// reference src/ray_tracing/embree_interface.cpp:30-51): a binned-SAH BVH2 over all meshes of the
// scene, flattened into 64-byte nodes that carry BOTH children's boxes (one 64-B node fetch decides
// both descents), plus triangles re-ordered into leaf order as three float4 (v0, e1, e2, index).
//
// Result contract (see oracle/tracer.h for the per-triangle arithmetic): closest hit = smallest t,
// ties -> smallest global triangle index; boxes are padded so that traversal never culls a triangle
// the per-triangle test would accept, which makes the answer independent of tree shape.
#pragma once
#include <cstdint>
#include <vector>

namespace romis {

struct alignas(64) BvhNode {
    float lo0[3], hi0[3];   // child 0 box
    float lo1[3], hi1[3];   // child 1 box
    int32_t child0, child1; // node index, or first triangle (leaf order) when count > 0
    int32_t count0, count1; // 0 = inner node, > 0 = leaf with that many triangles, -1 = empty slot
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be one 64-byte line half");

struct alignas(16) TriGeom {   // 48 B, leaf order
    float v0[3]; float e1x;
    float e1y, e1z, e2x, e2y;
    float e2z; uint32_t tri; uint32_t pad0, pad1;   // tri = global triangle index (mesh order)
};
static_assert(sizeof(TriGeom) == 48, "TriGeom is three float4");

struct Bvh {
    std::vector<BvhNode> nodes;     // nodes[0] = root
    std::vector<TriGeom> tris;      // leaf order
    int max_depth = 0;
};

// verts: 9 floats per triangle (v0, v1, v2) in global triangle order.
Bvh build_bvh(const float* verts, int ntri);

}  // namespace romis
