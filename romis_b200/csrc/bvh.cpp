// bvh.cpp -- binned-SAH BVH2 builder (host).  See bvh.hpp.
#include "bvh.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <numeric>

// Triangles per leaf.  Measured on B200 (C2 1080p, tools/quick_bench.py): 2 beats 1, 4 and 8 -- a triangle test costs about
// as much as a box pair, so small leaves win once the rays are incoherent (shadow rays).
#ifndef ROMIS_LEAF_MAX
#define ROMIS_LEAF_MAX 2
#endif

namespace romis {
namespace {

struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const float* p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box& b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const {
        float d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        if (d[0] < 0) return 0.0f;
        return 2.0f * (d[0] * d[1] + d[1] * d[2] + d[2] * d[0]);
    }
};

struct Builder {
    const float* verts;
    std::vector<Box> tbox;
    std::vector<float> cent;        // 3 per triangle
    std::vector<int> order;
    Bvh* out;
    float pad;
    static constexpr int kLeafMax = ROMIS_LEAF_MAX;
    static constexpr int kBins = 16;

    struct Ref { int32_t child; int32_t count; Box box; int depth; };

    // builds the subtree over order[first, first+count) and returns how its parent refers to it
    Ref build(int first, int count, int depth) {
        Box box, cbox;
        for (int i = first; i < first + count; i++) { box.grow(tbox[order[i]]); cbox.grow(&cent[3 * order[i]]); }
        out->max_depth = std::max(out->max_depth, depth);
        if (count <= kLeafMax) return Ref{first, count, box, depth};

        // binned SAH over the three axes
        int bestAxis = -1, bestSplit = -1; float bestCost = FLT_MAX;
        for (int a = 0; a < 3; a++) {
            float ext = cbox.hi[a] - cbox.lo[a];
            if (!(ext > 0.0f)) continue;
            Box bb[kBins]; int bc[kBins] = {0};
            float scale = kBins / ext;
            for (int i = first; i < first + count; i++) {
                int b = std::min(kBins - 1, std::max(0, int((cent[3 * order[i] + a] - cbox.lo[a]) * scale)));
                bb[b].grow(tbox[order[i]]); bc[b]++;
            }
            float rightArea[kBins]; int rightCnt[kBins];
            Box acc; int cnt = 0;
            for (int b = kBins - 1; b > 0; b--) { acc.grow(bb[b]); cnt += bc[b]; rightArea[b] = acc.area(); rightCnt[b] = cnt; }
            acc = Box(); cnt = 0;
            for (int b = 0; b < kBins - 1; b++) {
                acc.grow(bb[b]); cnt += bc[b];
                if (cnt == 0 || rightCnt[b + 1] == 0) continue;
                float cost = acc.area() * cnt + rightArea[b + 1] * rightCnt[b + 1];
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = b; }
            }
        }
        int mid;
        if (bestAxis < 0) {                         // all centroids coincide: split in the middle
            mid = first + count / 2;
        } else {
            float ext = cbox.hi[bestAxis] - cbox.lo[bestAxis];
            float scale = kBins / ext; float lo = cbox.lo[bestAxis]; int a = bestAxis, sp = bestSplit;
            auto it = std::stable_partition(order.begin() + first, order.begin() + first + count, [&](int t) {
                int b = std::min(kBins - 1, std::max(0, int((cent[3 * t + a] - lo) * scale)));
                return b <= sp;
            });
            mid = int(it - order.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        int id = int(out->nodes.size());
        out->nodes.emplace_back();
        Ref l = build(first, mid - first, depth + 1);
        Ref r = build(mid, first + count - mid, depth + 1);
        BvhNode& n = out->nodes[id];
        for (int a = 0; a < 3; a++) {
            n.lo0[a] = l.box.lo[a] - pad; n.hi0[a] = l.box.hi[a] + pad;
            n.lo1[a] = r.box.lo[a] - pad; n.hi1[a] = r.box.hi[a] + pad;
        }
        n.child0 = l.child; n.count0 = l.count; n.child1 = r.child; n.count1 = r.count;
        return Ref{id, 0, box, depth};
    }
};

}  // namespace

Bvh build_bvh(const float* verts, int ntri) {
    Bvh bvh;
    Builder b; b.verts = verts; b.out = &bvh;
    b.tbox.resize(ntri); b.cent.resize(3 * size_t(ntri)); b.order.resize(ntri);
    float ext = 0.0f;
    for (int t = 0; t < ntri; t++) {
        for (int k = 0; k < 3; k++) {
            b.tbox[t].grow(verts + 9 * t + 3 * k);
            for (int a = 0; a < 3; a++) ext = std::max(ext, std::fabs(verts[9 * t + 3 * k + a]));
        }
        for (int a = 0; a < 3; a++) b.cent[3 * t + a] = 0.5f * (b.tbox[t].lo[a] + b.tbox[t].hi[a]);
    }
    // Padding: keeps the slab test conservative w.r.t. the per-triangle test's rounding (see oracle/tracer.h).
    b.pad = 2e-5f * ext + 1e-30f;
    std::iota(b.order.begin(), b.order.end(), 0);

    // root is always an inner node so that traversal has one code path
    bvh.nodes.emplace_back();
    BvhNode root; std::memset(&root, 0, sizeof root);
    auto emptyBox = [](float* lo, float* hi) { for (int a = 0; a < 3; a++) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; } };
    emptyBox(root.lo0, root.hi0); emptyBox(root.lo1, root.hi1);
    root.child0 = root.child1 = 0; root.count0 = root.count1 = -1;
    if (ntri > 0) {
        if (ntri <= Builder::kLeafMax) {
            Builder::Ref l = b.build(0, ntri, 1);
            for (int a = 0; a < 3; a++) { root.lo0[a] = l.box.lo[a] - b.pad; root.hi0[a] = l.box.hi[a] + b.pad; }
            root.child0 = l.child; root.count0 = l.count;
            bvh.nodes[0] = root;
        } else {
            bvh.nodes.clear();
            b.build(0, ntri, 1);        // emits the root as nodes[0]
        }
    } else {
        bvh.nodes[0] = root;
    }

    bvh.tris.resize(ntri);
    for (int i = 0; i < ntri; i++) {
        int t = b.order[i];
        const float* v = verts + 9 * t;
        TriGeom& g = bvh.tris[i];
        float e1[3], e2[3];
        for (int a = 0; a < 3; a++) { g.v0[a] = v[a]; e1[a] = v[3 + a] - v[a]; e2[a] = v[6 + a] - v[a]; }
        g.e1x = e1[0]; g.e1y = e1[1]; g.e1z = e1[2];
        g.e2x = e2[0]; g.e2y = e2[1]; g.e2z = e2[2];
        g.tri = uint32_t(t); g.pad0 = g.pad1 = 0;
    }
    return bvh;
}

}  // namespace romis
