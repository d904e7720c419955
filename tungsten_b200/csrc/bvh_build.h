// Host-side BVH construction for the sm_100a traversal kernels.
//
// Replaces (reference): Embree's BVH4.Triangle4 object-split SAH builder used by TriangleMesh
// (src/thirdparty/embree/kernels/bvh/bvh_builder_sah.cpp:759, selected in kernels/common/scene.cpp:92-131)
// and the top-level user-geometry BVH of TraceableScene (src/core/renderer/TraceableScene.hpp:112-134).
// Not a port: one flattened world-space binned-SAH BVH2 over every mesh triangle of the scene, then laid
// out as 64-byte "two children per node" records so that one 64 B load feeds both child slab tests.
#pragma once
#include <cstdint>
#include <vector>

namespace tgb {

struct BuildTri { float v0[3], v1[3], v2[3]; };

// 64 B node: both children's boxes + links.
//   f[0..3]  = c0.lo.x c0.hi.x c0.lo.y c0.hi.y
//   f[4..7]  = c1.lo.x c1.hi.x c1.lo.y c1.hi.y
//   f[8..11] = c0.lo.z c0.hi.z c1.lo.z c1.hi.z
//   link[0], link[1]: >= 0 inner node index; < 0 leaf: ~link = (first << 3) | (count-1)
struct alignas(64) Node2 {
    float f[12];
    int32_t link[2];
    int32_t pad[2];
};

struct Bvh2 {
    std::vector<Node2> nodes;      // nodes[0] is the root pair; empty if there are no triangles
    std::vector<uint32_t> order;   // leaf order -> input triangle index
    int32_t root_link = 0;         // link of the root itself (leaf if the whole scene fits one leaf)
    float lo[3], hi[3];
    uint32_t max_depth = 0;
    double sah_cost = 0.0;
};

// threads <= 0: hardware concurrency.
// abs_pad: every triangle box is grown by this absolute amount (covers the slab test's rounding).
void build_bvh2(const BuildTri *tris, uint32_t n, Bvh2 &out, int threads = 0, float abs_pad = 0.0f);

}  // namespace tgb
