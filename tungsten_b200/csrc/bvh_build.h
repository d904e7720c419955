// Host-side BVH construction for the sm_100a traversal kernels.
//
// Replaces (reference): Embree's BVH4.Triangle4 object-split SAH builder used by TriangleMesh
// (src/thirdparty/embree/kernels/bvh/bvh_builder_sah.cpp:759, selected in kernels/common/scene.cpp:92-131)
// and the top-level user-geometry BVH of TraceableScene (src/core/renderer/TraceableScene.hpp:112-134).
// Not a port: one flattened world-space binned-SAH binary tree over every mesh triangle of the scene, collapsed
// into 4-ary 128-byte nodes for the GPU traversal kernels.
#pragma once
#include <cstdint>
#include <vector>

namespace tgb {

struct BuildTri { float v0[3], v1[3], v2[3]; };

// 128-byte 4-ary node (the binned-SAH binary tree is collapsed: the child with the largest surface area is replaced by
// its own two children until there are four): per axis the four children's lo / hi planes as one float4 each, then links.
//   f[0..3] = lo.x of children 0..3, f[4..7] = hi.x, f[8..11] = lo.y, f[12..15] = hi.y, f[16..19] = lo.z, f[20..23] = hi.z
//   link[k] >= 0 inner node index; link[k] < 0 leaf: ~link = (first << 3) | (count-1); link[k] == kEmptyLink: no child
// Half as many dependent node fetches per ray as the binary layout (the traversal kernels are latency bound,
// profiles/r01_b_k_trace.md), leaves hold <= 4 triangles.
constexpr int32_t kEmptyLink = int32_t(0x80000000u);
struct alignas(128) Node4 {
    float f[24];
    int32_t link[4];
    int32_t pad[4];
};

struct Bvh4 {
    std::vector<Node4> nodes;      // nodes[0] is the root; empty if there are no triangles
    std::vector<uint32_t> order;   // leaf order -> input triangle index
    float lo[3], hi[3];
    uint32_t max_depth = 0;        // depth of the 4-ary tree (stack need <= 3*depth)
    double sah_cost = 0.0;
};

// A primitive given by its bounds and the point the SAH binning sorts it by (curve segments).
struct BuildBox { float lo[3], hi[3], centroid[3]; };

// threads <= 0: hardware concurrency.
// abs_pad: every triangle box is grown by this absolute amount (covers the slab test's rounding).
void build_bvh4(const BuildTri *tris, uint32_t n, Bvh4 &out, int threads = 0, float abs_pad = 0.0f,
                uint32_t max_leaf = 4, float isect_cost = 1.0f);
// max_leaf (1..4) and isect_cost (SAH cost of one primitive test relative to one node visit) let expensive primitives get
// smaller leaves.
void build_bvh4_boxes(const BuildBox *boxes, uint32_t n, Bvh4 &out, int threads = 0, float abs_pad = 0.0f,
                      uint32_t max_leaf = 4, float isect_cost = 1.0f);

}  // namespace tgb
