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

// ---- compressed device layout ("QNode4", 64 bytes = 4 x 16-byte loads per visit instead of 7) -------------------
// The four child boxes are stored as 8-bit offsets from the node's own origin on a power-of-two grid:
//     plane = origin + q * 2^e   (per axis),   q_lo = floor, q_hi = ceil  ->  the stored box CONTAINS the float box,
// so a traversal over QNode4 visits a superset of what the float tree visits and finds the same closest hit.
//   chunk 0: origin.x, origin.y, origin.z, Sx         S = 32768 * 2^e (the kernel builds 1 + q*2^-15 with one PRMT and
//   chunk 1: Sy, Sz, lo.x[4], hi.x[4]                   evaluates t = fma(1 + q*2^-15, S/d, (origin - o)/d - S/d))
//   chunk 2: lo.y[4], hi.y[4], lo.z[4], hi.z[4]       one byte per child, child k in bits 8k..8k+7
//   chunk 3: link[4]                                  as in Node4
// Rounding of the kernel's evaluation is covered by (a) the builder's abs_pad, exactly as for the float nodes, and
// (b) 1/256 of a grid step subtracted/added before floor/ceil (the S/d cancellation costs at most 1/512 step).
struct alignas(64) QNode4 {
    float ox, oy, oz, sx;
    float sy, sz; uint32_t lox, hix;
    uint32_t loy, hiy, loz, hiz;
    int32_t link[4];
};
static_assert(sizeof(QNode4) == 64, "QNode4 layout");

struct QBvh4 {
    std::vector<QNode4> nodes;        // renumbered: [0, n_treelet) = the top treelet (largest surface area first), rest depth-first
    std::vector<int32_t> old_index;   // new index -> index in the Node4 array it was made from
    uint32_t n_treelet = 0;
    std::vector<QNode4> treelet_image;   // the first n_treelet nodes with their 16-byte chunks XOR-swizzled for shared memory:
                                         // chunk c of node i sits at chunk c ^ ((i >> 1) & 3)  (conflict-free-ish LDS.128)
};
// Quantise `in` (float 4-ary BVH) into `out`; at most max_treelet nodes go to the treelet (0 = none).
// Empty child slots get the inverted box (lo 255, hi 0), which no ray can enter, and `empty_link` instead of kEmptyLink: the
// kernels do not test for empty slots, so the link must be harmless to follow (the caller passes a one-primitive leaf that
// holds an all-zero record, which every primitive test rejects).
void quantize_bvh4(const Bvh4 &in, uint32_t max_treelet, QBvh4 &out, int32_t empty_link = kEmptyLink);
// Host evaluation of one node visit with the kernels' arithmetic (tests + tgb200_qbvh_selftest): entry distance of the
// four children for the ray (o, 1/d), INFINITY where the slab test fails.
void qnode_slab_host(const QNode4 &nd, const float o[3], const float inv_d[3], float tnear, float tfar, float t_out[4]);

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
