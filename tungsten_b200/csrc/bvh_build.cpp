// Binned-SAH builder (host, multi-threaded) + collapse to the 4-ary device layout.  See bvh_build.h for what it replaces.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>

namespace tgb {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::infinity(); hi[a] = -lo[a]; } }
    void grow(const Box &b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    void grow(const float *p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (!(dx >= 0.0f)) return 0.0f;
        return dx*dy + dy*dz + dz*dx;
    }
};

struct Tmp { Box box; int32_t left, right; uint32_t first, count; };

constexpr int kBins = 16;
constexpr float kTraversalCost = 1.0f;

struct Builder {
    const Box *tb; const float *cent; uint32_t *idx; std::vector<Tmp> nodes; std::atomic<uint32_t> n_nodes{0};
    std::atomic<int> free_threads{0};
    uint32_t kMaxLeaf = 4; float kIntersectCost = 1.0f;      // leaf size cap (<= 4: two bits in the leaf link) and SAH cost of one primitive test

    uint32_t alloc() { return n_nodes.fetch_add(1); }

    void build(uint32_t node, uint32_t first, uint32_t count, int depth) {
        Tmp &nd = nodes[node];
        Box box, cb; box.reset(); cb.reset();
        for (uint32_t i = 0; i < count; ++i) { uint32_t t = idx[first + i]; box.grow(tb[t]); cb.grow(cent + 3*t); }
        nd.box = box; nd.first = first; nd.count = count; nd.left = nd.right = -1;
        if (count <= 1) return;

        int best_axis = -1, best_split = -1; float best_cost = std::numeric_limits<float>::infinity();
        for (int a = 0; a < 3; ++a) {
            float ext = cb.hi[a] - cb.lo[a];
            if (!(ext > 0.0f)) continue;
            Box bb[kBins]; uint32_t bc[kBins];
            for (int b = 0; b < kBins; ++b) { bb[b].reset(); bc[b] = 0; }
            float scale = kBins/ext;
            for (uint32_t i = 0; i < count; ++i) {
                uint32_t t = idx[first + i];
                int b = std::min(kBins - 1, std::max(0, int((cent[3*t + a] - cb.lo[a])*scale)));
                bb[b].grow(tb[t]); bc[b]++;
            }
            float right_area[kBins]; Box acc; acc.reset();
            for (int b = kBins - 1; b > 0; --b) { acc.grow(bb[b]); right_area[b] = acc.half_area(); }
            acc.reset(); uint32_t nl = 0;
            for (int b = 0; b < kBins - 1; ++b) {
                acc.grow(bb[b]); nl += bc[b];
                if (nl == 0 || nl == count) continue;
                float cost = acc.half_area()*nl + right_area[b + 1]*(count - nl);
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
            }
        }
        float leaf_cost = kIntersectCost*count;
        float split_cost = best_axis >= 0 ? kTraversalCost + kIntersectCost*best_cost/std::max(box.half_area(), 1e-30f)
                                          : std::numeric_limits<float>::infinity();
        if (count <= kMaxLeaf && !(split_cost < leaf_cost)) return;

        uint32_t mid;
        if (best_axis >= 0) {
            float ext = cb.hi[best_axis] - cb.lo[best_axis], scale = kBins/ext, lo = cb.lo[best_axis];
            const float *c = cent; int ax = best_axis, sp = best_split;
            uint32_t *m = std::partition(idx + first, idx + first + count, [=](uint32_t t) {
                int b = std::min(kBins - 1, std::max(0, int((c[3*t + ax] - lo)*scale)));
                return b <= sp;
            });
            mid = uint32_t(m - (idx + first));
        } else {
            mid = count/2;   // all centroids coincide: split by index
        }
        if (mid == 0 || mid == count) mid = count/2;
        uint32_t l = alloc(), r = alloc();
        nodes[node].left = int32_t(l); nodes[node].right = int32_t(r);
        bool spawned = false; std::thread th;
        if (count > 32768) {
            int avail = free_threads.load();
            while (avail > 0 && !free_threads.compare_exchange_weak(avail, avail - 1)) {}
            if (avail > 0) { spawned = true; th = std::thread([=] { build(l, first, mid, depth + 1); free_threads.fetch_add(1); }); }
        }
        if (!spawned) build(l, first, mid, depth + 1);
        build(r, first + mid, count - mid, depth + 1);
        if (spawned) th.join();
    }
};

inline float pad_down(float v) { return std::nextafter(std::nextafter(v, -std::numeric_limits<float>::infinity()), -std::numeric_limits<float>::infinity()); }
inline float pad_up(float v) { return std::nextafter(std::nextafter(v, std::numeric_limits<float>::infinity()), std::numeric_limits<float>::infinity()); }

}  // namespace

static void build_from_boxes(std::vector<Box> &tb, std::vector<float> &cent, uint32_t n, Bvh4 &out, int threads, uint32_t max_leaf = 4, float isect_cost = 1.0f);

void build_bvh4(const BuildTri *tris, uint32_t n, Bvh4 &out, int threads, float abs_pad, uint32_t max_leaf, float isect_cost) {
    out.nodes.clear(); out.order.clear(); out.max_depth = 0; out.sah_cost = 0.0;
    for (int a = 0; a < 3; ++a) { out.lo[a] = std::numeric_limits<float>::infinity(); out.hi[a] = -out.lo[a]; }
    if (n == 0) return;
    std::vector<Box> tb(n); std::vector<float> cent(3*size_t(n));
    for (uint32_t i = 0; i < n; ++i) {
        Box b; b.reset(); b.grow(tris[i].v0); b.grow(tris[i].v1); b.grow(tris[i].v2);
        for (int a = 0; a < 3; ++a) { cent[3*size_t(i) + a] = 0.5f*b.lo[a] + 0.5f*b.hi[a]; b.lo[a] = pad_down(b.lo[a] - abs_pad); b.hi[a] = pad_up(b.hi[a] + abs_pad); }
        tb[i] = b;
    }
    build_from_boxes(tb, cent, n, out, threads, std::min<uint32_t>(std::max<uint32_t>(max_leaf, 1), 4), isect_cost);
}

void build_bvh4_boxes(const BuildBox *boxes, uint32_t n, Bvh4 &out, int threads, float abs_pad, uint32_t max_leaf, float isect_cost) {
    out.nodes.clear(); out.order.clear(); out.max_depth = 0; out.sah_cost = 0.0;
    for (int a = 0; a < 3; ++a) { out.lo[a] = std::numeric_limits<float>::infinity(); out.hi[a] = -out.lo[a]; }
    if (n == 0) return;
    std::vector<Box> tb(n); std::vector<float> cent(3*size_t(n));
    for (uint32_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            cent[3*size_t(i) + a] = boxes[i].centroid[a];
            tb[i].lo[a] = pad_down(boxes[i].lo[a] - abs_pad); tb[i].hi[a] = pad_up(boxes[i].hi[a] + abs_pad);
        }
    build_from_boxes(tb, cent, n, out, threads, std::min<uint32_t>(std::max<uint32_t>(max_leaf, 1), 4), isect_cost);
}

static void build_from_boxes(std::vector<Box> &tb, std::vector<float> &cent, uint32_t n, Bvh4 &out, int threads, uint32_t max_leaf, float isect_cost) {
    out.order.resize(n);
    for (uint32_t i = 0; i < n; ++i) out.order[i] = i;
    Builder bl; bl.tb = tb.data(); bl.cent = cent.data(); bl.idx = out.order.data(); bl.kMaxLeaf = max_leaf; bl.kIntersectCost = isect_cost;
    const float kIntersectCost = isect_cost;
    bl.nodes.resize(2*size_t(n) + 1);
    if (threads <= 0) threads = int(std::thread::hardware_concurrency());
    bl.free_threads = std::max(0, threads - 1);
    uint32_t root = bl.alloc();
    bl.build(root, 0, n, 0);
    for (int a = 0; a < 3; ++a) { out.lo[a] = bl.nodes[root].box.lo[a]; out.hi[a] = bl.nodes[root].box.hi[a]; }

    auto leaf_link = [](const Tmp &t) { return ~int32_t((t.first << 3) | (t.count - 1)); };
    const Tmp &rt = bl.nodes[root];
    double root_area = std::max(double(rt.box.half_area()), 1e-30);
    auto set_child = [](Node4 &nd, int k, const Box &b, int32_t link) {
        nd.f[k] = b.lo[0]; nd.f[4 + k] = b.hi[0]; nd.f[8 + k] = b.lo[1]; nd.f[12 + k] = b.hi[1]; nd.f[16 + k] = b.lo[2]; nd.f[20 + k] = b.hi[2];
        nd.link[k] = link;
    };
    auto empty_node = []() {
        Node4 nd; std::memset(&nd, 0, sizeof(nd));
        for (int k = 0; k < 4; ++k) nd.link[k] = kEmptyLink;
        return nd;
    };
    if (rt.left < 0) {                       // the whole scene fits one leaf
        Node4 nd = empty_node();
        set_child(nd, 0, rt.box, leaf_link(rt));
        out.nodes.push_back(nd);
        out.max_depth = 1;
        return;
    }
    struct Item { uint32_t tmp; int32_t out_index; uint32_t depth; };
    std::vector<Item> stack; stack.push_back({root, 0, 1});
    out.nodes.resize(1);
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        const Tmp &t = bl.nodes[it.tmp];
        // collapse: start with the two children, split the largest inner child until four
        uint32_t ch[4]; int nc = 0;
        ch[nc++] = uint32_t(t.left); ch[nc++] = uint32_t(t.right);
        while (nc < 4) {
            int best = -1; float best_area = -1.0f;
            for (int k = 0; k < nc; ++k) {
                const Tmp &c = bl.nodes[ch[k]];
                if (c.left >= 0 && c.box.half_area() > best_area) { best_area = c.box.half_area(); best = k; }
            }
            if (best < 0) break;
            const Tmp &c = bl.nodes[ch[best]];
            ch[best] = uint32_t(c.left); ch[nc++] = uint32_t(c.right);
        }
        Node4 nd = empty_node();
        out.sah_cost += kTraversalCost*t.box.half_area()/root_area;
        out.max_depth = std::max(out.max_depth, it.depth + 1);
        for (int k = 0; k < nc; ++k) {
            const Tmp &c = bl.nodes[ch[k]];
            if (c.left < 0) {
                set_child(nd, k, c.box, leaf_link(c));
                out.sah_cost += kIntersectCost*c.count*c.box.half_area()/root_area;
            } else {
                int32_t idx = int32_t(out.nodes.size());
                set_child(nd, k, c.box, idx);
                out.nodes.emplace_back();
                stack.push_back({ch[k], idx, it.depth + 1});
            }
        }
        out.nodes[size_t(it.out_index)] = nd;
    }
}

// ---- quantised layout -------------------------------------------------------------------------------------------
namespace {
inline float node_half_area(const Node4 &nd, int k) {
    float dx = nd.f[4 + k] - nd.f[k], dy = nd.f[12 + k] - nd.f[8 + k], dz = nd.f[20 + k] - nd.f[16 + k];
    return dx*dy + dy*dz + dz*dx;
}
}  // namespace

void quantize_bvh4(const Bvh4 &in, uint32_t max_treelet, QBvh4 &out, int32_t empty_link) {
    out.nodes.clear(); out.old_index.clear(); out.treelet_image.clear(); out.n_treelet = 0;
    const size_t n = in.nodes.size();
    if (n == 0) return;
    // ---- order: treelet = repeatedly take the open inner node with the largest surface area (the nodes a random ray is
    // most likely to visit), parents before children; the rest depth-first under each treelet frontier node.
    std::vector<int32_t> order; order.reserve(n);
    std::vector<uint8_t> taken(n, 0);
    {
        struct Open { float area; int32_t node; };
        auto cmp = [](const Open &a, const Open &b) { return a.area < b.area || (a.area == b.area && a.node > b.node); };
        std::vector<Open> heap; heap.push_back({std::numeric_limits<float>::infinity(), 0});
        while (!heap.empty() && order.size() < size_t(max_treelet)) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            Open o = heap.back(); heap.pop_back();
            order.push_back(o.node); taken[size_t(o.node)] = 1;
            const Node4 &nd = in.nodes[size_t(o.node)];
            for (int k = 0; k < 4; ++k) if (nd.link[k] >= 0) { heap.push_back({node_half_area(nd, k), nd.link[k]}); std::push_heap(heap.begin(), heap.end(), cmp); }
        }
        out.n_treelet = uint32_t(order.size());
        // remaining nodes: depth-first from every frontier node, in treelet order (children near their parents)
        std::vector<int32_t> stack;
        auto dfs = [&](int32_t root) {
            stack.clear(); stack.push_back(root);
            while (!stack.empty()) {
                int32_t x = stack.back(); stack.pop_back();
                if (taken[size_t(x)]) continue;
                taken[size_t(x)] = 1; order.push_back(x);
                const Node4 &nd = in.nodes[size_t(x)];
                for (int k = 3; k >= 0; --k) if (nd.link[k] >= 0) stack.push_back(nd.link[k]);
            }
        };
        if (out.n_treelet == 0) dfs(0);
        for (uint32_t i = 0; i < out.n_treelet; ++i) {
            const Node4 &nd = in.nodes[size_t(order[i])];
            for (int k = 0; k < 4; ++k) if (nd.link[k] >= 0 && !taken[size_t(nd.link[k])]) dfs(nd.link[k]);
        }
    }
    std::vector<int32_t> new_index(n, -1);
    for (size_t i = 0; i < order.size(); ++i) new_index[size_t(order[i])] = int32_t(i);
    out.old_index = order;
    out.nodes.resize(order.size());
    for (size_t i = 0; i < order.size(); ++i) {
        const Node4 &nd = in.nodes[size_t(order[i])];
        QNode4 q; std::memset(&q, 0, sizeof(q));
        float org[3], S[3]; uint32_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a) {
            double mn = std::numeric_limits<double>::infinity(), mx = -mn;
            for (int k = 0; k < 4; ++k) if (nd.link[k] != kEmptyLink) { mn = std::min(mn, double(nd.f[8*a + k])); mx = std::max(mx, double(nd.f[8*a + 4 + k])); }
            if (!(mx >= mn)) { mn = 0.0; mx = 0.0; }
            // grid: plane(q) = org + q*2^e.  org sits at least 1/128 step below the lowest child plane so that q = 0 keeps the
            // same 1/256-step guard as every other value; e = the smallest exponent for which the highest plane still maps to <= 255.
            const double ext = mx - mn;
            int e = ext > 0.0 ? std::max(-120, int(std::floor(std::log2(ext/255.0))) - 1) : -120;
            for (;; ++e) {
                const double st = std::ldexp(1.0, e);
                float og = float(mn - st/64.0);
                while (double(og) > mn - st/128.0) og = std::nextafter(og, -std::numeric_limits<float>::infinity());
                if (std::ceil((mx - double(og))/st + 1.0/256.0) <= 255.0) { org[a] = og; break; }
            }
            const double step = std::ldexp(1.0, e);
            S[a] = float(std::ldexp(1.0, e + 15));
            for (int k = 0; k < 4; ++k) {
                uint32_t ql = 255, qh = 0;                      // empty child: inverted box
                if (nd.link[k] != kEmptyLink) {
                    double l = (double(nd.f[8*a + k]) - double(org[a]))/step - 1.0/256.0;
                    double h = (double(nd.f[8*a + 4 + k]) - double(org[a]))/step + 1.0/256.0;
                    ql = uint32_t(std::min(255.0, std::max(0.0, std::floor(l))));
                    qh = uint32_t(std::min(255.0, std::max(0.0, std::ceil(h))));
                }
                lo[a] |= ql << (8*k); hi[a] |= qh << (8*k);
            }
        }
        q.ox = org[0]; q.oy = org[1]; q.oz = org[2]; q.sx = S[0]; q.sy = S[1]; q.sz = S[2];
        q.lox = lo[0]; q.hix = hi[0]; q.loy = lo[1]; q.hiy = hi[1]; q.loz = lo[2]; q.hiz = hi[2];
        for (int k = 0; k < 4; ++k) q.link[k] = nd.link[k] >= 0 ? new_index[size_t(nd.link[k])] : (nd.link[k] == kEmptyLink ? empty_link : nd.link[k]);
        out.nodes[i] = q;
    }
    out.treelet_image.resize(out.n_treelet);
    for (uint32_t i = 0; i < out.n_treelet; ++i) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(&out.nodes[i]);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&out.treelet_image[i]);
        for (uint32_t c = 0; c < 4; ++c) std::memcpy(dst + 4*(c ^ ((i >> 1) & 3u)), src + 4*c, 16);
    }
}

void qnode_slab_host(const QNode4 &nd, const float o[3], const float inv_d[3], float tnear, float tfar, float t_out[4]) {
    // mirrors Traversal::visit_loaded in tgb_wavefront.cuh operation for operation (fmaf = one rounding, as FFMA)
    const float org[3] = {nd.ox, nd.oy, nd.oz}, S[3] = {nd.sx, nd.sy, nd.sz};
    const uint32_t lo[3] = {nd.lox, nd.loy, nd.loz}, hi[3] = {nd.hix, nd.hiy, nd.hiz};
    float a[3], b[3]; uint32_t nearw[3], farw[3];
    for (int ax = 0; ax < 3; ++ax) {
        const float ood = o[ax]*inv_d[ax];
        a[ax] = S[ax]*inv_d[ax];
        b[ax] = std::fmaf(org[ax], inv_d[ax], -ood) - a[ax];
        const bool neg = inv_d[ax] < 0.0f;
        nearw[ax] = neg ? hi[ax] : lo[ax]; farw[ax] = neg ? lo[ax] : hi[ax];
    }
    auto F = [](uint32_t w, int k) { uint32_t bits = 0x3F800000u | (((w >> (8*k)) & 0xFFu) << 8); float f; std::memcpy(&f, &bits, 4); return f; };
    for (int k = 0; k < 4; ++k) {
        float tn = tnear, tf = tfar;
        for (int ax = 0; ax < 3; ++ax) {
            tn = std::fmax(tn, std::fmaf(F(nearw[ax], k), a[ax], b[ax]));
            tf = std::fmin(tf, std::fmaf(F(farw[ax], k), a[ax], b[ax]));
        }
        t_out[k] = (tn <= tf && nd.link[k] != kEmptyLink) ? tn : std::numeric_limits<float>::infinity();
    }
}

}  // namespace tgb
