// Wavefront kernels of the B200 path tracer (sm_100a).  One thread per path / per ray query.
//   k_raygen  : SobolPathSampler::startPath + ReconstructionFilter::sample + PinholeCamera::sampleDirection
//   k_trace   : TraceableScene::intersect  (closest hit; analytic prims + BVH2 over all mesh triangles)
//   k_shade   : makeLocalScatterEvent + handleSurface (NEE/MIS query generation, emission, BSDF sample, RR)
//   k_shadow  : attenuatedEmission/generalizedShadowRay for the NEE and MIS queries (same traversal + epilogue)
//   k_accum   : folds the bounce's direct-light estimate into the path, NaN guards, compacts survivors
//   k_resolve : OutputBuffer::addSample running mean, samples folded in sample-index order
// State lives in SoA arrays indexed by a fixed path slot; queues hold slot ids.
#pragma once
#include "tgb_device.cuh"

namespace tgb {

// ---- path state (SoA, one entry per slot) ----------------------------------------------------
struct PathState {
    float *ox, *oy, *oz, *dx, *dy, *dz, *tmin;          // current ray (tmax is always +inf for path rays)
    float *tx, *ty, *tz;                                // throughput
    float *ex, *ey, *ez;                                // accumulated emission (the sample's radiance)
    uint64_t *pcg;                                      // supplemental PCG state
    uint32_t *info;                                     // dimension[0:16) | bounce[16:24) | flags[24:32)
    float *ht, *hu, *hv; int *hid;                      // closest hit of the current ray
    float *px, *py, *pz;                                // shading point of this bounce (origin of NEE/MIS queries)
    // direct-light estimate of this bounce, folded in by k_accum
    float *lx, *ly, *lz, *bx, *by, *bz, *wl, *sx, *sy, *sz, *ux, *uy, *uz;   // L, B, light weight, surface emission term, throughput before
    // NEE query payload: direction, expected distance, f, pdfL, pdfB ; MIS payload: direction, weight, pdfB
    float *ndx, *ndy, *ndz, *ndist, *nfx, *nfy, *nfz, *npl, *npb;
    float *mdx, *mdy, *mdz, *mwx, *mwy, *mwz, *mpb;
    int *qlight;                                        // light primitive of this bounce's queries
};
constexpr int kPathFloatArrays = 7 + 3 + 3 + 3 + 3 + 13 + 9 + 7;   // float-sized arrays in PathState (excl. pcg/info/hid/qlight)

enum : uint32_t { F_WAS_SPECULAR = 1u << 24, F_ALIVE = 1u << 25, F_FINAL_CHECK = 1u << 26, F_HAS_NEE = 1u << 27,
                  F_HAS_SURF = 1u << 28 };
constexpr int HID_MISS = -1;

struct Counters { unsigned long long rays, hits, shadow_rays, shadow_hits; };   // path queries / NEE+MIS queries

// ---- closest-hit traversal -------------------------------------------------------------------
struct Hit { float t, u, v; int id; };

TGB_D float xor_sign(float a, uint32_t sgn) { return __uint_as_float(__float_as_uint(a) ^ sgn); }
// Embree's Vec3 dot: x*x' + (y*y' + z*z') (thirdparty/embree/common/math/vec3.h:182)
TGB_D float edot(V3 a, V3 b) { return a.x*b.x + (a.y*b.y + a.z*b.z); }

// Quad::intersect (primitives/Quad.cpp:71-99)
TGB_D void quad_intersect(const DPrim &q, int self, V3 o, V3 d, float tnear, Hit &h) {
    float nDotW = dot(d, q.normal);
    if (fabsf(nDotW) < 1e-6f) return;
    float t = dot(q.normal, q.base - o)/nDotW;
    if (t < tnear || t > h.t) return;
    V3 qq = o + d*t;
    V3 v = qq - q.base;
    float l0 = dot(v, q.edge0)*q.inv_uv_sq0;
    float l1 = dot(v, q.edge1)*q.inv_uv_sq1;
    if (l0 < 0.0f || l0 > 1.0f || l1 < 0.0f || l1 > 1.0f) return;
    h.t = t; h.u = l0; h.v = l1; h.id = -(self + 2);
}
// Cube::intersect (primitives/Cube.cpp:94-127); backside flag travels in h.u
TGB_D void cube_intersect(const DPrim &c, int self, V3 o, V3 d, float tnear, Hit &h) {
    V3 p = m3mul(c.inv_rot, o - c.pos);
    V3 dl = m3mul(c.inv_rot, d);
    V3 invD = v3(1.0f/dl.x, 1.0f/dl.y, 1.0f/dl.z);
    V3 relMin = -c.scale - p;
    V3 relMax = c.scale - p;
    float ttMin = tnear, ttMax = h.t;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float id = comp(invD, i), rmin = comp(relMin, i), rmax = comp(relMax, i);
        if (id >= 0.0f) { ttMin = maxf(ttMin, rmin*id); ttMax = minf(ttMax, rmax*id); }
        else            { ttMax = minf(ttMax, rmin*id); ttMin = maxf(ttMin, rmax*id); }
    }
    if (ttMin <= ttMax) {
        if (ttMin > tnear && ttMin < h.t) { h.t = ttMin; h.u = 0.0f; h.v = 0.0f; h.id = -(self + 2); }
        else if (ttMax > tnear && ttMax < h.t) { h.t = ttMax; h.u = 1.0f; h.v = 0.0f; h.id = -(self + 2); }
    }
}

constexpr int kStackSize = 64;

// Closest hit over the whole scene.  Triangle test = Embree's MoellerTrumboreIntersector1
// (thirdparty/embree/kernels/geometry/triangle_intersector_moeller.h:75-111) with IEEE division for t,u,v.
#ifndef TGB_TRAV
#define TGB_TRAV 0
#endif
#ifndef TGB_FMA_SLAB
#define TGB_FMA_SLAB 0
#endif
#if TGB_TRAV == 0
// Per-lane while-while traversal (independent thread scheduling interleaves the divergent lane groups).
// ANY = true: occlusion query for a light whose own hit was already resolved analytically -- returns at the
// first surface in [tnear, tfar] other than primitive `ignore` (generalizedShadowRay's blocker test).
template <bool ANY>
TGB_D Hit trace_scene(const DScene &sc, bool active, V3 o, V3 d, float tnear, float tfar, int ignore) {
    Hit h; h.t = tfar; h.u = 0.0f; h.v = 0.0f; h.id = HID_MISS;
    if (!active) return h;
    for (int i = 0; i < sc.n_analytic; ++i) {
        int pi = sc.analytic[i];
        if (ANY && pi == ignore) continue;
        const DPrim &p = sc.prims[pi];
        if (p.type == TGB_PRIM_QUAD) quad_intersect(p, pi, o, d, tnear, h);
        else cube_intersect(p, pi, o, d, tnear, h);
        if (ANY && h.id != HID_MISS) return h;
    }
    if (sc.n_nodes == 0) return h;

    const float ooeps = 1e-30f;
    float idx = 1.0f/(fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
    float idy = 1.0f/(fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
    float idz = 1.0f/(fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
    int stack[kStackSize]; int sp = 0;
    int cur = 0;
    const float4 *nodes = sc.nodes;
    while (true) {
        while (cur >= 0) {
            const float4 n0 = __ldg(nodes + 4*cur), n1 = __ldg(nodes + 4*cur + 1), n2 = __ldg(nodes + 4*cur + 2);
            const float4 lk = __ldg(nodes + 4*cur + 3);
            float c0lox = (n0.x - o.x)*idx, c0hix = (n0.y - o.x)*idx, c0loy = (n0.z - o.y)*idy, c0hiy = (n0.w - o.y)*idy;
            float c1lox = (n1.x - o.x)*idx, c1hix = (n1.y - o.x)*idx, c1loy = (n1.z - o.y)*idy, c1hiy = (n1.w - o.y)*idy;
            float c0loz = (n2.x - o.z)*idz, c0hiz = (n2.y - o.z)*idz, c1loz = (n2.z - o.z)*idz, c1hiz = (n2.w - o.z)*idz;
            float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tnear));
            float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), h.t));
            float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tnear));
            float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), h.t));
            // 2-ulp slack on the far side: slab rounding must never cull a triangle the exact test accepts
            bool t0 = c0min <= c0max*1.0000003f, t1 = c1min <= c1max*1.0000003f;
            int l0 = __float_as_int(lk.x), l1 = __float_as_int(lk.y);
            if (t0 && t1) {
                bool swp = c1min < c0min;
                int nearc = swp ? l1 : l0, farc = swp ? l0 : l1;
                cur = nearc;
                if (sp < kStackSize) stack[sp++] = farc;
            } else if (t0) cur = l0;
            else if (t1) cur = l1;
            else {
                if (sp == 0) return h;
                cur = stack[--sp];
            }
            if (cur < 0) break;
        }
        // leaf
        {
            int code = ~cur;
            int first = code >> 3, count = (code & 7) + 1;
            for (int i = 0; i < count; ++i) {
                const float4 *tr = sc.tri_isect + 3*size_t(first + i);
                const float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
                V3 v0 = v3(a.x, a.y, a.z), e1 = v3(a.w, b.x, b.y), e2 = v3(b.z, b.w, c.x), ng = v3(c.y, c.z, c.w);
                V3 C = v0 - o;
                V3 R = cross(d, C);
                float den = edot(ng, d);
                float absDen = fabsf(den);
                uint32_t sgn = __float_as_uint(den) & 0x80000000u;
                float U = xor_sign(edot(R, e2), sgn);
                float V = xor_sign(edot(R, e1), sgn);
                if (!(den != 0.0f && U >= 0.0f && V >= 0.0f && U + V <= absDen)) continue;
                float T = xor_sign(edot(ng, C), sgn);
                if (!(T > absDen*tnear && T < absDen*h.t)) continue;
                h.t = T/absDen; h.u = U/absDen; h.v = V/absDen; h.id = first + i;
                if (ANY) return h;
            }
            if (sp == 0) return h;
            cur = stack[--sp];
        }
    }
}

#elif TGB_TRAV == 1
// ANY = true: occlusion query for a light whose own hit was already resolved analytically -- stops at the
// first surface in [tnear, tfar] other than primitive `ignore` (generalizedShadowRay's blocker test).
//
// WARP-SYNCHRONOUS: must be called by all 32 lanes of a converged warp; lanes without a ray pass active=false.
// Traversal is the "speculative while-while" scheme: all lanes walk inner nodes together until every lane
// has found a leaf (one leaf may be postponed so a lane can keep descending), then all lanes intersect their
// leaves together.  Warp votes (__any_sync) keep the two phases converged; without them leaf tests ran with
// ~1.6 of 32 lanes active (profiles/r01_k_trace_baseline.md).  The stack lives in local memory (L1-resident).
constexpr int TRAV_DONE = int(0x80000000u);

template <bool ANY>
TGB_D Hit trace_scene(const DScene &sc, bool active, V3 o, V3 d, float tnear, float tfar, int ignore) {
    const unsigned FULL = 0xffffffffu;
    Hit h; h.t = tfar; h.u = 0.0f; h.v = 0.0f; h.id = HID_MISS;
    if (active) {
        for (int i = 0; i < sc.n_analytic; ++i) {
            int pi = sc.analytic[i];
            if (ANY && pi == ignore) continue;
            const DPrim &p = sc.prims[pi];
            if (p.type == TGB_PRIM_QUAD) quad_intersect(p, pi, o, d, tnear, h);
            else cube_intersect(p, pi, o, d, tnear, h);
            if (ANY && h.id != HID_MISS) break;
        }
    }
    int cur = (active && sc.n_nodes != 0 && !(ANY && h.id != HID_MISS)) ? 0 : TRAV_DONE;
    if (!__any_sync(FULL, cur != TRAV_DONE)) return h;

    const float ooeps = 1e-30f;
    float idx = 1.0f/(fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
    float idy = 1.0f/(fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
    float idz = 1.0f/(fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
    int stack[kStackSize]; int sp = 0;
    int leaf = 0;                                   // postponed leaf (negative code) or 0
    const float4 *nodes = sc.nodes;
    while (true) {
        // ---- phase 1: inner nodes, until every lane holds a leaf or is done
        while (__any_sync(FULL, cur >= 0)) {
            if (cur >= 0) {
                const float4 n0 = __ldg(nodes + 4*cur), n1 = __ldg(nodes + 4*cur + 1), n2 = __ldg(nodes + 4*cur + 2);
                const float4 lk = __ldg(nodes + 4*cur + 3);
                float c0lox = (n0.x - o.x)*idx, c0hix = (n0.y - o.x)*idx, c0loy = (n0.z - o.y)*idy, c0hiy = (n0.w - o.y)*idy;
                float c1lox = (n1.x - o.x)*idx, c1hix = (n1.y - o.x)*idx, c1loy = (n1.z - o.y)*idy, c1hiy = (n1.w - o.y)*idy;
                float c0loz = (n2.x - o.z)*idz, c0hiz = (n2.y - o.z)*idz, c1loz = (n2.z - o.z)*idz, c1hiz = (n2.w - o.z)*idz;
                float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tnear));
                float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), h.t));
                float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tnear));
                float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), h.t));
                // 2-ulp slack on the far side: slab rounding must never cull a triangle the exact test accepts
                bool t0 = c0min <= c0max*1.0000003f, t1 = c1min <= c1max*1.0000003f;
                int l0 = __float_as_int(lk.x), l1 = __float_as_int(lk.y);
                if (t0 && t1) {
                    bool swp = c1min < c0min;
                    cur = swp ? l1 : l0;
                    if (sp < kStackSize) stack[sp++] = swp ? l0 : l1;
                } else if (t0) cur = l0;
                else if (t1) cur = l1;
                else cur = sp ? stack[--sp] : TRAV_DONE;
                if (cur < 0 && cur != TRAV_DONE && leaf == 0) {        // postpone the first leaf, keep descending
                    leaf = cur;
                    cur = sp ? stack[--sp] : TRAV_DONE;
                }
            }
        }
        // ---- phase 2: every lane intersects the leaves it holds (postponed first, then current)
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            int code;
            if (k == 0) { code = leaf; leaf = 0; }
            else { code = (cur != TRAV_DONE) ? cur : 0; if (code) cur = sp ? stack[--sp] : TRAV_DONE; }
            if (!__any_sync(FULL, code != 0)) continue;
            int first = (~code) >> 3, count = code ? ((~code) & 7) + 1 : 0;
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                if (!__any_sync(FULL, i < count)) break;
                if (i < count) {
                    const float4 *tr = sc.tri_isect + 3*size_t(first + i);
                    const float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
                    V3 v0 = v3(a.x, a.y, a.z), e1 = v3(a.w, b.x, b.y), e2 = v3(b.z, b.w, c.x), ng = v3(c.y, c.z, c.w);
                    V3 C = v0 - o;
                    V3 R = cross(d, C);
                    float den = edot(ng, d);
                    float absDen = fabsf(den);
                    uint32_t sgn = __float_as_uint(den) & 0x80000000u;
                    float U = xor_sign(edot(R, e2), sgn);
                    float V = xor_sign(edot(R, e1), sgn);
                    if (den != 0.0f && U >= 0.0f && V >= 0.0f && U + V <= absDen) {
                        float T = xor_sign(edot(ng, C), sgn);
                        if (T > absDen*tnear && T < absDen*h.t) {
                            h.t = T/absDen; h.u = U/absDen; h.v = V/absDen; h.id = first + i;
                        }
                    }
                }
            }
            if (ANY && h.id != HID_MISS) { cur = TRAV_DONE; leaf = 0; sp = 0; }
        }
        if (!__any_sync(FULL, cur != TRAV_DONE)) break;
    }
    return h;
}

#else
// "if-if" traversal: every loop iteration a lane performs ONE step -- either one inner-node visit (two child
// slab tests) or one triangle test of its current leaf -- and the lanes reconverge at the end of the
// iteration.  Compared with the per-lane while-while above this keeps the triangle tests from running with
// 1-2 active lanes (profiles/r01_a_k_trace_baseline.md).
template <bool ANY>
TGB_D Hit trace_scene(const DScene &sc, bool active, V3 o, V3 d, float tnear, float tfar, int ignore) {
    Hit h; h.t = tfar; h.u = 0.0f; h.v = 0.0f; h.id = HID_MISS;
    if (!active) return h;
    for (int i = 0; i < sc.n_analytic; ++i) {
        int pi = sc.analytic[i];
        if (ANY && pi == ignore) continue;
        const DPrim &p = sc.prims[pi];
        if (p.type == TGB_PRIM_QUAD) quad_intersect(p, pi, o, d, tnear, h);
        else cube_intersect(p, pi, o, d, tnear, h);
        if (ANY && h.id != HID_MISS) return h;
    }
    if (sc.n_nodes == 0) return h;
    const float ooeps = 1e-30f;
    float idx = 1.0f/(fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
    float idy = 1.0f/(fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
    float idz = 1.0f/(fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
#if TGB_FMA_SLAB
    float oodx = o.x*idx, oody = o.y*idy, oodz = o.z*idz;
#endif
    int stack[kStackSize]; int sp = 0;
    int cur = 0;
    const float4 *nodes = sc.nodes;
    const int DONE = int(0x80000000u);
    while (cur != DONE) {
        if (cur >= 0) {
            const float4 n0 = __ldg(nodes + 4*cur), n1 = __ldg(nodes + 4*cur + 1), n2 = __ldg(nodes + 4*cur + 2);
            const float4 lk = __ldg(nodes + 4*cur + 3);
#if TGB_FMA_SLAB
            float c0lox = __fmaf_rn(n0.x, idx, -oodx), c0hix = __fmaf_rn(n0.y, idx, -oodx), c0loy = __fmaf_rn(n0.z, idy, -oody), c0hiy = __fmaf_rn(n0.w, idy, -oody);
            float c1lox = __fmaf_rn(n1.x, idx, -oodx), c1hix = __fmaf_rn(n1.y, idx, -oodx), c1loy = __fmaf_rn(n1.z, idy, -oody), c1hiy = __fmaf_rn(n1.w, idy, -oody);
            float c0loz = __fmaf_rn(n2.x, idz, -oodz), c0hiz = __fmaf_rn(n2.y, idz, -oodz), c1loz = __fmaf_rn(n2.z, idz, -oodz), c1hiz = __fmaf_rn(n2.w, idz, -oodz);
#else
            float c0lox = (n0.x - o.x)*idx, c0hix = (n0.y - o.x)*idx, c0loy = (n0.z - o.y)*idy, c0hiy = (n0.w - o.y)*idy;
            float c1lox = (n1.x - o.x)*idx, c1hix = (n1.y - o.x)*idx, c1loy = (n1.z - o.y)*idy, c1hiy = (n1.w - o.y)*idy;
            float c0loz = (n2.x - o.z)*idz, c0hiz = (n2.y - o.z)*idz, c1loz = (n2.z - o.z)*idz, c1hiz = (n2.w - o.z)*idz;
#endif
            float c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tnear));
            float c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), h.t));
            float c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tnear));
            float c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), h.t));
            bool t0 = c0min <= c0max*1.0000003f, t1 = c1min <= c1max*1.0000003f;
            int l0 = __float_as_int(lk.x), l1 = __float_as_int(lk.y);
            if (t0 && t1) {
                bool swp = c1min < c0min;
                cur = swp ? l1 : l0;
                if (sp < kStackSize) stack[sp++] = swp ? l0 : l1;
            } else if (t0) cur = l0;
            else if (t1) cur = l1;
            else cur = sp ? stack[--sp] : DONE;
        } else {
            int code = ~cur;
            int first = code >> 3, rem = code & 7;            // rem = triangles left after this one
            const float4 *tr = sc.tri_isect + 3*size_t(first);
            const float4 a = __ldg(tr), b = __ldg(tr + 1), c = __ldg(tr + 2);
            V3 v0 = v3(a.x, a.y, a.z), e1 = v3(a.w, b.x, b.y), e2 = v3(b.z, b.w, c.x), ng = v3(c.y, c.z, c.w);
            V3 C = v0 - o;
            V3 R = cross(d, C);
            float den = edot(ng, d);
            float absDen = fabsf(den);
            uint32_t sgn = __float_as_uint(den) & 0x80000000u;
            float U = xor_sign(edot(R, e2), sgn);
            float V = xor_sign(edot(R, e1), sgn);
            bool hit = false;
            if (den != 0.0f && U >= 0.0f && V >= 0.0f && U + V <= absDen) {
                float T = xor_sign(edot(ng, C), sgn);
                if (T > absDen*tnear && T < absDen*h.t) {
                    h.t = T/absDen; h.u = U/absDen; h.v = V/absDen; h.id = first; hit = true;
                }
            }
            if (ANY && hit) return h;
            cur = rem ? ~(((first + 1) << 3) | (rem - 1)) : (sp ? stack[--sp] : DONE);
        }
    }
    return h;
}
#endif
TGB_D Hit trace_closest(const DScene &sc, bool active, V3 o, V3 d, float tnear, float tfar) { return trace_scene<false>(sc, active, o, d, tnear, tfar, -1); }

// Fill a Surface from a hit: Primitive::intersectionInfo for mesh/quad/cube
// (TriangleMesh.cpp:323-331,344-355; Quad.cpp:112-120; Cube.cpp:157-171) + TraceableScene::intersect (:183-188)
TGB_D void make_surface(const DScene &sc, const Hit &h, V3 o, V3 d, Surface &s) {
    s.p = o + d*h.t;
    s.w = d;
    if (h.id >= 0) {
        uint32_t g = __ldg(sc.tri_global + h.id);
        int pi = int(__ldg(sc.tri_prim + g));
        const DPrim &m = sc.prims[pi];
        s.prim = pi;
        const float4 c = __ldg(sc.tri_isect + 3*size_t(h.id) + 2);
        // (p1-p0)x(p2-p0) == -(e1 x e2) exactly (e1 = p0-p1, e2 = p2-p0; negation is exact in IEEE)
        V3 isectNg = v3(-c.y, -c.z, -c.w);
        s.backside = dot(isectNg, d) > 0.0f;
        s.Ng = normalize(isectNg);
        const float4 *sh = sc.tri_shade + 4*size_t(g);
        const float4 s0 = __ldg(sh), s1 = __ldg(sh + 1), s2 = __ldg(sh + 2), s3 = __ldg(sh + 3);
        float u = h.u, v = h.v;
        if (m.flags & PF_SMOOTH) {
            V3 n0 = v3(s0.x, s0.y, s0.z), n1 = v3(s0.w, s1.x, s1.y), n2 = v3(s1.z, s1.w, s2.x);
            s.Ns = normalize(n0*(1.0f - u - v) + n1*u + n2*v);
        } else s.Ns = s.Ng;
        float w0 = 1.0f - u - v;
        s.u = w0*s2.y + u*s2.w + v*s3.y;
        s.v = w0*s2.z + u*s3.x + v*s3.z;
        int mat = __float_as_int(s3.w);
        s.bsdf = int(__ldg(sc.slots + m.bsdf_first + mat));
    } else {
        int pi = -h.id - 2;
        const DPrim &p = sc.prims[pi];
        s.prim = pi;
        s.bsdf = int(__ldg(sc.slots + p.bsdf_first));
        if (p.type == TGB_PRIM_QUAD) {
            s.Ng = s.Ns = p.normal; s.u = h.u; s.v = h.v;
            s.backside = dot(d, p.normal) >= 0.0f;
        } else {
            V3 q = m3mul(p.inv_rot, s.p - p.pos);
            V3 dd = vabs(q) - p.scale;
            int dim = 0; float mx = dd.x; if (dd.y > mx) { mx = dd.y; dim = 1; } if (dd.z > mx) { mx = dd.z; dim = 2; }
            float sgn = comp(q, dim) < 0.0f ? -1.0f : 1.0f;
            V3 n = v3(dim == 0 ? sgn : 0.0f, dim == 1 ? sgn : 0.0f, dim == 2 ? sgn : 0.0f);
            V3 uvw = v3((q.x/p.scale.x)*0.5f + 0.5f, (q.y/p.scale.y)*0.5f + 0.5f, (q.z/p.scale.z)*0.5f + 0.5f);
            s.Ns = s.Ng = m3mul(p.rot, n);
            s.u = comp(uvw, (dim + 1) % 3); s.v = comp(uvw, (dim + 2) % 3);
            s.backside = h.u != 0.0f;
        }
    }
}

// Primitive::evalDirect (Quad.cpp:235-238, TriangleMesh.cpp:493-496, Cube evalDirect)
TGB_D V3 eval_direct(const DScene &sc, const Surface &s) {
    const DPrim &p = sc.prims[s.prim];
    if (p.emission_tex < 0 || s.backside) return v3s(0.0f);
    return tex_eval(sc.tex[p.emission_tex], s.u, s.v);
}

// light.intersect(ray) for a quad light, resolved in the shading kernel (Quad.cpp:71-99 on the light alone,
// as TraceBase::attenuatedEmission does at TraceBase.cpp:160) so that the traced query is a pure occlusion test.
TGB_D bool quad_light_hit(const DPrim &l, V3 p, V3 d, float tnear, float &t, float &l0, float &l1, bool &backside) {
    float nDotW = dot(d, l.normal);
    if (fabsf(nDotW) < 1e-6f) return false;
    t = dot(l.normal, l.base - p)/nDotW;
    if (t < tnear) return false;
    V3 v = (p + d*t) - l.base;
    l0 = dot(v, l.edge0)*l.inv_uv_sq0;
    l1 = dot(v, l.edge1)*l.inv_uv_sq1;
    if (l0 < 0.0f || l0 > 1.0f || l1 < 0.0f || l1 > 1.0f) return false;
    backside = nDotW >= 0.0f;
    return true;
}

// ---- kernels ---------------------------------------------------------------------------------
struct BatchInfo { const uint32_t *pix_id, *pix_seed; uint32_t n_pix, spp_begin, n_paths; };

__global__ void __launch_bounds__(256) k_raygen(DScene sc, PathState st, BatchInfo bi, uint32_t *queue, uint32_t *count) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i == 0) *count = bi.n_paths;
    if (i >= bi.n_paths) return;
    uint32_t pix = i % bi.n_pix, smp_i = bi.spp_begin + i/bi.n_pix;
    uint32_t pixel_id = __ldg(bi.pix_id + pix);
    Sampler smp; sampler_start(smp, sc.sobol, __ldg(bi.pix_seed + pix), pixel_id, smp_i);
    uint32_t px = pixel_id % sc.cam.res_x, py = pixel_id/sc.cam.res_x;
    // PinholeCamera::sampleDirection (cameras/PinholeCamera.cpp:70-86)
    float xa = sampler_next1d(smp), xb = sampler_next1d(smp);
    float fu, fv;
    if (sc.cam.filter == TGB_FILTER_DIRAC) { fu = 0.0f; fv = 0.0f; }
    else if (sc.cam.filter == TGB_FILTER_BOX) { fu = xa - 0.5f; fv = xb - 0.5f; }
    else { fu = filter_sample1(sc.cam, xa); fv = filter_sample1(sc.cam, xb); }
    V3 localD = normalize(v3(-1.0f + (float(px) + 0.5f + fu)*2.0f*sc.cam.pixel_size_x,
                             sc.cam.ratio - (float(py) + 0.5f + fv)*2.0f*sc.cam.pixel_size_x,
                             sc.cam.plane_dist));
    V3 d = m3mul(sc.cam.m, localD);
    st.ox[i] = sc.cam.pos.x; st.oy[i] = sc.cam.pos.y; st.oz[i] = sc.cam.pos.z;
    st.dx[i] = d.x; st.dy[i] = d.y; st.dz[i] = d.z; st.tmin[i] = 1e-4f;                   // math/Ray.hpp:24
    st.tx[i] = 1.0f; st.ty[i] = 1.0f; st.tz[i] = 1.0f;
    st.ex[i] = 0.0f; st.ey[i] = 0.0f; st.ez[i] = 0.0f;
    st.pcg[i] = smp.pcg;
    st.info[i] = smp.dimension | F_WAS_SPECULAR | F_ALIVE;
    queue[i] = i;
}

template <bool SHADOW>
TGB_D void count_rays(Counters *ctr, bool valid, bool hit) {
    unsigned mv = __ballot_sync(0xffffffffu, valid), mh = __ballot_sync(0xffffffffu, valid && hit);
    if ((threadIdx.x & 31) == 0 && mv) {
        atomicAdd(SHADOW ? &ctr->shadow_rays : &ctr->rays, (unsigned long long)__popc(mv));
        if (mh) atomicAdd(SHADOW ? &ctr->shadow_hits : &ctr->hits, (unsigned long long)__popc(mh));
    }
}

__global__ void __launch_bounds__(128) k_trace(DScene sc, PathState st, const uint32_t *queue, const uint32_t *count, Counters *ctr) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    bool valid = i < *count;
    uint32_t s = 0; V3 o = v3s(0.0f), d = v3(0.0f, 0.0f, 1.0f); float tmin = 0.0f;
    if (valid) {
        s = queue[i];
        o = v3(st.ox[s], st.oy[s], st.oz[s]); d = v3(st.dx[s], st.dy[s], st.dz[s]); tmin = st.tmin[s];
    }
    Hit h = trace_closest(sc, valid, o, d, tmin, INFINITY);
    if (valid) { st.ht[s] = h.t; st.hu[s] = h.u; st.hv[s] = h.v; st.hid[s] = h.id; }
    count_rays<false>(ctr, valid, h.id != HID_MISS);
}

// Parity hook: rays in AoS tgb_ray, hits out as tgb_hit (tgb200_trace_closest).
__global__ void __launch_bounds__(128) k_trace_rays(DScene sc, const tgb_ray *rays, tgb_hit *hits, uint32_t n) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    bool valid = i < n;
    V3 o = v3s(0.0f), d = v3(0.0f, 0.0f, 1.0f); float t0 = 0.0f, t1 = 0.0f;
    if (valid) { o = v3(rays[i].o[0], rays[i].o[1], rays[i].o[2]); d = v3(rays[i].d[0], rays[i].d[1], rays[i].d[2]); t0 = rays[i].tmin; t1 = rays[i].tmax; }
    Hit h = trace_closest(sc, valid, o, d, t0, t1);
    if (!valid) return;
    tgb_hit out; out.primitive = -1; out.prim_id = 0; out.t = h.t; out.u = 0.0f; out.v = 0.0f; out.backside = 0;
    if (h.id != HID_MISS) {
        Surface s; make_surface(sc, h, o, d, s);
        out.primitive = s.prim; out.backside = s.backside ? 1u : 0u;
        if (h.id >= 0) { uint32_t g = sc.tri_global[h.id]; out.prim_id = int(g - sc.prims[s.prim].tri_first); out.u = h.u; out.v = h.v; }
        else if (sc.prims[s.prim].type == TGB_PRIM_QUAD) { out.u = h.u; out.v = h.v; }
    }
    hits[i] = out;
}

// handleSurface (integrators/TraceBase.cpp:516-568) + the loop tail of traceSample (PathTracer.cpp:108-126)
__global__ void __launch_bounds__(128) k_shade(DScene sc, PathState st, BatchInfo bi, const uint32_t *queue, const uint32_t *count,
                                               uint32_t *squeue, uint32_t *scount) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    bool valid = i < *count;
    bool qn = false, qm = false, qn_any = false, qm_any = false;
    uint32_t s = 0;
    if (valid) {
        s = queue[i];
        uint32_t info = st.info[s];
        int bounce = int((info >> 16) & 0xFFu);
        bool wasSpecular = (info & F_WAS_SPECULAR) != 0;
        V3 o = v3(st.ox[s], st.oy[s], st.oz[s]), d = v3(st.dx[s], st.dy[s], st.dz[s]);
        V3 thr = v3(st.tx[s], st.ty[s], st.tz[s]);
        Hit h; h.t = st.ht[s]; h.u = st.hu[s]; h.v = st.hv[s]; h.id = st.hid[s];
        const tgb_settings &set = sc.set;
        uint32_t flags = 0;

        if (h.id == HID_MISS) {
            // loop exit with didHit == false: handleInfiniteLights (TraceBase.cpp:570-578, TraceableScene.hpp:194-209)
            if (bounce >= set.min_bounces && bounce < set.max_bounces && sc.n_inf_lights > 0) {
                int li = sc.inf_lights[sc.n_inf_lights - 1];
                const DPrim &l = sc.prims[li];
                if (!set.enable_light_sampling || wasSpecular || !(l.flags & PF_SAMPLABLE)) {
                    float u, v; direction_to_uv(l, d, u, v, nullptr);
                    V3 em = tex_eval(sc.tex[l.emission_tex], u, v);
                    st.ex[s] += thr.x*em.x; st.ey[s] += thr.y*em.y; st.ez[s] += thr.z*em.z;
                }
            }
            st.info[s] = (info & 0x00FFFFFFu) | F_FINAL_CHECK;
        } else {
            Sampler smp; smp.sobol = sc.sobol; smp.pcg = st.pcg[s]; smp.dimension = info & 0xFFFFu;
            {
                uint32_t pix = s % bi.n_pix;
                smp.index = bi.spp_begin + s/bi.n_pix;
                smp.scramble = __ldg(bi.pix_seed + pix) ^ hash32(__ldg(bi.pix_id + pix));
            }
            Surface sf; make_surface(sc, h, o, d, sf);
            const DBsdf &b = sc.bsdfs[sf.bsdf];
            const float epsilon = 5e-4f;                                                     // TraceableScene.hpp:39

            // makeLocalScatterEvent (TraceBase.cpp:24-51)
            Event e;
            {
                Frame frame = frame_from_normal(sf.Ns);
                bool hitBackside = dot(frame.n, d) > 0.0f;
                bool isTransmissive = (b.lobes & LOBE_TRANSMISSIVE) != 0;
                bool flipFrame = set.enable_two_sided_shading && hitBackside && !isTransmissive;
                if (flipFrame) { frame.n = -frame.n; frame.t = -frame.t; }
                e.frame = frame; e.wi = to_local(frame, -d); e.wo = v3s(0.0f); e.weight = v3s(1.0f); e.pdf = 1.0f;
                e.requested = LOBE_ALL; e.sampled = 0; e.flipped = flipFrame;
            }
            // forward-transparency coin flip: transparency == 0 for every lobe in scope, the draw is still made (:525-529)
            (void)sampler_boolean(smp, 0.0f);

            st.px[s] = sf.p.x; st.py[s] = sf.p.y; st.pz[s] = sf.p.z;
            if (set.enable_light_sampling && bounce < set.max_bounces - 1) {
                // estimateDirect -> chooseLight -> sampleDirect (TraceBase.cpp:483-494,416-459,383-400)
                float weight;
                int li = choose_light(sc, sf.p, smp, weight);
                bool pureSpecular = b.lobes != 0 && (b.lobes & ~uint32_t(LOBE_SPECULAR)) == 0;
                if (li >= 0 && !pureSpecular && b.lobes != LOBE_FORWARD) {
                    const DPrim &l = sc.prims[li];
                    // lightSample (TraceBase.cpp:246-285).  For quad / environment lights the light's own hit
                    // (attenuatedEmission, :160-165) and lightF (:279-284) are resolved here and the query that is
                    // traced is a pure blocker test; mesh lights keep the closest-hit query + epilogue.
                    LightSample ls;
                    if (light_sample_direct(sc, l, sf.p, smp, ls)) {
                        e.wo = to_local(e.frame, ls.d);
                        bool ok = true;
                        if (set.enable_consistency_checks)                                     // isConsistent (:53-60)
                            ok = (dot(ls.d, sf.Ng) < 0.0f) == ((e.wo.z < 0.0f) != e.flipped);
                        if (ok) {
                            e.requested = LOBE_ALL_BUT_SPECULAR;
                            V3 f = bsdf_eval(sc, b, sf, e);
                            if (!is_zero(f)) {
                                float pdfB = bsdf_pdf(sc, b, sf, e);
                                if (l.type == TGB_PRIM_MESH) {
                                    qn = true; qn_any = false;
                                    st.ndist[s] = ls.dist; st.nfx[s] = f.x; st.nfy[s] = f.y; st.nfz[s] = f.z;
                                    st.npl[s] = ls.pdf; st.npb[s] = pdfB;
                                } else {
                                    V3 em = v3s(0.0f); float tfar = INFINITY; bool hitL = true;
                                    if (l.type == TGB_PRIM_QUAD) {
                                        float t, l0, l1; bool back;
                                        hitL = quad_light_hit(l, sf.p, ls.d, epsilon, t, l0, l1, back) && !(t*(1.0f + 1e-3f) < ls.dist);
                                        if (hitL && !back) em = tex_eval(sc.tex[l.emission_tex], l0, l1);
                                        tfar = t;
                                    } else {
                                        float u, v; direction_to_uv(l, ls.d, u, v, nullptr);
                                        em = tex_eval(sc.tex[l.emission_tex], u, v);
                                    }
                                    if (hitL && !is_zero(em)) {
                                        V3 lightF = (f*em)/ls.pdf;
                                        lightF = lightF*power_heuristic(ls.pdf, pdfB);
                                        qn = true; qn_any = true;
                                        st.ndist[s] = tfar; st.nfx[s] = lightF.x; st.nfy[s] = lightF.y; st.nfz[s] = lightF.z;
                                    }
                                }
                                if (qn) { st.ndx[s] = ls.d.x; st.ndy[s] = ls.d.y; st.ndz[s] = ls.d.z; }
                            }
                        }
                    }
                    // bsdfSample (TraceBase.cpp:287-321)
                    e.requested = LOBE_ALL_BUT_SPECULAR;
                    if (bsdf_sample(sc, b, sf, smp, e) && !is_zero(e.weight)) {
                        V3 wo = to_global(e.frame, e.wo);
                        bool ok = true;
                        if (set.enable_consistency_checks)
                            ok = (dot(wo, sf.Ng) < 0.0f) == ((e.wo.z < 0.0f) != e.flipped);
                        if (ok) {
                            if (l.type == TGB_PRIM_MESH) {
                                qm = true; qm_any = false;
                                st.mwx[s] = e.weight.x; st.mwy[s] = e.weight.y; st.mwz[s] = e.weight.z; st.mpb[s] = e.pdf;
                            } else {
                                V3 em = v3s(0.0f); float tfar = INFINITY, directPdf; bool hitL = true;
                                if (l.type == TGB_PRIM_QUAD) {
                                    float t, l0, l1; bool back;
                                    hitL = quad_light_hit(l, sf.p, wo, epsilon, t, l0, l1, back);
                                    if (hitL && !back) em = tex_eval(sc.tex[l.emission_tex], l0, l1);
                                    tfar = t;
                                    directPdf = t*t/(fabsf(dot(l.normal, wo))*l.area);                 // Quad::directPdf (Quad.cpp:216-222)
                                } else {
                                    float u, v, sinTheta; direction_to_uv(l, wo, u, v, &sinTheta);
                                    const DTex &et = sc.tex[l.emission_tex];
                                    em = tex_eval(et, u, v);                                              // InfiniteSphere.cpp:218-229,241-244
                                    directPdf = et.type == TGB_TEX_CONSTANT ? INV_FOUR_PI_F : INV_PI_F*INV_TWO_PI_F*bitmap_pdf_uv(et, u, v)/sinTheta;
                                }
                                if (hitL && !is_zero(em)) {
                                    V3 bsdfF = em*e.weight;
                                    bsdfF = bsdfF*power_heuristic(e.pdf, directPdf);
                                    qm = true; qm_any = true;
                                    st.mwx[s] = bsdfF.x; st.mwy[s] = bsdfF.y; st.mwz[s] = bsdfF.z; st.mpb[s] = tfar;
                                }
                            }
                            if (qm) { st.mdx[s] = wo.x; st.mdy[s] = wo.y; st.mdz[s] = wo.z; }
                        }
                    }
                    // generalizedShadowRay returns 0 unless bounce+1 >= minBounces (TraceBase.cpp:117)
                    if (bounce + 1 < set.min_bounces) { qn = false; qm = false; }
                    if (qn || qm) {
                        flags |= F_HAS_NEE;
                        st.qlight[s] = li; st.wl[s] = weight;
                        st.ux[s] = thr.x; st.uy[s] = thr.y; st.uz[s] = thr.z;
                        st.lx[s] = 0.0f; st.ly[s] = 0.0f; st.lz[s] = 0.0f; st.bx[s] = 0.0f; st.by[s] = 0.0f; st.bz[s] = 0.0f;
                    }
                }
            }

            // emission of the surface itself (TraceBase.cpp:540-543)
            const DPrim &prim = sc.prims[sf.prim];
            if ((prim.flags & PF_EMISSIVE) && bounce >= set.min_bounces) {
                if (!set.enable_light_sampling || wasSpecular || !(prim.flags & PF_SAMPLABLE)) {
                    V3 em = eval_direct(sc, sf)*thr;
                    st.sx[s] = em.x; st.sy[s] = em.y; st.sz[s] = em.z;
                    flags |= F_HAS_SURF;
                }
            }

            // continuation sample (TraceBase.cpp:545-565)
            uint32_t status = 0;
            e.requested = LOBE_ALL;
            bool cont = bsdf_sample(sc, b, sf, smp, e);
            V3 wo = v3s(0.0f);
            if (cont) {
                wo = to_global(e.frame, e.wo);
                if (set.enable_consistency_checks && (dot(wo, sf.Ng) < 0.0f) != ((e.wo.z < 0.0f) != e.flipped)) cont = false;
            }
            if (cont) {
                thr = thr*e.weight;
                wasSpecular = (e.sampled & LOBE_SPECULAR) != 0;
                st.ox[s] = sf.p.x; st.oy[s] = sf.p.y; st.oz[s] = sf.p.z;      // ray.hitpoint() == info.p for every primitive in scope
                st.dx[s] = wo.x; st.dy[s] = wo.y; st.dz[s] = wo.z; st.tmin[s] = epsilon;
                // PathTracer.cpp:108-117
                if (max_comp(thr) == 0.0f) status = F_FINAL_CHECK;
                else {
                    float roulettePdf = max_comp(vabs(thr));
                    bool alive = true;
                    if (bounce > 2 && roulettePdf < 0.1f) {
                        if (sampler_boolean(smp, roulettePdf)) thr = thr/roulettePdf;
                        else alive = false;
                    }
                    if (alive) {
                        bounce++;
                        status = bounce < set.max_bounces ? F_ALIVE : (F_ALIVE | F_FINAL_CHECK);
                    }
                }
                st.tx[s] = thr.x; st.ty[s] = thr.y; st.tz[s] = thr.z;
            }
            st.pcg[s] = smp.pcg;
            st.info[s] = (smp.dimension & 0xFFFFu) | (uint32_t(bounce) << 16) | (wasSpecular ? F_WAS_SPECULAR : 0u) | status | flags;
        }
    }
    // enqueue this bounce's shadow queries: warp-vote compaction, one atomic per warp
    {
        unsigned mn = __ballot_sync(0xffffffffu, qn), mm = __ballot_sync(0xffffffffu, qm);
        unsigned total = __popc(mn) + __popc(mm);
        if (total) {
            unsigned lane = threadIdx.x & 31, base = 0;
            if (lane == 0) base = atomicAdd(scount, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            unsigned lt = (1u << lane) - 1u;
            if (qn) squeue[base + __popc(mn & lt)] = (s << 2) | (qn_any ? 2u : 0u);
            if (qm) squeue[base + __popc(mn) + __popc(mm & lt)] = (s << 2) | 1u | (qm_any ? 2u : 0u);
        }
    }
}

// attenuatedEmission + generalizedShadowRay (TraceBase.cpp:144-174,62-125) for one NEE or MIS query.
//  * "any" queries (quad / environment lights): lightF / bsdfF were finished by k_shade; this kernel only looks for
//    a blocker in [epsilon, t_light] other than the light and stores the value if there is none;
//  * mesh-light queries: one closest-hit query decides visibility (the light is part of the scene) and the epilogue
//    evaluates evalDirect / directPdf on the light hit: lightF (TraceBase.cpp:279-284) or bsdfF (:316-320).
__global__ void __launch_bounds__(128) k_shadow(DScene sc, PathState st, const uint32_t *squeue, const uint32_t *scount, Counters *ctr) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    bool valid = i < *scount;
    uint32_t s = 0; bool mis = false, any = false; int li = -1;
    V3 p = v3s(0.0f), d = v3(0.0f, 0.0f, 1.0f); float tfar = INFINITY;
    if (valid) {
        uint32_t q = squeue[i]; s = q >> 2; mis = q & 1u; any = q & 2u;
        p = v3(st.px[s], st.py[s], st.pz[s]);
        d = mis ? v3(st.mdx[s], st.mdy[s], st.mdz[s]) : v3(st.ndx[s], st.ndy[s], st.ndz[s]);
        li = st.qlight[s];
        if (any) tfar = mis ? st.mpb[s] : st.ndist[s];
    }
    // both query kinds can share a warp (scenes mixing mesh lights with quad/environment lights): run the
    // occlusion traversal for the "any" lanes, then the closest-hit traversal for the rest
    bool anyhit = false;
    if (__any_sync(0xffffffffu, valid && any)) {
        Hit h = trace_scene<true>(sc, valid && any, p, d, 5e-4f, tfar, li);
        if (valid && any) {
            anyhit = h.id != HID_MISS;
            if (!anyhit) {
                if (!mis) { st.lx[s] = st.nfx[s]; st.ly[s] = st.nfy[s]; st.lz[s] = st.nfz[s]; }
                else { st.bx[s] = st.mwx[s]; st.by[s] = st.mwy[s]; st.bz[s] = st.mwz[s]; }
            }
        }
    }
    if (__any_sync(0xffffffffu, valid && !any)) {
        Hit h = trace_scene<false>(sc, valid && !any, p, d, 5e-4f, INFINITY, -1);
        if (valid && !any) {
            const DPrim &l = sc.prims[li];
            anyhit = h.id != HID_MISS;
            if (anyhit) {
                Surface ls; make_surface(sc, h, p, d, ls);
                bool visible = ls.prim == li;
                if (visible && !mis && h.t*(1.0f + 1e-3f) < st.ndist[s]) visible = false;       // TraceBase.cpp:160
                if (visible) {
                    V3 em = eval_direct(sc, ls);
                    if (!is_zero(em)) {
                        if (!mis) {
                            V3 f = v3(st.nfx[s], st.nfy[s], st.nfz[s]);
                            float pdfL = st.npl[s];
                            V3 lightF = (f*em)/pdfL;
                            lightF = lightF*power_heuristic(pdfL, st.npb[s]);
                            st.lx[s] = lightF.x; st.ly[s] = lightF.y; st.lz[s] = lightF.z;
                        } else {
                            // TriangleMesh::directPdf (TriangleMesh.cpp:477-481)
                            float directPdf = length_sq(p - ls.p)/(-dot(d, ls.Ng)*l.total_area);
                            V3 w = v3(st.mwx[s], st.mwy[s], st.mwz[s]);
                            V3 bsdfF = em*w;
                            bsdfF = bsdfF*power_heuristic(st.mpb[s], directPdf);
                            st.bx[s] = bsdfF.x; st.by[s] = bsdfF.y; st.bz[s] = bsdfF.z;
                        }
                    }
                }
            }
        }
    }
    count_rays<true>(ctr, valid, anyhit);
}

// Fold this bounce's direct light + surface emission into the path (order as in handleSurface:537-543),
// apply the NaN guards of traceSample (PathTracer.cpp:119-122,130) and compact the survivors.
__global__ void __launch_bounds__(256) k_accum(DScene sc, PathState st, const uint32_t *queue, const uint32_t *count,
                                               uint32_t *next_queue, uint32_t *next_count) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    bool valid = i < *count;
    bool alive = false; uint32_t s = 0;
    if (valid) {
        s = queue[i];
        uint32_t info = st.info[s];
        V3 em = v3(st.ex[s], st.ey[s], st.ez[s]);
        bool touched = false;
        if (info & F_HAS_NEE) {
            // generalizedShadowRay's `bounce >= minBounces` test (TraceBase.cpp:117) on bounce+1 of the shading bounce
            V3 L = v3(st.lx[s], st.ly[s], st.lz[s]), B = v3(st.bx[s], st.by[s], st.bz[s]);
            V3 r = ((L + B)*st.wl[s])*v3(st.ux[s], st.uy[s], st.uz[s]);
            em = em + r; touched = true;
        }
        if (info & F_HAS_SURF) { em = em + v3(st.sx[s], st.sy[s], st.sz[s]); touched = true; }
        alive = (info & F_ALIVE) != 0;
        bool finalCheck = (info & F_FINAL_CHECK) != 0;
        if (alive || finalCheck) {
            V3 thr = v3(st.tx[s], st.ty[s], st.tz[s]);
            bool bad = false;
            if (alive) {
                V3 o = v3(st.ox[s], st.oy[s], st.oz[s]), d = v3(st.dx[s], st.dy[s], st.dz[s]);
                bad = isnan(sum(d) + sum(o));
            }
            bad = bad || isnan(sum(thr) + sum(em));
            if (bad) { em = v3s(0.0f); alive = false; touched = true; }
        }
        if (alive && finalCheck) alive = false;                  // bounce reached maxBounces: loop exits
        if (touched) { st.ex[s] = em.x; st.ey[s] = em.y; st.ez[s] = em.z; }
        st.info[s] = (info & ~(F_HAS_NEE | F_HAS_SURF | F_ALIVE | F_FINAL_CHECK)) | (alive ? F_ALIVE : 0u);
    }
    unsigned m = __ballot_sync(0xffffffffu, alive);
    if (m) {
        unsigned lane = threadIdx.x & 31, base = 0;
        if (lane == 0) base = atomicAdd(next_count, unsigned(__popc(m)));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (alive) next_queue[base + __popc(m & ((1u << lane) - 1u))] = s;
    }
}

// OutputBuffer::addSample (cameras/OutputBuffer.hpp:104-132): running mean in sample order, NaN/Inf samples dropped
__global__ void __launch_bounds__(256) k_resolve(PathState st, BatchInfo bi, uint32_t spp_count, float *fb, uint32_t *fb_count) {
    uint32_t pix = blockIdx.x*blockDim.x + threadIdx.x;
    if (pix >= bi.n_pix) return;
    uint32_t pid = bi.pix_id[pix];
    float mx = fb[3*size_t(pid)], my = fb[3*size_t(pid) + 1], mz = fb[3*size_t(pid) + 2];
    uint32_t cnt = fb_count[pid];
    for (uint32_t k = 0; k < spp_count; ++k) {
        size_t s = size_t(k)*bi.n_pix + pix;
        float cx = st.ex[s], cy = st.ey[s], cz = st.ez[s];
        if (isnan(cx) || isnan(cy) || isnan(cz) || isinf(cx) || isinf(cy) || isinf(cz)) continue;
        float n = float(cnt + 1u); cnt++;
        mx += (cx - mx)/n; my += (cy - my)/n; mz += (cz - mz)/n;
    }
    fb[3*size_t(pid)] = mx; fb[3*size_t(pid) + 1] = my; fb[3*size_t(pid) + 2] = mz;
    fb_count[pid] = cnt;
}

// Tile-major pack / unpack of the resident framebuffer: the send/receive side of the one collective on
// this path (all-gather of the rendered tiles across GPUs).
__global__ void __launch_bounds__(256) k_pack_tiles(const uint32_t *pix_id, uint32_t n_pix, const float *fb, float *out) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    size_t p = size_t(pix_id[i])*3;
    out[3*size_t(i)] = fb[p]; out[3*size_t(i) + 1] = fb[p + 1]; out[3*size_t(i) + 2] = fb[p + 2];
}
__global__ void __launch_bounds__(256) k_unpack_tiles(const uint32_t *pix_id, uint32_t n_pix, const float *in, float *fb, uint32_t *fb_count, uint32_t count) {
    uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    size_t p = size_t(pix_id[i])*3;
    fb[p] = in[3*size_t(i)]; fb[p + 1] = in[3*size_t(i) + 1]; fb[p + 2] = in[3*size_t(i) + 2];
    fb_count[pix_id[i]] = count;
}

}  // namespace tgb
