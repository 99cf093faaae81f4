// Flattened BVH layouts and the single-ray traversal loop.
//
// Triangles are re-laid as 64-byte pre-gathered records in leaf order (one 256-bit + one 128-bit
// load: the traversal is bound by load INSTRUCTIONS per lane, not bytes, so the padded record costs
// two L1 wavefronts per lane instead of the three of a packed 48-byte one; no index indirection --
// replaces the reference's Box<dyn Primitive> -> Arc<TriangleMesh> -> indices -> vertices chain,
// bvh.rs:181 / triangle.rs:96-111):
//   tri:   (p0.xyz, prim id), (p1.xyz, mesh id), (p2.xyz, unused), (unused)
//
// Interior nodes, FTN_BVH_WIDTH = 2 (default) -- "BVH2x64": 64-byte record with both children's boxes:
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)  n1 = (c1...)  nz = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//   ci = (child0, child1, -, -)
// FTN_BVH_WIDTH = 4 -- "BVH4x128" (built with -DFTN_BVH_WIDTH=4; measured SLOWER on B200, kept for
// A/B): one 128-byte record holding the boxes of up to four children in SoA form.  It halves the
// node visits (C3 interior rays: 35.3 -> 18.3 per ray) but the four exact slab tests + the sorting
// network cost more issue slots than the saved round trips buy back: 1866 vs 2049 Mrays/s
// (profiles/r01_ab_bvh4_vs_bvh2.txt).  The traversal is issue-bound, not latency-bound.
//   q0 = lo.x[0..3]  q1 = hi.x[0..3]  q2 = lo.y[..]  q3 = hi.y[..]  q4 = lo.z[..]  q5 = hi.z[..]
//   q6 = child[0..3] (int bits)       q7 = unused
//   unused child slots carry child = FTN_TRAVERSAL_DONE and are masked out of the box test.
// child >= 0: interior node index;  child < 0: leaf, ~child = first*4 + (count-1), count <= 4.
//
// The box test is the reference's slab test with 1/d hoisted (bounds.rs:214-233): same
// roundings, same 1+2*gamma(3) widening, same NaN-ignoring min/max.
#pragma once
#include "ftn_geom.cuh"

namespace ftn {

#ifndef FTN_BVH_WIDTH
#define FTN_BVH_WIDTH 2
#endif
#if FTN_BVH_WIDTH == 4
#define FTN_NODE_F4 8
#else
#define FTN_NODE_F4 4
#endif
#define FTN_NODE_BYTES (16 * FTN_NODE_F4)
#define FTN_TRI_F4 4
#define FTN_TRI_BYTES (16 * FTN_TRI_F4)

#ifndef FTN_LEAF_MAX
#define FTN_LEAF_MAX 4
#endif
#define FTN_STACK_SIZE 256  /* traversal stack (local memory, touched only as deep as the tree): radix trees are <= 62 deep, PLOC trees reach 175 at 50M triangles */
#define FTN_NO_HIT_SLOT 0xFFFFFFFFu
#define FTN_SPHERE_SLOT_FLAG 0x80000000u
#define FTN_TRAVERSAL_DONE ((int)0x80000000)

struct F4 { float x, y, z, w; };

struct BvhView {
    const F4* nodes;      // FTN_NODE_F4 x F4 per node
    const F4* tris;       // FTN_TRI_F4 x F4 per triangle, leaf order
    uint32_t n_nodes;     // 0 => no triangles
    uint32_t n_tris;
    uint32_t wide;        // 0: BVH2x64 records (this file); 1: BVH8q compressed 8-wide records (ftn_bvh8.cuh)
};

#ifndef FTN_PREFETCH_CHILDREN
#define FTN_PREFETCH_CHILDREN 0
#endif
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
FTN_HD F4 ld4(const F4* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}
// One 256-bit read-only load (LDG.E.256, sm_100+) of two consecutive F4; p must be 32-byte aligned.
// ncu on the 128-bit version showed the traversal kernels bound by the L1 data pipe
// (l1tex__data_pipe_lsu_wavefronts 90 % of peak): every lane reads its own node, so each load
// instruction costs one L1 wavefront PER LANE however few bytes it moves -- halving the number of
// load instructions per node halves that traffic.
__device__ __forceinline__ void ld8(const F4* p, F4& a, F4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
#else
FTN_HD F4 ld4(const F4* p) { return *p; }
FTN_HD void ld8(const F4* p, F4& a, F4& b) { a = p[0]; b = p[1]; }
#endif

// `nan_free`: every component of 1/d is finite and non-zero, so (bound - o) * (1/d) can never be
// 0 * inf = NaN and the cheaper, bit-identical form of the slab test below applies.
struct RaySlab { V3 o; V3 inv_d; float widen; bool nan_free; };
FTN_HD bool slab_component_regular(float v) { return v != 0.0f && fabsf(v) <= 3.402823466e+38f; }   // finite, non-zero, not NaN
FTN_HD RaySlab make_ray_slab(V3 o, V3 d) {
    RaySlab s; s.o = o;
    s.inv_d = V3(rn_div(1.0f, d.x), rn_div(1.0f, d.y), rn_div(1.0f, d.z));
    s.widen = rn_add(1.0f, rn_mul(2.0f, gamma_n(3)));
    s.nan_free = slab_component_regular(s.inv_d.x) && slab_component_regular(s.inv_d.y) && slab_component_regular(s.inv_d.z);
    return s;
}
// bounds.rs:214-233 for one box given as (lo,hi) per axis; returns hit and the entry distance.
// Branch-free: the reference returns early as soon as t0 > t1; t0 only grows and t1 only shrinks
// (f32::max/min ignore NaNs, like fmaxf/fminf), so testing once at the end accepts exactly the
// same boxes without divergent branches inside the warp.
FTN_HD bool slab_test_exact(const RaySlab& r, float lox, float hix, float loy, float hiy, float loz, float hiz,
                            float t_max, float* t_entry) {
    const float ax = rn_mul(rn_sub(lox, r.o.x), r.inv_d.x), bx = rn_mul(rn_sub(hix, r.o.x), r.inv_d.x);
    const float ay = rn_mul(rn_sub(loy, r.o.y), r.inv_d.y), by = rn_mul(rn_sub(hiy, r.o.y), r.inv_d.y);
    const float az = rn_mul(rn_sub(loz, r.o.z), r.inv_d.z), bz = rn_mul(rn_sub(hiz, r.o.z), r.inv_d.z);
    // `if t_near > t_far { swap }`: near = (a > b) ? b : a, far = (a > b) ? a : b  (NaN: no swap)
    const bool sx = ax > bx, sy = ay > by, sz = az > bz;
    const float nx = sx ? bx : ax, fx = rn_mul(sx ? ax : bx, r.widen);
    const float ny = sy ? by : ay, fy = rn_mul(sy ? ay : by, r.widen);
    const float nz = sz ? bz : az, fz = rn_mul(sz ? az : bz, r.widen);
    const float t0 = fmaxf(fmaxf(fmaxf(0.0f, nx), ny), nz);
    const float t1 = fminf(fminf(fminf(t_max, fx), fy), fz);
    *t_entry = t0;
    return !(t0 > t1);
}
// The same test for a ray with r.nan_free: without NaNs the swap is min/max, and because
// x -> round(x * widen) is monotone (widen > 0) the three widened far distances share one multiply:
// min(fx*w, fy*w, fz*w) == min(fx, fy, fz)*w bit for bit.  Identical accept/reject and entry
// distance as slab_test_exact (tests/test_hostsim_parity.py checks both forms against the oracle).
FTN_HD bool slab_test_nan_free(const RaySlab& r, float lox, float hix, float loy, float hiy, float loz, float hiz,
                               float t_max, float* t_entry) {
    const float ax = rn_mul(rn_sub(lox, r.o.x), r.inv_d.x), bx = rn_mul(rn_sub(hix, r.o.x), r.inv_d.x);
    const float ay = rn_mul(rn_sub(loy, r.o.y), r.inv_d.y), by = rn_mul(rn_sub(hiy, r.o.y), r.inv_d.y);
    const float az = rn_mul(rn_sub(loz, r.o.z), r.inv_d.z), bz = rn_mul(rn_sub(hiz, r.o.z), r.inv_d.z);
    const float t0 = fmaxf(fmaxf(fmaxf(0.0f, fminf(ax, bx)), fminf(ay, by)), fminf(az, bz));
    const float far_ = rn_mul(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), r.widen);
    const float t1 = fminf(t_max, far_);
    *t_entry = t0;
    return !(t0 > t1);
}
template <bool NAN_FREE>
FTN_HD bool slab_test(const RaySlab& r, float lox, float hix, float loy, float hiy, float loz, float hiz,
                      float t_max, float* t_entry) {
    return NAN_FREE ? slab_test_nan_free(r, lox, hix, loy, hiy, loz, hiz, t_max, t_entry)
                    : slab_test_exact(r, lox, hix, loy, hiy, loz, hiz, t_max, t_entry);
}

struct TraceCounters { uint32_t nodes, tris; };

// the three rows of triangle record `slot` (256-bit + 128-bit load)
FTN_HD void load_tri(const BvhView& bvh, uint32_t slot, F4* a, F4* b, F4* c) {
    const F4* t = bvh.tris + (size_t)FTN_TRI_F4 * (size_t)slot;
    ld8(t, *a, *b);
    *c = ld4(t + 2);
}

FTN_HD void cswap(float& ka, int& va, float& kb, int& vb) {   // compare-exchange on (key, value)
    const bool s = kb < ka;
    const float k0 = s ? kb : ka, k1 = s ? ka : kb;
    const int v0 = s ? vb : va, v1 = s ? va : vb;
    ka = k0; va = v0; kb = k1; vb = v1;
}

// Tests the children of interior node `cur` and returns the next reference to visit: the nearest
// entered child (the others are pushed far-to-near), or a popped entry, or FTN_TRAVERSAL_DONE.
// Front-to-back by entry distance (the reference orders by split-axis sign, bvh.rs:194-201; the
// order only decides which of two exactly tied hits is kept).
template <bool NAN_FREE>
FTN_HD int node_step_impl(const BvhView& bvh, int cur, const RaySlab& slab, float t_max, int* stack, int& sp) {
    const F4* nd = bvh.nodes + (size_t)FTN_NODE_F4 * (size_t)cur;
#if FTN_BVH_WIDTH == 4
    const F4 lx = ld4(nd), hx = ld4(nd + 1), ly = ld4(nd + 2), hy = ld4(nd + 3), lz = ld4(nd + 4), hz = ld4(nd + 5), cr = ld4(nd + 6);
    const float inf = FTN_INF;
    float e0, e1, e2, e3;
    int c0 = (int)f2u(cr.x), c1 = (int)f2u(cr.y), c2 = (int)f2u(cr.z), c3 = (int)f2u(cr.w);
    // an unused slot is marked by its child reference (the swap-on-inverted rule of the slab test
    // would otherwise read an inverted or NaN box as an infinite one)
    const bool h0 = slab_test<NAN_FREE>(slab, lx.x, hx.x, ly.x, hy.x, lz.x, hz.x, t_max, &e0) && c0 != FTN_TRAVERSAL_DONE;
    const bool h1 = slab_test<NAN_FREE>(slab, lx.y, hx.y, ly.y, hy.y, lz.y, hz.y, t_max, &e1) && c1 != FTN_TRAVERSAL_DONE;
    const bool h2 = slab_test<NAN_FREE>(slab, lx.z, hx.z, ly.z, hy.z, lz.z, hz.z, t_max, &e2) && c2 != FTN_TRAVERSAL_DONE;
    const bool h3 = slab_test<NAN_FREE>(slab, lx.w, hx.w, ly.w, hy.w, lz.w, hz.w, t_max, &e3) && c3 != FTN_TRAVERSAL_DONE;
    e0 = h0 ? e0 : inf; e1 = h1 ? e1 : inf; e2 = h2 ? e2 : inf; e3 = h3 ? e3 : inf;
    // 5-comparator sorting network on (entry distance, child); misses sort to the back with +inf.
    // Entry distances are >= 0 and finite for entered boxes, so +inf marks exactly the misses.
    cswap(e0, c0, e1, c1); cswap(e2, c2, e3, c3); cswap(e0, c0, e2, c2); cswap(e1, c1, e3, c3); cswap(e1, c1, e2, c2);
    const int nh = (int)h0 + (int)h1 + (int)h2 + (int)h3;
    if (nh == 0) return (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE;
    if (nh > 3) stack[sp++] = c3;
    if (nh > 2) stack[sp++] = c2;
    if (nh > 1) stack[sp++] = c1;
    return c0;
#else
    F4 n0, n1, nz, ci;
    ld8(nd, n0, n1); ld8(nd + 2, nz, ci);
    float e0, e1;
    const int c0 = (int)f2u(ci.x), c1 = (int)f2u(ci.y);
#if defined(__CUDA_ARCH__) && FTN_PREFETCH_CHILDREN
    // the next record this ray reads is one of the two children: start both L2 -> L1 fetches now,
    // under the ~60 instructions of the two slab tests, instead of after them
    prefetch_l1(c0 >= 0 ? (const void*)(bvh.nodes + (size_t)FTN_NODE_F4 * (size_t)c0)
                        : (FTN_PREFETCH_CHILDREN > 1 ? (const void*)(bvh.tris + FTN_TRI_F4 * (size_t)((~(uint32_t)c0) >> 2)) : (const void*)nd));
    if (c1 != FTN_TRAVERSAL_DONE)
        prefetch_l1(c1 >= 0 ? (const void*)(bvh.nodes + (size_t)FTN_NODE_F4 * (size_t)c1)
                            : (FTN_PREFETCH_CHILDREN > 1 ? (const void*)(bvh.tris + FTN_TRI_F4 * (size_t)((~(uint32_t)c1) >> 2)) : (const void*)nd));
#endif
    const bool h0 = slab_test<NAN_FREE>(slab, n0.x, n0.y, n0.z, n0.w, nz.x, nz.y, t_max, &e0);
    const bool h1 = slab_test<NAN_FREE>(slab, n1.x, n1.y, n1.z, n1.w, nz.z, nz.w, t_max, &e1) && c1 != FTN_TRAVERSAL_DONE;
    if (h0 && h1) {
        const bool swap = e1 < e0;
        stack[sp++] = swap ? c0 : c1;
        return swap ? c1 : c0;
    }
    if (h0) return c0;
    if (h1) return c1;
    return (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE;
#endif
}
FTN_HD int node_step(const BvhView& bvh, int cur, const RaySlab& slab, float t_max, int* stack, int& sp) {
    // per-ray choice; rays with a zero / denormal direction component are rare, so warps seldom diverge here
    return slab.nan_free ? node_step_impl<true>(bvh, cur, slab, t_max, stack, sp) : node_step_impl<false>(bvh, cur, slab, t_max, stack, sp);
}

// Tests the triangles of leaf reference `leaf` (< 0); returns true if ANY and a hit was accepted.
template <bool ANY, bool COUNT>
FTN_HD bool leaf_step(const BvhView& bvh, int leaf, V3 ro, const RayShear& shear, float* t_max, uint32_t* best, TriHit* hit, TraceCounters* ctr) {
    const uint32_t ref = ~(uint32_t)leaf;
    const uint32_t first = ref >> 2, count = (ref & 3u) + 1u;
    for (uint32_t i = 0; i < count; ++i) {
        const F4* t = bvh.tris + (size_t)FTN_TRI_F4 * (size_t)(first + i);
        F4 a, b; ld8(t, a, b);
        const F4 c = ld4(t + 2);
        if (COUNT) ctr->tris++;
        TriHit h;
        if (triangle_intersect(V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), ro, shear, *t_max, &h)) {
            *t_max = h.t; *best = first + i; *hit = h;
            if (ANY) return true;
        }
    }
    return false;
}

// Closest hit (ANY = false: Scene::intersect, bvh.rs:160-215) or any hit (ANY = true:
// Scene::intersect_test, bvh.rs:217-266) against the triangle BVH, one ray.  `t_max` in/out.
// Returns the leaf-order slot of the accepted triangle or FTN_NO_HIT_SLOT.  This is the plain
// statement of the traversal (and what the host test harness runs); the kernels use the
// warp-persistent form of the same steps in ftn_trace_persistent.cuh.
template <bool ANY, bool COUNT>
FTN_HD uint32_t bvh_traverse(const BvhView& bvh, V3 ro, V3 rd, float* t_max_io, TriHit* hit_out, TraceCounters* ctr) {
    uint32_t best = FTN_NO_HIT_SLOT;
    if (bvh.n_nodes == 0u) return best;
    float t_max = *t_max_io;
    const RaySlab slab = make_ray_slab(ro, rd);
    const RayShear shear = make_ray_shear(rd);
    int stack[FTN_STACK_SIZE];
    int sp = 0;
    int cur = 0;
    while (cur != FTN_TRAVERSAL_DONE) {
        if (cur >= 0) {
            if (COUNT) ctr->nodes++;
            cur = node_step(bvh, cur, slab, t_max, stack, sp);
        } else {
            if (leaf_step<ANY, COUNT>(bvh, cur, ro, shear, &t_max, &best, hit_out, ctr)) break;
            cur = (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE;
        }
    }
    *t_max_io = t_max;
    return best;
}

}  // namespace ftn
