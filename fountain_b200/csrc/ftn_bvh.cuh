// Flattened BVH layouts and the traversal loops.
//
// Layout "BVH2x64": one 64-byte, 64-byte-aligned record per interior node holding BOTH
// children's boxes (so one node fetch = 4 x 128-bit loads decides two subtrees) and
// triangles re-laid as 48-byte pre-gathered records in leaf order (3 x 128-bit loads, no
// index indirection; replaces the reference's Box<dyn Primitive> -> Arc<TriangleMesh> ->
// indices -> vertices chain, bvh.rs:181 / triangle.rs:96-111).
//
//   node:  n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//          n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//          nz = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//          ci = (child0, child1, unused, unused)   child >= 0: interior node index
//                                                  child <  0: leaf, ~child = first*4 + (count-1)
//   tri:   (p0.xyz, prim id), (p1.xyz, mesh id), (p2.xyz, unused)
//
// The box test is the reference's slab test with 1/d hoisted (bounds.rs:214-233): same
// roundings, same 1+2*gamma(3) widening, same NaN-ignoring min/max.
#pragma once
#include "ftn_geom.cuh"

namespace ftn {

#define FTN_LEAF_MAX 4
#define FTN_STACK_SIZE 64
#define FTN_NO_HIT_SLOT 0xFFFFFFFFu
#define FTN_SPHERE_SLOT_FLAG 0x80000000u

struct F4 { float x, y, z, w; };

struct BvhView {
    const F4* nodes;      // 4 x F4 per node
    const F4* tris;       // 3 x F4 per triangle, leaf order
    uint32_t n_nodes;     // 0 => no triangles
    uint32_t n_tris;
};

#if defined(__CUDA_ARCH__)
FTN_HD F4 ld4(const F4* p) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
}
#else
FTN_HD F4 ld4(const F4* p) { return *p; }
#endif

struct RaySlab { V3 o; V3 inv_d; float widen; };
FTN_HD RaySlab make_ray_slab(V3 o, V3 d) {
    RaySlab s; s.o = o;
    s.inv_d = V3(rn_div(1.0f, d.x), rn_div(1.0f, d.y), rn_div(1.0f, d.z));
    s.widen = rn_add(1.0f, rn_mul(2.0f, gamma_n(3)));
    return s;
}
// bounds.rs:214-233 for one box given as (lo,hi) per axis; returns hit and the entry distance.
// Branch-free: the reference returns early as soon as t0 > t1; t0 only grows and t1 only shrinks
// (f32::max/min ignore NaNs, like fmaxf/fminf), so testing once at the end accepts exactly the
// same boxes without divergent branches inside the warp.
FTN_HD bool slab_test(const RaySlab& r, float lox, float hix, float loy, float hiy, float loz, float hiz,
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

struct TraceCounters { uint32_t nodes, tris; };

// Closest hit (ANY = false: Scene::intersect, bvh.rs:160-215) or any hit (ANY = true:
// Scene::intersect_test, bvh.rs:217-266) against the triangle BVH.  `t_max` in/out.
// Returns the leaf-order slot of the accepted triangle or FTN_NO_HIT_SLOT.
template <bool ANY, bool COUNT>
FTN_HD uint32_t bvh2_traverse(const BvhView& bvh, V3 ro, V3 rd, float* t_max_io, TriHit* hit_out, TraceCounters* ctr) {
    uint32_t best = FTN_NO_HIT_SLOT;
    if (bvh.n_nodes == 0u) return best;
    float t_max = *t_max_io;
    const RaySlab slab = make_ray_slab(ro, rd);
    const RayShear shear = make_ray_shear(rd);
    int stack[FTN_STACK_SIZE];
    int sp = 0;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            const F4* n = bvh.nodes + 4 * (size_t)cur;
            const F4 n0 = ld4(n), n1 = ld4(n + 1), nz = ld4(n + 2), ci = ld4(n + 3);
            if (COUNT) ctr->nodes++;
            float e0, e1;
            const bool h0 = slab_test(slab, n0.x, n0.y, n0.z, n0.w, nz.x, nz.y, t_max, &e0);
            const bool h1 = slab_test(slab, n1.x, n1.y, n1.z, n1.w, nz.z, nz.w, t_max, &e1);
            const int c0 = (int)f2u(ci.x), c1 = (int)f2u(ci.y);
            if (h0 && h1) {
                // front-to-back by entry distance (the reference orders by split-axis sign, bvh.rs:194-201;
                // order only affects which of two exactly-tied hits is kept)
                if (e1 < e0) { if (sp < FTN_STACK_SIZE) stack[sp++] = c0; cur = c1; }
                else { if (sp < FTN_STACK_SIZE) stack[sp++] = c1; cur = c0; }
                continue;
            } else if (h0) { cur = c0; continue; }
            else if (h1) { cur = c1; continue; }
        } else {
            const uint32_t ref = ~(uint32_t)cur;
            const uint32_t first = ref >> 2, count = (ref & 3u) + 1u;
            for (uint32_t i = 0; i < count; ++i) {
                const F4* t = bvh.tris + 3 * (size_t)(first + i);
                const F4 a = ld4(t), b = ld4(t + 1), c = ld4(t + 2);
                if (COUNT) ctr->tris++;
                TriHit h;
                if (triangle_intersect(V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), ro, shear, t_max, &h)) {
                    t_max = h.t; best = first + i; *hit_out = h;
                    if (ANY) { *t_max_io = t_max; return best; }
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    *t_max_io = t_max;
    return best;
}

}  // namespace ftn
