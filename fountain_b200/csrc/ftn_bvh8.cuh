// "BVH8q": the compressed 8-wide layout of the aggregate and its traversal steps.
//
// What it replaces: BVH::intersect / intersect_test (bvh.rs:160-266) over the reference's 32-byte
// LinearBVHNode (bvh.rs:269-302), and the box test Bounds3f::intersect_test (bounds.rs:214-233).
//
// Why: ncu on the 64-byte two-child records (profiles/r02_ncu_c3_batches_baseline.txt) shows the
// traversal bound by the L1 data pipe -- one wavefront per LANE per load instruction, because every
// lane reads its own record -- at 36 node visits x 2 loads + 3.8 triangles x 3 loads per incoherent
// ray.  An 8-wide node with 8-bit child boxes (Ylitie, Karras, Laine 2017) cuts the visits to about a
// third; here it is laid out for sm_100's 256-bit loads:
//
//   node, 96-byte stride, 32-byte aligned, 80 bytes used = LDG.256 + LDG.256 + LDG.128:
//     w0 = (p.x, p.y, p.z, ex | ey<<8 | ez<<16 | imask<<24)      quantisation frame: origin p (= box min, exact),
//                                                                per-axis scale 2^(e-127); imask: slot holds an interior child
//     w1 = (child_base, tri_base, counts16 | lmask<<16, 0)       interior children are consecutive records from child_base in slot
//                                                                order; leaf children own count(slot) = 2 bits of counts16 (1..3)
//                                                                consecutive triangles from tri_base in slot order; lmask: slot is a leaf
//     w2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])  child box planes, one byte each: plane = p + q * 2^(e-127)
//     w3 = (qlo.z[..],   qlo.z[..],   qhi.x[..],   qhi.x[..])    (lo rounded down, hi rounded up: the decoded box contains the child's)
//     w4 = (qhi.y[..],   qhi.y[..],   qhi.z[..],   qhi.z[..])
//   Child slots are assigned by octant (slot bit a set = the child lies towards +axis a of the node's centre), so a ray
//   visits the hit children in the order slot ^ octant-of-its-direction without sorting distances (Ylitie et al. section 3.3).
//
// The box test is CONSERVATIVE, not the reference's arithmetic: t = fma(q', 2^(e+15) / d, ((p - o) / d - 2^(e+15) / d) -+ err)
// with q' = 1 + q * 2^-15 built by one PRMT, and err = 20 eps |(p-o)/d| + 6 eps |2^(e+15)/d| bounding every rounding of this
// form AND of the reference's ((lo - o) * (1/d), far side widened by 1 + 2 gamma(3)), so every box the reference's slab test
// enters on the exact child bounds is entered here too (the decoded box is a superset, the interval only grows).  A
// direction component below 2^-60 (the reference gets +-inf or a huge 1/d) uses +-2^60: the slab then constrains nothing
// whenever the origin is inside it, as in the reference (its NaN / inf cases).  Hits are decided by the untouched exact
// watertight triangle test, so results do not depend on this test beyond "never cull a box the reference enters".
// tests/test_hostsim_parity.py checks the superset property against the exact slab test on random and degenerate rays.
#pragma once
#include "ftn_bvh.cuh"

namespace ftn {

// FTN_BVH8_PLANES16 = 1 (A/B build, measured SLOWER, kept like the BVH4 experiment of round 1): the child planes as 16-bit
// bf16 INTEGERS (0..255, exact) instead of bytes, two children per word -- 128-byte records (4 x LDG.256).  Idea: ncu on the
// byte layout shows the node test bound by the ALU pipe (69 % busy, FMA pipe 28 %): per child 6 PRMT byte->float conversions
// + 4 min/max + compare + add on the ALU pipe against 6 FFMA.  A bf16 in the LOW half of a word becomes a float with one
// shift (IMAD.SHL: FMA pipe); one in the HIGH half is used as it is -- the low half then adds < 1 quantisation step, which is
// conservative for hi planes, and lo planes of odd children are stored one step lower at build time.  SASS: PRMT 57 -> 0,
// IMAD 108 -> 114, registers 72 -> 80 (6 blocks / SM instead of 7).  B200 (profiles/r02_ab_bvh8.txt): coherent +3.5 %,
// incoherent diffuse +0 %, interior -5 %, C4 k_extend -5.5 %: the larger records and the lost block cost what the ALU pipe gains.
#ifndef FTN_BVH8_PLANES16
#define FTN_BVH8_PLANES16 0
#endif
#if FTN_BVH8_PLANES16
#define FTN_NODE8_F4 8
#else
#define FTN_NODE8_F4 6
#endif
#define FTN_NODE8_BYTES (16 * FTN_NODE8_F4)
#define FTN_LEAF8_MAX 3            /* triangles per leaf child (2-bit count) */
#define FTN_STACK8_SHARED 8        /* traversal stack entries per lane kept in shared memory */
#define FTN_STACK8_SIZE 256        /* total entries = FTN_STACK_SIZE: one per level of the tree at most, and the wide tree is never deeper than the binary tree it was collapsed from; entries beyond the shared ones live in local memory and are touched only that deep */

#if defined(__CUDA_ARCH__)
FTN_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
FTN_HD int popc32(uint32_t x) { return __popc(x); }
FTN_HD int bfind8(uint32_t x) { return 31 - __clz((int)x); }       // x != 0
#else
FTN_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7u))) & 0xFFu) << (8 * i);
    return r;
}
FTN_HD int popc32(uint32_t x) { return __builtin_popcount(x); }
FTN_HD int bfind8(uint32_t x) { return 31 - __builtin_clz(x); }
#endif

// ---- per-ray constants of the conservative box test ---------------------------------------------------------------
struct Ray8 { V3 o; V3 idir; uint32_t octinv; };
FTN_HD float ray8_inv(float d) {
    const float big = 1152921504606846976.0f;                       // 2^60
    return (fabsf(d) >= 8.673617379884035e-19f) ? rn_div(1.0f, d)   // |d| >= 2^-60
                                                : (sign_positive(d) ? big : -big);
}
FTN_HD Ray8 make_ray8(V3 o, V3 d) {
    Ray8 r; r.o = o;
    r.idir = V3(ray8_inv(d.x), ray8_inv(d.y), ray8_inv(d.z));
    const uint32_t oct = (r.idir.x < 0.0f ? 1u : 0u) | (r.idir.y < 0.0f ? 2u : 0u) | (r.idir.z < 0.0f ? 4u : 0u);
    r.octinv = oct ^ 7u;
    return r;
}

// slot -> visiting priority: mask'[j] = mask[j ^ octinv] for both bytes of m (inner hits | leaf hits << 8)
FTN_HD uint32_t bvh8_permute_hits(uint32_t m, uint32_t octinv) {
    if (octinv & 1u) m = ((m & 0x5555u) << 1) | ((m >> 1) & 0x5555u);
    if (octinv & 2u) m = ((m & 0x3333u) << 2) | ((m >> 2) & 0x3333u);
    if (octinv & 4u) m = ((m & 0x0F0Fu) << 4) | ((m >> 4) & 0x0F0Fu);
    return m;
}

struct Node8Hits {
    uint32_t child_base, ng_bits;   // ng_bits = permuted inner hits (8) | imask << 8
    uint32_t tri_base, tg_bits;     // tg_bits = permuted leaf hits (8) | counts16 << 8
};

#define FTN_BOX8_ERR_B (20.0f * FTN_MACHINE_EPS)
#define FTN_BOX8_ERR_A (6.0f * FTN_MACHINE_EPS)

// one child: bit = the ray's [0, t_max] overlaps the decoded box (conservatively)
#define FTN_BOX8_CHILD(k, nxw, nyw, nzw, fxw, fyw, fzw)                                                            \
    {                                                                                                               \
        const uint32_t sel = 0x7604u | (((k) & 3u) << 4);                                                           \
        const float tnx = fmaf(u2f(byte_perm(nxw, 0x3F800000u, sel)), ax, onx), tfx = fmaf(u2f(byte_perm(fxw, 0x3F800000u, sel)), ax, ofx); \
        const float tny = fmaf(u2f(byte_perm(nyw, 0x3F800000u, sel)), ay, ony), tfy = fmaf(u2f(byte_perm(fyw, 0x3F800000u, sel)), ay, ofy); \
        const float tnz = fmaf(u2f(byte_perm(nzw, 0x3F800000u, sel)), az, onz), tfz = fmaf(u2f(byte_perm(fzw, 0x3F800000u, sel)), az, ofz); \
        const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));                                                  \
        const float tf = fminf(fminf(tfx, tfy), fminf(tfz, t_max));                                                 \
        if (!(tn > tf)) hits |= 1u << (k);                                                                          \
    }

#if FTN_BVH8_PLANES16
#define FTN_BOX16_ERR_B (20.0f * FTN_MACHINE_EPS)
#define FTN_BOX16_ERR_A (20.0f * 256.0f * FTN_MACHINE_EPS)
#if defined(__CUDA_ARCH__)
// word << 16 as a multiply: ptxas turns it into IMAD.SHL on the FMA pipe (a plain shift goes to the saturated ALU pipe)
FTN_HD float bf16_lo(uint32_t w) { uint32_t r; asm("mul.lo.u32 %0, %1, 65536;" : "=r"(r) : "r"(w)); return __uint_as_float(r); }
#else
FTN_HD float bf16_lo(uint32_t w) { return u2f(w << 16); }
#endif
FTN_HD float bf16_hi(uint32_t w) { return u2f(w); }       // + the low half as mantissa garbage: < 1 unit above the stored integer
// children 2j (low halves) and 2j + 1 (high halves) of word j of each selected plane vector
#define FTN_BOX16_PAIR(j, NX, NY, NZ, FX, FY, FZ)                                                                   \
    {                                                                                                               \
        const float tnx0 = fmaf(bf16_lo(NX), ax, onx), tfx0 = fmaf(bf16_lo(FX), ax, ofx);                           \
        const float tny0 = fmaf(bf16_lo(NY), ay, ony), tfy0 = fmaf(bf16_lo(FY), ay, ofy);                           \
        const float tnz0 = fmaf(bf16_lo(NZ), az, onz), tfz0 = fmaf(bf16_lo(FZ), az, ofz);                           \
        const float tnx1 = fmaf(bf16_hi(NX), ax, onx), tfx1 = fmaf(bf16_hi(FX), ax, ofx);                           \
        const float tny1 = fmaf(bf16_hi(NY), ay, ony), tfy1 = fmaf(bf16_hi(FY), ay, ofy);                           \
        const float tnz1 = fmaf(bf16_hi(NZ), az, onz), tfz1 = fmaf(bf16_hi(FZ), az, ofz);                           \
        const float tn0 = fmaxf(fmaxf(tnx0, tny0), fmaxf(tnz0, 0.0f)), tf0 = fminf(fminf(tfx0, tfy0), fminf(tfz0, t_max)); \
        const float tn1 = fmaxf(fmaxf(tnx1, tny1), fmaxf(tnz1, 0.0f)), tf1 = fminf(fminf(tfx1, tfy1), fminf(tfz1, t_max)); \
        if (!(tn0 > tf0)) hits |= 1u << (2 * (j));                                                                  \
        if (!(tn1 > tf1)) hits |= 1u << (2 * (j) + 1);                                                              \
    }
FTN_HD Node8Hits node8_test(const F4* nodes, uint32_t idx, const Ray8& r, float t_max) {
    const F4* nd = nodes + (size_t)FTN_NODE8_F4 * (size_t)idx;
    F4 w0, w1, lx, ly, lz, hx, hy, hz;
    ld8(nd, w0, w1); ld8(nd + 2, lx, ly); ld8(nd + 4, lz, hx); ld8(nd + 6, hy, hz);
    const uint32_t eb = f2u(w0.w);
    // t = q * a + b with a = 2^e / d (exact: a power of two times 1/d), b = (p - o) / d, q the stored integer
    const float ax = rn_mul(u2f((eb & 0xFFu) << 23), r.idir.x);
    const float ay = rn_mul(u2f(((eb >> 8) & 0xFFu) << 23), r.idir.y);
    const float az = rn_mul(u2f(((eb >> 16) & 0xFFu) << 23), r.idir.z);
    const float bx = rn_mul(rn_sub(w0.x, r.o.x), r.idir.x), by = rn_mul(rn_sub(w0.y, r.o.y), r.idir.y), bz = rn_mul(rn_sub(w0.z, r.o.z), r.idir.z);
    const float ex = fmaf(FTN_BOX16_ERR_B, fabsf(bx), FTN_BOX16_ERR_A * fabsf(ax));
    const float ey = fmaf(FTN_BOX16_ERR_B, fabsf(by), FTN_BOX16_ERR_A * fabsf(ay));
    const float ez = fmaf(FTN_BOX16_ERR_B, fabsf(bz), FTN_BOX16_ERR_A * fabsf(az));
    const float onx = rn_sub(bx, ex), ofx = rn_add(bx, ex), ony = rn_sub(by, ey), ofy = rn_add(by, ey), onz = rn_sub(bz, ez), ofz = rn_add(bz, ez);
    const bool px = !(r.idir.x < 0.0f), py = !(r.idir.y < 0.0f), pz = !(r.idir.z < 0.0f);
    uint32_t hits = 0u;
#define FTN_SEL(p, a, b) ((p) ? f2u(a) : f2u(b))
    FTN_BOX16_PAIR(0, FTN_SEL(px, lx.x, hx.x), FTN_SEL(py, ly.x, hy.x), FTN_SEL(pz, lz.x, hz.x), FTN_SEL(px, hx.x, lx.x), FTN_SEL(py, hy.x, ly.x), FTN_SEL(pz, hz.x, lz.x))
    FTN_BOX16_PAIR(1, FTN_SEL(px, lx.y, hx.y), FTN_SEL(py, ly.y, hy.y), FTN_SEL(pz, lz.y, hz.y), FTN_SEL(px, hx.y, lx.y), FTN_SEL(py, hy.y, ly.y), FTN_SEL(pz, hz.y, lz.y))
    FTN_BOX16_PAIR(2, FTN_SEL(px, lx.z, hx.z), FTN_SEL(py, ly.z, hy.z), FTN_SEL(pz, lz.z, hz.z), FTN_SEL(px, hx.z, lx.z), FTN_SEL(py, hy.z, ly.z), FTN_SEL(pz, hz.z, lz.z))
    FTN_BOX16_PAIR(3, FTN_SEL(px, lx.w, hx.w), FTN_SEL(py, ly.w, hy.w), FTN_SEL(pz, lz.w, hz.w), FTN_SEL(px, hx.w, lx.w), FTN_SEL(py, hy.w, ly.w), FTN_SEL(pz, hz.w, lz.w))
#undef FTN_SEL
    const uint32_t imask = eb >> 24, meta = f2u(w1.z), lmask = (meta >> 16) & 0xFFu;
    const uint32_t m = bvh8_permute_hits((hits & imask) | ((hits & lmask) << 8), r.octinv);
    Node8Hits h;
    h.child_base = f2u(w1.x); h.ng_bits = (m & 0xFFu) | (imask << 8);
    h.tri_base = f2u(w1.y);   h.tg_bits = (m >> 8) | (meta << 8);
    return h;
}
#else
// Tests the eight child boxes of node `idx`; returns the hit children as the two groups the traversal carries.
FTN_HD Node8Hits node8_test(const F4* nodes, uint32_t idx, const Ray8& r, float t_max) {
    const F4* nd = nodes + (size_t)FTN_NODE8_F4 * (size_t)idx;
    F4 w0, w1, w2, w3;
    ld8(nd, w0, w1); ld8(nd + 2, w2, w3);
    const F4 w4 = ld4(nd + 4);
    const uint32_t eb = f2u(w0.w);
    // a = 2^(e+15) / d (exact: a power of two times 1/d);  b = (p - o) / d
    const float ax = rn_mul(u2f(((eb & 0xFFu) + 15u) << 23), r.idir.x);
    const float ay = rn_mul(u2f((((eb >> 8) & 0xFFu) + 15u) << 23), r.idir.y);
    const float az = rn_mul(u2f((((eb >> 16) & 0xFFu) + 15u) << 23), r.idir.z);
    const float bx = rn_mul(rn_sub(w0.x, r.o.x), r.idir.x), by = rn_mul(rn_sub(w0.y, r.o.y), r.idir.y), bz = rn_mul(rn_sub(w0.z, r.o.z), r.idir.z);
    const float ex = fmaf(FTN_BOX8_ERR_B, fabsf(bx), FTN_BOX8_ERR_A * fabsf(ax));
    const float ey = fmaf(FTN_BOX8_ERR_B, fabsf(by), FTN_BOX8_ERR_A * fabsf(ay));
    const float ez = fmaf(FTN_BOX8_ERR_B, fabsf(bz), FTN_BOX8_ERR_A * fabsf(az));
    const float cx = rn_sub(bx, ax), cy = rn_sub(by, ay), cz = rn_sub(bz, az);
    const float onx = rn_sub(cx, ex), ofx = rn_add(cx, ex), ony = rn_sub(cy, ey), ofy = rn_add(cy, ey), onz = rn_sub(cz, ez), ofz = rn_add(cz, ez);
    // near / far plane bytes by the sign of the direction
    const bool px = !(r.idir.x < 0.0f), py = !(r.idir.y < 0.0f), pz = !(r.idir.z < 0.0f);
    const uint32_t lx0 = f2u(w2.x), lx1 = f2u(w2.y), ly0 = f2u(w2.z), ly1 = f2u(w2.w), lz0 = f2u(w3.x), lz1 = f2u(w3.y);
    const uint32_t hx0 = f2u(w3.z), hx1 = f2u(w3.w), hy0 = f2u(w4.x), hy1 = f2u(w4.y), hz0 = f2u(w4.z), hz1 = f2u(w4.w);
    const uint32_t nx0 = px ? lx0 : hx0, nx1 = px ? lx1 : hx1, fx0 = px ? hx0 : lx0, fx1 = px ? hx1 : lx1;
    const uint32_t ny0 = py ? ly0 : hy0, ny1 = py ? ly1 : hy1, fy0 = py ? hy0 : ly0, fy1 = py ? hy1 : ly1;
    const uint32_t nz0 = pz ? lz0 : hz0, nz1 = pz ? lz1 : hz1, fz0 = pz ? hz0 : lz0, fz1 = pz ? hz1 : lz1;
    uint32_t hits = 0u;
    FTN_BOX8_CHILD(0u, nx0, ny0, nz0, fx0, fy0, fz0)
    FTN_BOX8_CHILD(1u, nx0, ny0, nz0, fx0, fy0, fz0)
    FTN_BOX8_CHILD(2u, nx0, ny0, nz0, fx0, fy0, fz0)
    FTN_BOX8_CHILD(3u, nx0, ny0, nz0, fx0, fy0, fz0)
    FTN_BOX8_CHILD(4u, nx1, ny1, nz1, fx1, fy1, fz1)
    FTN_BOX8_CHILD(5u, nx1, ny1, nz1, fx1, fy1, fz1)
    FTN_BOX8_CHILD(6u, nx1, ny1, nz1, fx1, fy1, fz1)
    FTN_BOX8_CHILD(7u, nx1, ny1, nz1, fx1, fy1, fz1)
    const uint32_t imask = eb >> 24, meta = f2u(w1.z), lmask = (meta >> 16) & 0xFFu;
    const uint32_t m = bvh8_permute_hits((hits & imask) | ((hits & lmask) << 8), r.octinv);
    Node8Hits h;
    h.child_base = f2u(w1.x); h.ng_bits = (m & 0xFFu) | (imask << 8);
    h.tri_base = f2u(w1.y);   h.tg_bits = (m >> 8) | (meta << 8);            // counts16 = meta & 0xFFFF lands in bits 8..23; lmask above it (unused)
    return h;
}
#endif

// Next interior child of a node group (its highest-priority hit bit), removed from the group.
FTN_HD uint32_t node8_pop_child(uint32_t base, uint32_t& bits, uint32_t octinv) {
    const int j = bfind8(bits & 0xFFu);
    bits &= ~(1u << j);
    const uint32_t slot = (uint32_t)j ^ octinv;
    return base + (uint32_t)popc32((bits >> 8) & 0xFFu & ((1u << slot) - 1u));
}
// Next leaf child of a triangle group: first triangle and count, removed from the group.
FTN_HD void node8_pop_leaf(uint32_t base, uint32_t& bits, uint32_t octinv, uint32_t* first, uint32_t* count) {
    const int j = bfind8(bits & 0xFFu);
    bits &= ~(1u << j);
    const uint32_t slot = (uint32_t)j ^ octinv;
    const uint32_t counts = (bits >> 8) & 0xFFFFu;
    const uint32_t below = counts & ((1u << (2u * slot)) - 1u);
    *count = (counts >> (2u * slot)) & 3u;
    *first = base + (uint32_t)popc32(below & 0x5555u) + 2u * (uint32_t)popc32(below & 0xAAAAu);
}

// the triangles [first, first + count) against the ray; true if ANY and one was accepted
template <bool ANY, bool COUNT>
FTN_HD bool tris8_test(const BvhView& bvh, uint32_t first, uint32_t count, V3 ro, const RayShear& shear, float* t_max, uint32_t* best, TriHit* hit, TraceCounters* ctr) {
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

// Plain single-ray statement of the traversal (host test harness, and the reference for the warp-persistent form in
// ftn_trace_persistent.cuh): closest hit (ANY = false) or any hit.  Root = node 0 = the group {base 0, one child in the
// priority slot of slot 0}.
template <bool ANY, bool COUNT>
FTN_HD uint32_t bvh8_traverse(const BvhView& bvh, V3 ro, V3 rd, float* t_max_io, TriHit* hit_out, TraceCounters* ctr, uint32_t* max_sp_out = nullptr) {
    uint32_t best = FTN_NO_HIT_SLOT;
    if (bvh.n_nodes == 0u) return best;
    float t_max = *t_max_io;
    const Ray8 r8 = make_ray8(ro, rd);
    const RayShear shear = make_ray_shear(rd);
    uint32_t stack_b[FTN_STACK8_SIZE], stack_m[FTN_STACK8_SIZE];
    uint32_t sp = 0, max_sp = 0;
    uint32_t ng_base = 0u, ng_bits = (1u << (0u ^ r8.octinv)) | (1u << 8);   // root: slot 0 of a virtual parent with imask = 1
    uint32_t tg_base = 0u, tg_bits = 0u;
    for (;;) {
        if (tg_bits & 0xFFu) {
            uint32_t first, count;
            node8_pop_leaf(tg_base, tg_bits, r8.octinv, &first, &count);
            if (tris8_test<ANY, COUNT>(bvh, first, count, ro, shear, &t_max, &best, hit_out, ctr)) break;
            continue;
        }
        if (!(ng_bits & 0xFFu)) {
            if (sp == 0u) break;
            --sp; ng_base = stack_b[sp]; ng_bits = stack_m[sp];
        }
        const uint32_t node = node8_pop_child(ng_base, ng_bits, r8.octinv);
        if (ng_bits & 0xFFu) { stack_b[sp] = ng_base; stack_m[sp] = ng_bits; ++sp; if (sp > max_sp) max_sp = sp; }
        if (COUNT) ctr->nodes++;
        const Node8Hits h = node8_test(bvh.nodes, node, r8, t_max);
        ng_base = h.child_base; ng_bits = h.ng_bits; tg_base = h.tri_base; tg_bits = h.tg_bits;
    }
    if (max_sp_out) *max_sp_out = max_sp;
    *t_max_io = t_max;
    return best;
}

}  // namespace ftn
