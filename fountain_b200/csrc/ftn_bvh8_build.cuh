// Build of the BVH8q layout (ftn_bvh8.cuh): greedy collapse of the binary tree (Karras radix tree or PLOC,
// LbvhArrays with first/last leaf ranges) into 8-wide nodes, octant slot assignment, conservative 8-bit
// quantisation of the child boxes.  Per-node body only (FTN_HD): the level-synchronous kernel is in scene.cu,
// the host test harness replays the same body breadth-first.
#pragma once
#include "ftn_bvh8.cuh"
#include "ftn_lbvh.cuh"

namespace ftn {

// ---- build: collapse of the binary tree (Karras or PLOC, LbvhArrays with first/last ranges in leaf order) -------------
FTN_HD void bvh8_ref_box(const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi, uint32_t ref, F4* lo, F4* hi) {
    if (ref & LBVH_LEAF_FLAG) { *lo = leaf_lo[ref & ~LBVH_LEAF_FLAG]; *hi = leaf_hi[ref & ~LBVH_LEAF_FLAG]; }
    else { *lo = a.node_lo[ref]; *hi = a.node_hi[ref]; }
}
FTN_HD uint32_t bvh8_ref_count(const LbvhArrays& a, uint32_t ref) { return (ref & LBVH_LEAF_FLAG) ? 1u : a.last[ref] - a.first[ref] + 1u; }
FTN_HD uint32_t bvh8_ref_first(const LbvhArrays& a, uint32_t ref) { return (ref & LBVH_LEAF_FLAG) ? (ref & ~LBVH_LEAF_FLAG) : a.first[ref]; }
FTN_HD float bvh8_half_area(F4 lo, F4 hi) {
    const float dx = rn_sub(hi.x, lo.x), dy = rn_sub(hi.y, lo.y), dz = rn_sub(hi.z, lo.z);
    return rn_add(rn_add(rn_mul(dx, dy), rn_mul(dy, dz)), rn_mul(dz, dx));
}

// smallest biased exponent E (scale = 2^(E-127)) with 255 * scale >= extent; clamped so that E + 15 stays a finite exponent
FTN_HD uint32_t bvh8_exponent(double extent) {
    int e = -126;
    if (extent > 0.0) {
        int k;
        const double m = frexp(extent / 255.0, &k);      // extent / 255 = m 2^k, m in [0.5, 1)
        e = (m == 0.5) ? k - 1 : k;
        while (ldexp(255.0, e) < extent) ++e;            // the division above rounds
    }
    if (e < -110) e = -110;
    if (e > 112) e = 112;
    return (uint32_t)(e + 127);
}
FTN_HD uint32_t bvh8_quant_lo(double v, double p, uint32_t E) {
    double q = floor(ldexp(v - p, 127 - (int)E));
    while (q > 0.0 && p + ldexp(q, (int)E - 127) > v) q -= 1.0;
    return q < 0.0 ? 0u : (q > 255.0 ? 255u : (uint32_t)q);
}
FTN_HD uint32_t bvh8_quant_hi(double v, double p, uint32_t E) {
    double q = ceil(ldexp(v - p, 127 - (int)E));
    while (q < 255.0 && p + ldexp(q, (int)E - 127) < v) q += 1.0;
    return q < 0.0 ? 0u : (q > 255.0 ? 255u : (uint32_t)q);
}

// One wide node.  `bref` = the binary subtree it covers (root call: the whole tree, or one leaf range for n <= 3 triangles:
// bref = LBVH_LEAF_FLAG with `single_count` triangles).  The caller allocates: alloc(n_inner, n_tris, &child_base, &tri_base).
// Writes the record, the binary refs of its interior children (brefs[child_base + k]) and the triangle order of its leaf
// children (order_out[tri_base + ..] = order_in[leaf position]).
template <class Alloc>
FTN_HD void bvh8_collapse_node(const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi, uint32_t bref, uint32_t single_count,
                               F4 box_lo, F4 box_hi, uint32_t w, Alloc& alloc, F4* nodes, uint32_t* brefs,
                               const uint32_t* order_in, uint32_t* order_out) {
    uint32_t ref[8]; F4 lo[8], hi[8]; float area[8]; uint32_t cnt[8];
    int n = 0;
    if (single_count != 0u) {                            // tiny scene: one leaf child holding everything
        ref[0] = LBVH_LEAF_FLAG; lo[0] = box_lo; hi[0] = box_hi; area[0] = 0.0f; cnt[0] = single_count; n = 1;
    } else {
        ref[0] = a.left[bref]; ref[1] = a.right[bref]; n = 2;
        for (int i = 0; i < 2; ++i) { bvh8_ref_box(a, leaf_lo, leaf_hi, ref[i], &lo[i], &hi[i]); area[i] = bvh8_half_area(lo[i], hi[i]); cnt[i] = bvh8_ref_count(a, ref[i]); }
        // 1) open the largest subtree that cannot be a leaf until eight children are listed;
        // 2) with slots to spare, split multi-triangle leaves (tighter boxes at no extra record)
        for (int phase = 0; phase < 2; ++phase) {
            const uint32_t keep = phase == 0 ? (uint32_t)FTN_LEAF8_MAX : 1u;
            while (n < 8) {
                int best = -1;
                for (int i = 0; i < n; ++i) if (cnt[i] > keep && (best < 0 || area[i] > area[best])) best = i;
                if (best < 0) break;
                const uint32_t r = ref[best], l2 = a.left[r], r2 = a.right[r];
                ref[best] = l2; bvh8_ref_box(a, leaf_lo, leaf_hi, l2, &lo[best], &hi[best]); area[best] = bvh8_half_area(lo[best], hi[best]); cnt[best] = bvh8_ref_count(a, l2);
                ref[n] = r2; bvh8_ref_box(a, leaf_lo, leaf_hi, r2, &lo[n], &hi[n]); area[n] = bvh8_half_area(lo[n], hi[n]); cnt[n] = bvh8_ref_count(a, r2);
                ++n;
            }
        }
    }
    // octant slots: greedy assignment maximising (centroid - node centre) . (+-1, +-1, +-1)(slot)
    float vx[8], vy[8], vz[8];
    const float ccx = 0.5f * (box_lo.x + box_hi.x), ccy = 0.5f * (box_lo.y + box_hi.y), ccz = 0.5f * (box_lo.z + box_hi.z);
    for (int i = 0; i < n; ++i) { vx[i] = 0.5f * (lo[i].x + hi[i].x) - ccx; vy[i] = 0.5f * (lo[i].y + hi[i].y) - ccy; vz[i] = 0.5f * (lo[i].z + hi[i].z) - ccz; }
    int slot_of[8]; uint32_t used_slots = 0u, assigned = 0u;
    for (int it = 0; it < n; ++it) {
        float bc = -FTN_INF; int bi = -1, bs = -1;
        for (int i = 0; i < n; ++i) {
            if (assigned & (1u << i)) continue;
            for (int s = 0; s < 8; ++s) {
                if (used_slots & (1u << s)) continue;
                const float c = ((s & 1) ? vx[i] : -vx[i]) + ((s & 2) ? vy[i] : -vy[i]) + ((s & 4) ? vz[i] : -vz[i]);
                if (bi < 0 || c > bc) { bc = c; bi = i; bs = s; }
            }
        }
        slot_of[bi] = bs; assigned |= 1u << bi; used_slots |= 1u << bs;
    }
    int child_in_slot[8];
    for (int s = 0; s < 8; ++s) child_in_slot[s] = -1;
    for (int i = 0; i < n; ++i) child_in_slot[slot_of[i]] = i;
    uint32_t imask = 0u, lmask = 0u, counts16 = 0u, n_inner = 0u, n_tris = 0u;
    for (int s = 0; s < 8; ++s) {
        const int i = child_in_slot[s];
        if (i < 0) continue;
        if (cnt[i] > (uint32_t)FTN_LEAF8_MAX) { imask |= 1u << s; ++n_inner; }
        else { lmask |= 1u << s; counts16 |= cnt[i] << (2 * s); n_tris += cnt[i]; }
    }
    uint32_t child_base = 0u, tri_base = 0u;
    alloc(n_inner, n_tris, &child_base, &tri_base);
    // quantisation frame
    const double px = box_lo.x, py = box_lo.y, pz = box_lo.z;
    const uint32_t Ex = bvh8_exponent((double)box_hi.x - px), Ey = bvh8_exponent((double)box_hi.y - py), Ez = bvh8_exponent((double)box_hi.z - pz);
#if FTN_BVH8_PLANES16
    uint32_t q[6][4] = {{0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}, {0u, 0u, 0u, 0u}};   // lo.x lo.y lo.z hi.x hi.y hi.z, four words of two bf16
#else
    uint32_t q[6][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}};   // lo.x lo.y lo.z hi.x hi.y hi.z, two words of four bytes
#endif
    uint32_t ci = 0u, ti = 0u;
    for (int s = 0; s < 8; ++s) {
        const int i = child_in_slot[s];
#if FTN_BVH8_PLANES16
        const int wsel = s >> 1, sh = 16 * (s & 1);
        // the integer as bf16 (exact up to 255); a lo plane in a HIGH half reads up to one unit too large, so it is stored one lower
        #define FTN_Q16(v) (f2u((float)(v)) >> 16)
        if (i < 0) { q[0][wsel] |= FTN_Q16(255u) << sh; q[1][wsel] |= FTN_Q16(255u) << sh; q[2][wsel] |= FTN_Q16(255u) << sh; continue; }
        uint32_t qlx = bvh8_quant_lo(lo[i].x, px, Ex), qly = bvh8_quant_lo(lo[i].y, py, Ey), qlz = bvh8_quant_lo(lo[i].z, pz, Ez);
        if (s & 1) { qlx = qlx ? qlx - 1u : 0u; qly = qly ? qly - 1u : 0u; qlz = qlz ? qlz - 1u : 0u; }
        q[0][wsel] |= FTN_Q16(qlx) << sh; q[1][wsel] |= FTN_Q16(qly) << sh; q[2][wsel] |= FTN_Q16(qlz) << sh;
        q[3][wsel] |= FTN_Q16(bvh8_quant_hi(hi[i].x, px, Ex)) << sh; q[4][wsel] |= FTN_Q16(bvh8_quant_hi(hi[i].y, py, Ey)) << sh; q[5][wsel] |= FTN_Q16(bvh8_quant_hi(hi[i].z, pz, Ez)) << sh;
        #undef FTN_Q16
#else
        const int wsel = s >> 2, sh = 8 * (s & 3);
        if (i < 0) {   // empty slot: inverted box (also masked out by imask | lmask)
            q[0][wsel] |= 255u << sh; q[1][wsel] |= 255u << sh; q[2][wsel] |= 255u << sh;
            continue;
        }
        q[0][wsel] |= bvh8_quant_lo(lo[i].x, px, Ex) << sh; q[1][wsel] |= bvh8_quant_lo(lo[i].y, py, Ey) << sh; q[2][wsel] |= bvh8_quant_lo(lo[i].z, pz, Ez) << sh;
        q[3][wsel] |= bvh8_quant_hi(hi[i].x, px, Ex) << sh; q[4][wsel] |= bvh8_quant_hi(hi[i].y, py, Ey) << sh; q[5][wsel] |= bvh8_quant_hi(hi[i].z, pz, Ez) << sh;
#endif
        if (imask & (1u << s)) { brefs[child_base + ci] = ref[i]; ++ci; }
        else {
            const uint32_t f = single_count != 0u ? 0u : bvh8_ref_first(a, ref[i]);
            for (uint32_t j = 0; j < cnt[i]; ++j) order_out[tri_base + ti + j] = order_in[f + j];
            ti += cnt[i];
        }
    }
    F4* out = nodes + (size_t)FTN_NODE8_F4 * (size_t)w;
    F4 v;
    v.x = box_lo.x; v.y = box_lo.y; v.z = box_lo.z; v.w = u2f(Ex | (Ey << 8) | (Ez << 16) | (imask << 24)); out[0] = v;
    v.x = u2f(child_base); v.y = u2f(tri_base); v.z = u2f(counts16 | (lmask << 16)); v.w = 0.0f; out[1] = v;
#if FTN_BVH8_PLANES16
    for (int pl = 0; pl < 6; ++pl) { v.x = u2f(q[pl][0]); v.y = u2f(q[pl][1]); v.z = u2f(q[pl][2]); v.w = u2f(q[pl][3]); out[2 + pl] = v; }
#else
    v.x = u2f(q[0][0]); v.y = u2f(q[0][1]); v.z = u2f(q[1][0]); v.w = u2f(q[1][1]); out[2] = v;
    v.x = u2f(q[2][0]); v.y = u2f(q[2][1]); v.z = u2f(q[3][0]); v.w = u2f(q[3][1]); out[3] = v;
    v.x = u2f(q[4][0]); v.y = u2f(q[4][1]); v.z = u2f(q[5][0]); v.w = u2f(q[5][1]); out[4] = v;
    v.x = v.y = v.z = v.w = 0.0f; out[5] = v;
#endif
}

// upper bound of the number of wide nodes for n triangles (every all-leaf node holds >= 4 triangles, every other node
// eight children: at most 2n/7 + 1 records)
FTN_HD size_t bvh8_max_nodes(size_t n) { return n / 3 + 8; }

}  // namespace ftn
