// PLOC -- parallel locally-ordered clustering (Meister & Bittner 2018) over the Morton-sorted
// triangles: the binary topology is built bottom-up by repeatedly merging MUTUAL nearest
// neighbours (smallest surface area of the joined box) found inside a window of +-FTN_PLOC_RADIUS
// positions of the Morton order.  It replaces the Karras radix-tree topology of ftn_lbvh.cuh, whose
// trees cost 15 % (1M-triangle sphere) to 30 % (gear ring + huge ground triangles) more issue slots
// per ray than a SAH tree (scripts/exp_sah.py); the sort, the leaf collapse and the emission of the
// traversal layout stay as they are.  Hit results do not depend on the topology.
//
// Per-element bodies only (FTN_HD): the kernels are in scene.cu, the host test harness replays
// the same bodies sequentially.  Deterministic: node ids come from prefix sums, not atomics.
//
// Build arrays (LbvhArrays): `arrive[id]` holds the triangle COUNT of internal node id during
// and after the build; ids are handed out from n-2 downwards so that the root (created last) is 0,
// as the emission expects.
#pragma once
#include "ftn_lbvh.cuh"

namespace ftn {

#ifndef FTN_PLOC_RADIUS
#define FTN_PLOC_RADIUS 16
#endif
#define PLOC_NONE 0xFFFFFFFFu
// cost model of the leaf collapse, in units of one node visit
#ifndef FTN_SAH_C_TRAV
#define FTN_SAH_C_TRAV 1.0f
#endif
#ifndef FTN_SAH_C_ISECT
#define FTN_SAH_C_ISECT 1.0f
#endif

FTN_HD void ploc_box(const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi, uint32_t ref, F4* lo, F4* hi) {
    if (ref & LBVH_LEAF_FLAG) { *lo = leaf_lo[ref & ~LBVH_LEAF_FLAG]; *hi = leaf_hi[ref & ~LBVH_LEAF_FLAG]; }
    else { *lo = a.node_lo[ref]; *hi = a.node_hi[ref]; }
}
FTN_HD uint32_t ploc_count(const LbvhArrays& a, uint32_t ref) { return (ref & LBVH_LEAF_FLAG) ? 1u : a.arrive[ref]; }

// surface area of the union of two boxes (explicitly rounded operations: host replay == device)
FTN_HD float ploc_join_area(F4 alo, F4 ahi, F4 blo, F4 bhi) {
    const float dx = rn_sub(fmaxf(ahi.x, bhi.x), fminf(alo.x, blo.x));
    const float dy = rn_sub(fmaxf(ahi.y, bhi.y), fminf(alo.y, blo.y));
    const float dz = rn_sub(fmaxf(ahi.z, bhi.z), fminf(alo.z, blo.z));
    return rn_add(rn_add(rn_mul(dx, dy), rn_mul(dy, dz)), rn_mul(dz, dx));
}

// nearest neighbour of cluster i among positions [i-R, i+R]: smallest joined area; equal areas are ordered by a
// SYMMETRIC key of the pair -- adjacent positions (2k, 2k+1) first, then distance, then the smaller position.
// Any strict total order on pairs guarantees progress (the globally smallest pair is mutual); this one makes
// runs of equal areas (regular tessellations) pair up (0,1), (2,3), ... in ONE round instead of growing a
// chain one merge per round (184 rounds / depth 65 for the 1M-triangle sphere with a plain smaller-position rule).
FTN_HD uint32_t ploc_pair_key(uint32_t i, uint32_t j) {
    const uint32_t lo = i < j ? i : j, hi = i < j ? j : i;
    const uint32_t even_adjacent = (hi == lo + 1u && (lo & 1u) == 0u) ? 0u : 1u;
    return (even_adjacent << 31) | ((hi - lo) << 24) | (lo & 0x00FFFFFFu);   // distance <= 2R < 128
}
FTN_HD uint32_t ploc_nearest(const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi, const uint32_t* cl, uint32_t c, uint32_t i) {
    F4 lo, hi;
    ploc_box(a, leaf_lo, leaf_hi, cl[i], &lo, &hi);
    const uint32_t j0 = i > (uint32_t)FTN_PLOC_RADIUS ? i - FTN_PLOC_RADIUS : 0u;
    const uint32_t j1 = (i + FTN_PLOC_RADIUS < c - 1u) ? i + FTN_PLOC_RADIUS : c - 1u;
    float best = FTN_INF; uint32_t best_key = 0xFFFFFFFFu, bj = PLOC_NONE;
    for (uint32_t j = j0; j <= j1; ++j) {
        if (j == i) continue;
        F4 l2, h2;
        ploc_box(a, leaf_lo, leaf_hi, cl[j], &l2, &h2);
        const float ar = ploc_join_area(lo, hi, l2, h2);
        const uint32_t key = ploc_pair_key(i, j);
        if (bj == PLOC_NONE || ar < best || (ar == best && key < best_key)) { best = ar; best_key = key; bj = j; }
    }
    return bj;
}

// merge[i] = 1: cluster i absorbs its mutual nearest neighbour nn[i] > i;  valid[i] = 0: cluster i is absorbed
FTN_HD void ploc_flags(const uint32_t* nn, uint32_t i, uint32_t* merge, uint32_t* valid) {
    const uint32_t j = nn[i];
    const bool mutual = j != PLOC_NONE && nn[j] == i;
    merge[i] = (mutual && i < j) ? 1u : 0u;
    valid[i] = (mutual && i > j) ? 0u : 1u;
}

// writes the cluster list of the next round; creates the internal node of a merging pair
FTN_HD void ploc_merge(const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi, const uint32_t* cl_in, uint32_t* cl_out,
                       const uint32_t* nn, const uint32_t* merge, const uint32_t* valid, const uint32_t* mscan, const uint32_t* vscan,
                       uint32_t n, uint32_t created, uint32_t i) {
    if (!valid[i]) return;
    const uint32_t pos = vscan[i];
    if (!merge[i]) { cl_out[pos] = cl_in[i]; return; }
    const uint32_t id = (n - 2u) - (created + mscan[i]);
    const uint32_t l = cl_in[i], r = cl_in[nn[i]];
    F4 llo, lhi, rlo, rhi;
    ploc_box(a, leaf_lo, leaf_hi, l, &llo, &lhi);
    ploc_box(a, leaf_lo, leaf_hi, r, &rlo, &rhi);
    F4 lo, hi;   // Bounds3::join, bounds.rs:129-143
    lo.x = fminf(llo.x, rlo.x); lo.y = fminf(llo.y, rlo.y); lo.z = fminf(llo.z, rlo.z);
    hi.x = fmaxf(lhi.x, rhi.x); hi.y = fmaxf(lhi.y, rhi.y); hi.z = fmaxf(lhi.z, rhi.z);
    const uint32_t count = ploc_count(a, l) + ploc_count(a, r);
    // surface-area-heuristic cost of the cheapest form of this subtree (lo.w) and whether that form is
    // ONE leaf (hi.w): cost = min(C_isect * count * A, C_trav * A + cost(left) + cost(right))
    const float area = ploc_join_area(llo, lhi, rlo, rhi);
    const float cl_ = (l & LBVH_LEAF_FLAG) ? rn_mul(FTN_SAH_C_ISECT, ploc_join_area(llo, lhi, llo, lhi)) : llo.w;
    const float cr_ = (r & LBVH_LEAF_FLAG) ? rn_mul(FTN_SAH_C_ISECT, ploc_join_area(rlo, rhi, rlo, rhi)) : rlo.w;
    const float c_split = rn_add(rn_add(rn_mul(FTN_SAH_C_TRAV, area), cl_), cr_);
    const float c_leaf = rn_mul(rn_mul(FTN_SAH_C_ISECT, (float)count), area);
    const bool as_leaf = count <= (uint32_t)FTN_LEAF_MAX && c_leaf <= c_split;
    lo.w = as_leaf ? c_leaf : c_split;
    hi.w = as_leaf ? 1.0f : 0.0f;
    a.node_lo[id] = lo; a.node_hi[id] = hi;
    a.left[id] = l; a.right[id] = r;
    a.arrive[id] = count;
    a.parent[(l & LBVH_LEAF_FLAG) ? (n - 1u + (l & ~LBVH_LEAF_FLAG)) : l] = id;
    a.parent[(r & LBVH_LEAF_FLAG) ? (n - 1u + (r & ~LBVH_LEAF_FLAG)) : r] = id;
    cl_out[pos] = id;
}

// Position of the first triangle of subtree `ref` in the depth-first order of the finished tree
// (= number of triangles in left siblings along the path to the root), and the depth of `ref`.
FTN_HD uint32_t ploc_dfs_position(const LbvhArrays& a, uint32_t n, uint32_t ref, uint32_t* depth_out) {
    uint32_t child = ref;
    uint32_t node = a.parent[(ref & LBVH_LEAF_FLAG) ? (n - 1u + (ref & ~LBVH_LEAF_FLAG)) : ref];
    uint32_t pos = 0, depth = 0;
    while (node != PLOC_NONE) {
        if (a.right[node] == child) pos += ploc_count(a, a.left[node]);
        child = node;
        node = a.parent[node];
        ++depth;
    }
    *depth_out = depth;
    return pos;
}

// PLOC trees: an internal node stays an interior node unless it, or an ancestor of <= FTN_LEAF_MAX
// triangles, prefers to be a single leaf (hi.w set by ploc_merge).
FTN_HD uint32_t ploc_survives(const LbvhArrays& a, int i) {
    uint32_t node = (uint32_t)i;
    while (node != PLOC_NONE && a.arrive[node] <= (uint32_t)FTN_LEAF_MAX) {
        if (a.node_hi[node].w != 0.0f) return 0u;
        node = a.parent[node];
    }
    return 1u;
}

}  // namespace ftn
