// Karras-2012 LBVH topology over sorted (Morton code, index) keys and the emission of the
// BVH2x64 traversal layout (ftn_bvh.cuh).  Pure index arithmetic; FTN_HD so the host test
// harness can run it on the CPU.
#pragma once
#include "ftn_bvh.cuh"
#include "ftn_scene.h"

namespace ftn {

#if defined(__CUDA_ARCH__)
FTN_HD int clz32(uint32_t x) { return __clz((int)x); }
#else
FTN_HD int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
#endif

// Length of the common prefix of keys i and j; equal codes are disambiguated by the index
// (the sort is stable, so index order == position order).  -1 outside [0, n).
FTN_HD int lbvh_delta(const uint32_t* codes, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = codes[i], b = codes[j];
    if (a == b) return 32 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz32(a ^ b);
}

// Internal node i (0 <= i < n-1): covered key range [first,last] and split position gamma:
// left child covers [first,gamma], right child [gamma+1,last].
FTN_HD void lbvh_node_range(const uint32_t* codes, int n, int i, int* first, int* last, int* gamma) {
    int d = (lbvh_delta(codes, n, i, i + 1) - lbvh_delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = lbvh_delta(codes, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(codes, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int g = i + s * d + (d < 0 ? -1 : 0);
    *first = i < j ? i : j;
    *last = i < j ? j : i;
    *gamma = g;
}

// child encoding inside the build arrays: leaf k -> k | LBVH_LEAF_FLAG, internal k -> k
#define LBVH_LEAF_FLAG 0x80000000u

FTN_HD int lbvh_encode_leaf_ref(uint32_t first, uint32_t count) { return (int)~((first << 2) | (count - 1u)); }

struct LbvhArrays {
    uint32_t* left;      // n-1   child refs (LBVH_LEAF_FLAG | leaf) or internal index
    uint32_t* right;     // n-1
    uint32_t* first;     // n-1   covered range
    uint32_t* last;      // n-1
    uint32_t* parent;    // 2n-1  [0,n-1): internal nodes, [n-1, 2n-1): leaves
    uint32_t* arrive;    // n-1   refit arrival counters
    F4* node_lo;         // n-1
    F4* node_hi;         // n-1
};

// Triangle::world_bound (triangle.rs:151-157) and Bounds3::centroid (bounds.rs:161-163):
// centroid = min + (max - min) / 2.  lo.w / hi.w carry centroid.x / centroid.y.
FTN_HD void tri_bounds_centroid(const float* pos, const uint32_t* idx, uint32_t i, F4* lo_out, F4* hi_out, float cen[3]) {
    const uint32_t v0 = idx[3 * i], v1 = idx[3 * i + 1], v2 = idx[3 * i + 2];
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        const float p0 = pos[3 * v0 + a], p1 = pos[3 * v1 + a], p2 = pos[3 * v2 + a];
        lo[a] = fminf(fminf(fminf(FLT_MAX, p0), p1), p2);
        hi[a] = fmaxf(fmaxf(fmaxf(-FLT_MAX, p0), p1), p2);
        cen[a] = rn_add(lo[a], rn_div(rn_sub(hi[a], lo[a]), 2.0f));
    }
    lo_out->x = lo[0]; lo_out->y = lo[1]; lo_out->z = lo[2]; lo_out->w = cen[0];
    hi_out->x = hi[0]; hi_out->y = hi[1]; hi_out->z = hi[2]; hi_out->w = cen[1];
}
// morton3 (morton.rs:3-36) of Bounds3f::offset (bounds.rs:200-206) in the centroid bounds
FTN_HD uint32_t tri_morton(F4 lo, F4 hi, const float cmin[3], const float cmax[3]) {
    const float c[3] = {lo.w, hi.w, rn_add(lo.z, rn_div(rn_sub(hi.z, lo.z), 2.0f))};
    float o[3];
    for (int a = 0; a < 3; ++a) {
        o[a] = rn_sub(c[a], cmin[a]);
        if (cmax[a] > cmin[a]) o[a] = rn_div(o[a], rn_sub(cmax[a], cmin[a]));
    }
    return morton3_clamped(o[0], o[1], o[2]);
}

FTN_HD void lbvh_topology_node(const uint32_t* codes, int n, int i, const LbvhArrays& a) {
    int first, last, gamma;
    lbvh_node_range(codes, n, i, &first, &last, &gamma);
    const uint32_t l = (gamma == first) ? (LBVH_LEAF_FLAG | (uint32_t)gamma) : (uint32_t)gamma;
    const uint32_t r = (gamma + 1 == last) ? (LBVH_LEAF_FLAG | (uint32_t)(gamma + 1)) : (uint32_t)(gamma + 1);
    a.left[i] = l; a.right[i] = r; a.first[i] = (uint32_t)first; a.last[i] = (uint32_t)last;
    a.parent[(l & LBVH_LEAF_FLAG) ? (n - 1 + (int)(l & ~LBVH_LEAF_FLAG)) : (int)l] = (uint32_t)i;
    a.parent[(r & LBVH_LEAF_FLAG) ? (n - 1 + (int)(r & ~LBVH_LEAF_FLAG)) : (int)r] = (uint32_t)i;
    if (i == 0) a.parent[0] = 0xFFFFFFFFu;
}

FTN_HD void lbvh_load_child_box(const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi, uint32_t ref, F4* lo, F4* hi) {
    if (ref & LBVH_LEAF_FLAG) { *lo = leaf_lo[ref & ~LBVH_LEAF_FLAG]; *hi = leaf_hi[ref & ~LBVH_LEAF_FLAG]; }
    else {
#if defined(__CUDA_ARCH__)
        // boxes of interior nodes are produced by other SMs during the refit: read them through L2
        // (__ldcg) so a stale L1 line holding a neighbouring node cannot be served
        const float4 l = __ldcg(reinterpret_cast<const float4*>(a.node_lo + ref)), h = __ldcg(reinterpret_cast<const float4*>(a.node_hi + ref));
        lo->x = l.x; lo->y = l.y; lo->z = l.z; lo->w = l.w; hi->x = h.x; hi->y = h.y; hi->z = h.z; hi->w = h.w;
#else
        *lo = a.node_lo[ref]; *hi = a.node_hi[ref];
#endif
    }
}
FTN_HD void lbvh_join_children(const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi, uint32_t node) {
    F4 llo, lhi, rlo, rhi;
    lbvh_load_child_box(a, leaf_lo, leaf_hi, a.left[node], &llo, &lhi);
    lbvh_load_child_box(a, leaf_lo, leaf_hi, a.right[node], &rlo, &rhi);
    F4 lo, hi;   // Bounds3::join, bounds.rs:129-143 (min/max are exact)
    lo.x = fminf(llo.x, rlo.x); lo.y = fminf(llo.y, rlo.y); lo.z = fminf(llo.z, rlo.z); lo.w = 0.0f;
    hi.x = fmaxf(lhi.x, rhi.x); hi.y = fmaxf(lhi.y, rhi.y); hi.z = fmaxf(lhi.z, rhi.z); hi.w = 0.0f;
    a.node_lo[node] = lo; a.node_hi[node] = hi;
}

FTN_HD uint32_t lbvh_survives(const LbvhArrays& a, int i) { return (a.last[i] - a.first[i] + 1u > (uint32_t)FTN_LEAF_MAX) ? 1u : 0u; }

FTN_HD int lbvh_emit_child(const LbvhArrays& a, const uint32_t* survive, const uint32_t* new_index, uint32_t ref) {
    if (ref & LBVH_LEAF_FLAG) return lbvh_encode_leaf_ref(ref & ~LBVH_LEAF_FLAG, 1u);
    if (survive[ref]) return (int)new_index[ref];
    return lbvh_encode_leaf_ref(a.first[ref], a.last[ref] - a.first[ref] + 1u);
}
// ---- emission of the traversal layout (ftn_bvh.cuh) ---------------------------------------------------
// depth of interior node i in the binary tree (root = 0), by walking the parent links
FTN_HD uint32_t lbvh_depth(const LbvhArrays& a, int i) {
    uint32_t d = 0;
    uint32_t p = a.parent[i];
    while (p != 0xFFFFFFFFu) { ++d; p = a.parent[p]; }
    return d;
}
// Which surviving binary nodes become records of the emitted layout.  Width 2: all of them.
// Width 4: those at even depth; a surviving node at odd depth is absorbed into its parent's
// record, which then lists its two children directly (2-level collapse, 2..4 children per record).
FTN_HD uint32_t lbvh_is_record(const LbvhArrays& a, const uint32_t* survive, int i) {
#if FTN_BVH_WIDTH == 4
    return (survive[i] && (lbvh_depth(a, i) & 1u) == 0u) ? 1u : 0u;
#else
    return survive[i];
#endif
}

#if FTN_BVH_WIDTH == 4
struct Node4Builder {
    float lx[4], hx[4], ly[4], hy[4], lz[4], hz[4]; int child[4]; int n;
};
FTN_HD void node4_add(Node4Builder* b, F4 lo, F4 hi, int ref) {
    const int k = b->n++;
    b->lx[k] = lo.x; b->hx[k] = hi.x; b->ly[k] = lo.y; b->hy[k] = hi.y; b->lz[k] = lo.z; b->hz[k] = hi.z; b->child[k] = ref;
}
FTN_HD void node4_store(const Node4Builder& b, F4* out) {
    float lx[4], hx[4], ly[4], hy[4], lz[4], hz[4]; int ch[4];
    const float inf = FTN_INF;
    for (int k = 0; k < 4; ++k) {
        const bool used = k < b.n;
        lx[k] = used ? b.lx[k] : inf; hx[k] = used ? b.hx[k] : -inf;
        ly[k] = used ? b.ly[k] : inf; hy[k] = used ? b.hy[k] : -inf;
        lz[k] = used ? b.lz[k] : inf; hz[k] = used ? b.hz[k] : -inf;
        ch[k] = used ? b.child[k] : FTN_TRAVERSAL_DONE;
    }
    F4 q;
    q.x = lx[0]; q.y = lx[1]; q.z = lx[2]; q.w = lx[3]; out[0] = q;
    q.x = hx[0]; q.y = hx[1]; q.z = hx[2]; q.w = hx[3]; out[1] = q;
    q.x = ly[0]; q.y = ly[1]; q.z = ly[2]; q.w = ly[3]; out[2] = q;
    q.x = hy[0]; q.y = hy[1]; q.z = hy[2]; q.w = hy[3]; out[3] = q;
    q.x = lz[0]; q.y = lz[1]; q.z = lz[2]; q.w = lz[3]; out[4] = q;
    q.x = hz[0]; q.y = hz[1]; q.z = hz[2]; q.w = hz[3]; out[5] = q;
    q.x = u2f((uint32_t)ch[0]); q.y = u2f((uint32_t)ch[1]); q.z = u2f((uint32_t)ch[2]); q.w = u2f((uint32_t)ch[3]); out[6] = q;
    q.x = q.y = q.z = q.w = 0.0f; out[7] = q;
}
// child `ref` of an even-depth record: a leaf-ish subtree becomes a leaf reference; a surviving
// (odd-depth) interior node is replaced by its own two children.
FTN_HD void node4_add_subtree(Node4Builder* b, const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi,
                              const uint32_t* survive, const uint32_t* new_index, uint32_t ref, bool expand) {
    F4 lo, hi;
    if ((ref & LBVH_LEAF_FLAG) || !survive[ref]) {
        lbvh_load_child_box(a, leaf_lo, leaf_hi, ref, &lo, &hi);
        node4_add(b, lo, hi, lbvh_emit_child(a, survive, new_index, ref));   // leaf reference
    } else if (expand) {
        node4_add_subtree(b, a, leaf_lo, leaf_hi, survive, new_index, a.left[ref], false);
        node4_add_subtree(b, a, leaf_lo, leaf_hi, survive, new_index, a.right[ref], false);
    } else {
        lbvh_load_child_box(a, leaf_lo, leaf_hi, ref, &lo, &hi);
        node4_add(b, lo, hi, (int)new_index[ref]);                          // an even-depth record
    }
}
#endif

// one record of the traversal layout for binary node i (for which lbvh_is_record is 1)
FTN_HD void lbvh_emit_node(const LbvhArrays& a, const F4* leaf_lo, const F4* leaf_hi, const uint32_t* survive,
                           const uint32_t* new_index, int i, F4* nodes) {
#if FTN_BVH_WIDTH == 4
    Node4Builder b; b.n = 0;
    node4_add_subtree(&b, a, leaf_lo, leaf_hi, survive, new_index, a.left[i], true);
    node4_add_subtree(&b, a, leaf_lo, leaf_hi, survive, new_index, a.right[i], true);
    node4_store(b, nodes + (size_t)FTN_NODE_F4 * (size_t)new_index[i]);
#else
    F4 l0, h0, l1, h1;
    const uint32_t lr = a.left[i], rr = a.right[i];
    lbvh_load_child_box(a, leaf_lo, leaf_hi, lr, &l0, &h0);
    lbvh_load_child_box(a, leaf_lo, leaf_hi, rr, &l1, &h1);
    const int c0 = lbvh_emit_child(a, survive, new_index, lr), c1 = lbvh_emit_child(a, survive, new_index, rr);
    F4* out = nodes + 4 * (size_t)new_index[i];
    F4 n0, n1, nz, ci;
    n0.x = l0.x; n0.y = h0.x; n0.z = l0.y; n0.w = h0.y;
    n1.x = l1.x; n1.y = h1.x; n1.z = l1.y; n1.w = h1.y;
    nz.x = l0.z; nz.y = h0.z; nz.z = l1.z; nz.w = h1.z;
    ci.x = u2f((uint32_t)c0); ci.y = u2f((uint32_t)c1); ci.z = 0.0f; ci.w = 0.0f;
    out[0] = n0; out[1] = n1; out[2] = nz; out[3] = ci;
#endif
}
// root when the whole scene fits one leaf (n <= FTN_LEAF_MAX): one child = the leaf
FTN_HD void lbvh_emit_single(uint32_t n, const float lo[3], const float hi[3], F4* nodes) {
    const int leaf = lbvh_encode_leaf_ref(0u, n);
#if FTN_BVH_WIDTH == 4
    Node4Builder b; b.n = 0;
    F4 l, h; l.x = lo[0]; l.y = lo[1]; l.z = lo[2]; l.w = 0.0f; h.x = hi[0]; h.y = hi[1]; h.z = hi[2]; h.w = 0.0f;
    node4_add(&b, l, h, leaf);
    node4_store(b, nodes);
#else
    F4 n0, n1, nz, ci;
    const float inf = FTN_INF;
    n0.x = lo[0]; n0.y = hi[0]; n0.z = lo[1]; n0.w = hi[1];
    n1.x = inf; n1.y = -inf; n1.z = inf; n1.w = -inf;
    nz.x = lo[2]; nz.y = hi[2]; nz.z = inf; nz.w = -inf;
    ci.x = u2f((uint32_t)leaf); ci.y = u2f((uint32_t)FTN_TRAVERSAL_DONE); ci.z = 0.0f; ci.w = 0.0f;
    nodes[0] = n0; nodes[1] = n1; nodes[2] = nz; nodes[3] = ci;
#endif
}
// 64-byte pre-gathered triangle record for leaf-order slot i
FTN_HD void lbvh_gather_tri(const float* pos, const uint32_t* idx, const uint32_t* order, uint32_t i,
                            const MeshData* meshes, uint32_t n_meshes, F4* tris) {
    const uint32_t prim = order[i];
    const uint32_t v0 = idx[3 * prim], v1 = idx[3 * prim + 1], v2 = idx[3 * prim + 2];
    uint32_t lo = 0, hi = n_meshes;   // mesh of this triangle: last mesh whose first_tri <= prim
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (meshes[mid].first_tri <= prim) lo = mid; else hi = mid; }
    F4 a, b, c;
    a.x = pos[3 * v0]; a.y = pos[3 * v0 + 1]; a.z = pos[3 * v0 + 2]; a.w = u2f(prim);
    b.x = pos[3 * v1]; b.y = pos[3 * v1 + 1]; b.z = pos[3 * v1 + 2]; b.w = u2f(lo);
    c.x = pos[3 * v2]; c.y = pos[3 * v2 + 1]; c.z = pos[3 * v2 + 2]; c.w = 0.0f;
    F4 z; z.x = z.y = z.z = z.w = 0.0f;
    F4* out = tris + (size_t)FTN_TRI_F4 * (size_t)i;
    out[0] = a; out[1] = b; out[2] = c; out[3] = z;
}

}  // namespace ftn
