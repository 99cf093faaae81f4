// Scene upload into flat SoA device buffers and the Morton-code LBVH build
// (replaces Scene::new scene/mod.rs:32-49 and BVH::build bvh.rs:27-158).
//
// Build pipeline, all on the device:
//   k_tri_bounds      per-triangle AABB + centroid (Triangle::world_bound triangle.rs:151-157,
//                     Bounds3::centroid bounds.rs:161-163) and the scene / centroid bounds
//   k_morton          morton3 (morton.rs:3-36) of Bounds3f::offset (bounds.rs:200-206)
//   radix_sort_pairs  stable sort of (code, primitive) -- sort_scan.cu
//   k_lbvh_topology   Karras hierarchy over the sorted keys
//   k_lbvh_refit      bottom-up boxes with per-node arrival counters
//   k_lbvh_emit       leaf collapse (<= FTN_LEAF_MAX triangles) + BVH2x64 node records
//   k_bvh8_level      (default layout) level-synchronous collapse of the binary tree into BVH8q records (ftn_bvh8.cuh)
//   k_gather_tris     64-byte pre-gathered triangle records in leaf order
#include "ftn_scene.h"
#include "ftn_lbvh.cuh"
#include "ftn_ploc.cuh"
#include "ftn_bvh8_build.cuh"
#include <cstdio>
#include <cstdlib>
#define FTN_REFILL_THRESHOLD_DEFAULT 16
#define FTN_VOTE_BIAS_DEFAULT 14
#define FTN_VOTE_BIAS_WIDE_DEFAULT 28
#define FTN_VOTE_MIN_TRIS 65536u
#define FTN_PLOC_MIN_TRIS 65536u
#define FTN_WIDE_MIN_TRIS 65536u
#include <chrono>
#include <cmath>
#include <cstring>

namespace ftn {

// ---- ordered-float atomics: exact (min/max do not round), order independent --------------------
__device__ __forceinline__ uint32_t float_flip(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float float_unflip(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }

struct BuildBounds { uint32_t scene_lo[3], scene_hi[3], cen_lo[3], cen_hi[3]; };

__global__ void k_init_bounds(BuildBounds* b) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int i = 0; i < 3; ++i) {
            b->scene_lo[i] = float_flip(FLT_MAX); b->scene_hi[i] = float_flip(-FLT_MAX);   // Bounds3::empty, bounds.rs:125-127
            b->cen_lo[i] = float_flip(FLT_MAX); b->cen_hi[i] = float_flip(-FLT_MAX);
        }
    }
}

__global__ void __launch_bounds__(256)
k_tri_bounds(const float* __restrict__ pos, const uint32_t* __restrict__ idx, uint32_t n_tris,
             F4* __restrict__ tri_lo, F4* __restrict__ tri_hi, BuildBounds* __restrict__ gb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n_tris) {
        F4 l, h; float c[3];
        tri_bounds_centroid(pos, idx, i, &l, &h, c);
        lo[0] = l.x; lo[1] = l.y; lo[2] = l.z; hi[0] = h.x; hi[1] = h.y; hi[2] = h.z;
        for (int a = 0; a < 3; ++a) { clo[a] = c[a]; chi[a] = c[a]; }
        tri_lo[i] = l; tri_hi[i] = h;   // centroid.x / .y ride in the w lanes; centroid.z is recomputed
    }
    // warp reduce, then one atomic per warp per component
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
            clo[a] = fminf(clo[a], __shfl_xor_sync(0xffffffffu, clo[a], o));
            chi[a] = fmaxf(chi[a], __shfl_xor_sync(0xffffffffu, chi[a], o));
        }
    }
    // ... then a block reduce in shared memory and ONE atomic per block per component (12 contended
    // addresses: per-warp atomics made this kernel 240 us for 1M triangles)
    __shared__ float red[8][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) for (int a = 0; a < 3; ++a) { red[warp][a] = lo[a]; red[warp][3 + a] = hi[a]; red[warp][6 + a] = clo[a]; red[warp][9 + a] = chi[a]; }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int k = threadIdx.x;
        const bool is_min = (k < 3) || (k >= 6 && k < 9);
        float v = red[0][k];
        for (int w = 1; w < 8; ++w) v = is_min ? fminf(v, red[w][k]) : fmaxf(v, red[w][k]);
        uint32_t* dst = (k < 3) ? &gb->scene_lo[k] : (k < 6) ? &gb->scene_hi[k - 3] : (k < 9) ? &gb->cen_lo[k - 6] : &gb->cen_hi[k - 9];
        if (is_min) atomicMin(dst, float_flip(v)); else atomicMax(dst, float_flip(v));
    }
}

__global__ void __launch_bounds__(256)
k_morton(const F4* __restrict__ tri_lo, const F4* __restrict__ tri_hi, uint32_t n_tris, const BuildBounds* __restrict__ gb,
         uint32_t* __restrict__ codes_in_order, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tris) return;
    float cmin[3], cmax[3];
    for (int a = 0; a < 3; ++a) { cmin[a] = float_unflip(gb->cen_lo[a]); cmax[a] = float_unflip(gb->cen_hi[a]); }
    const uint32_t code = tri_morton(tri_lo[i], tri_hi[i], cmin, cmax);
    codes_in_order[i] = code;
    keys[i] = code;
    vals[i] = i;
}

// ---- Karras topology, refit, emission (bodies in ftn_lbvh.cuh) ----------------------------------------------
__global__ void __launch_bounds__(256)
k_lbvh_topology(const uint32_t* __restrict__ codes, int n, LbvhArrays a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) lbvh_topology_node(codes, n, i, a);
}

// one thread per leaf climbs; the second arrival at a node owns it (both children are complete)
__global__ void __launch_bounds__(256)
k_lbvh_refit(int n, LbvhArrays a, const F4* __restrict__ leaf_lo, const F4* __restrict__ leaf_hi) {
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    uint32_t node = a.parent[n - 1 + leaf];
    while (node != 0xFFFFFFFFu) {
        __threadfence();
        if (atomicAdd(&a.arrive[node], 1u) == 0u) return;
        __threadfence();
        lbvh_join_children(a, leaf_lo, leaf_hi, node);
        node = a.parent[node];
    }
}

// ---- PLOC topology (bodies in ftn_ploc.cuh) -------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ploc_init(uint32_t n, uint32_t* __restrict__ cl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cl[i] = LBVH_LEAF_FLAG | i;
}
// The multi-launch rounds keep the cluster count and the number of nodes created so far ON THE DEVICE
// (state = {c, created}, double-buffered per round) and are launched for an upper bound c_bound of the
// count (it only shrinks): the host reads the state back once per GROUP of rounds, not once per round.
__global__ void __launch_bounds__(256)
k_ploc_nearest(LbvhArrays a, const F4* __restrict__ leaf_lo, const F4* __restrict__ leaf_hi, const uint32_t* __restrict__ cl,
               const uint32_t* __restrict__ state, uint32_t* __restrict__ nn) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = state[0];
    if (i < c) nn[i] = ploc_nearest(a, leaf_lo, leaf_hi, cl, c, i);
}
__global__ void __launch_bounds__(256)
k_ploc_flags(const uint32_t* __restrict__ nn, const uint32_t* __restrict__ state, uint32_t c_bound, uint32_t* __restrict__ merge,
             uint32_t* __restrict__ valid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c_bound) return;
    if (i < state[0]) ploc_flags(nn, i, merge, valid);
    else { merge[i] = 0u; valid[i] = 0u; }     // beyond the live clusters: nothing to scan
}
__global__ void __launch_bounds__(256)
k_ploc_merge(LbvhArrays a, const F4* __restrict__ leaf_lo, const F4* __restrict__ leaf_hi, const uint32_t* __restrict__ cl_in,
             uint32_t* __restrict__ cl_out, const uint32_t* __restrict__ nn, const uint32_t* __restrict__ merge, const uint32_t* __restrict__ valid,
             const uint32_t* __restrict__ mscan, const uint32_t* __restrict__ vscan, uint32_t n, const uint32_t* __restrict__ state,
             uint32_t* __restrict__ state_next) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = state[0], created = state[1];
    if (i >= c) return;
    ploc_merge(a, leaf_lo, leaf_hi, cl_in, cl_out, nn, merge, valid, mscan, vscan, n, created, i);
    if (i == c - 1u) { const uint32_t merged = mscan[i] + merge[i]; state_next[0] = c - merged; state_next[1] = created + merged; }
}
// The tail of the clustering in ONE block: once at most PLOC_FINISH_MAX clusters are left, the remaining
// rounds (the majority: ~200 rounds for 1M triangles, ~14 of them above this size) run inside a single
// kernel with the cluster lists in shared memory -- no launches and no host read-back per round.
// Same per-element bodies, same prefix-sum node numbering: the tree is identical to the multi-launch form.
#define PLOC_FINISH_MAX 1024
__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = warp_sums[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
        warp_sums[lane] = wi - w;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    const uint32_t r = warp_sums[warp] + incl - v;
    __syncthreads();   // warp_sums / total are reused by the next scan
    return r;
}
__global__ void __launch_bounds__(PLOC_FINISH_MAX)
k_ploc_finish(LbvhArrays a, const F4* __restrict__ leaf_lo, const F4* __restrict__ leaf_hi, const uint32_t* __restrict__ cl_global,
              uint32_t c, uint32_t n, uint32_t created, uint32_t* __restrict__ status) {
    __shared__ uint32_t cl[2][PLOC_FINISH_MAX], nn[PLOC_FINISH_MAX], mg[PLOC_FINISH_MAX], va[PLOC_FINISH_MAX], ms[PLOC_FINISH_MAX], vs[PLOC_FINISH_MAX];
    __shared__ uint32_t warp_sums[32], tot_m, tot_v;
    const uint32_t i = threadIdx.x;
    if (i < c) cl[0][i] = cl_global[i];
    __syncthreads();
    int cur = 0;
    while (c > 1u) {
        if (i < c) nn[i] = ploc_nearest(a, leaf_lo, leaf_hi, cl[cur], c, i);
        __syncthreads();
        if (i < c) ploc_flags(nn, i, mg, va);
        __syncthreads();
        const uint32_t m_ex = block_exclusive_scan_1024(i < c ? mg[i] : 0u, warp_sums, &tot_m);
        const uint32_t merged = tot_m;
        const uint32_t v_ex = block_exclusive_scan_1024(i < c ? va[i] : 0u, warp_sums, &tot_v);
        if (i < c) { ms[i] = m_ex; vs[i] = v_ex; }
        __syncthreads();
        if (merged == 0u) { if (i == 0) status[0] = 1u; return; }   // cannot happen (see ploc_nearest); reported, not spun on
        if (i < c) ploc_merge(a, leaf_lo, leaf_hi, cl[cur], cl[cur ^ 1], nn, mg, va, ms, vs, n, created, i);
        __threadfence_block();
        __syncthreads();
        created += merged; c -= merged; cur ^= 1;
    }
    if (i == 0) { status[1] = created; status[2] = cl[cur][0]; }
}

// depth-first position of every leaf (and the deepest leaf), range of every internal node
__global__ void __launch_bounds__(256)
k_ploc_leaf_positions(LbvhArrays a, uint32_t n, uint32_t* __restrict__ newpos, uint32_t* __restrict__ max_depth) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    uint32_t d;
    newpos[l] = ploc_dfs_position(a, n, LBVH_LEAF_FLAG | l, &d);
    atomicMax(max_depth, d);
}
__global__ void __launch_bounds__(256)
k_ploc_node_ranges(LbvhArrays a, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1u) return;
    uint32_t d;
    const uint32_t f = ploc_dfs_position(a, n, i, &d);
    a.first[i] = f; a.last[i] = f + a.arrive[i] - 1u;
}
__global__ void __launch_bounds__(256)
k_ploc_permute_leaves(uint32_t n, const uint32_t* __restrict__ newpos, const F4* __restrict__ lo_in, const F4* __restrict__ hi_in,
                      const uint32_t* __restrict__ order_in, F4* __restrict__ lo_out, F4* __restrict__ hi_out, uint32_t* __restrict__ order_out) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const uint32_t p = newpos[l];
    lo_out[p] = lo_in[l]; hi_out[p] = hi_in[l]; order_out[p] = order_in[l];
}
__global__ void __launch_bounds__(256)
k_ploc_rewrite_refs(LbvhArrays a, uint32_t n, const uint32_t* __restrict__ newpos) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1u) return;
    const uint32_t l = a.left[i], r = a.right[i];
    if (l & LBVH_LEAF_FLAG) a.left[i] = LBVH_LEAF_FLAG | newpos[l & ~LBVH_LEAF_FLAG];
    if (r & LBVH_LEAF_FLAG) a.right[i] = LBVH_LEAF_FLAG | newpos[r & ~LBVH_LEAF_FLAG];
}
__global__ void __launch_bounds__(256)
k_ploc_survive(int n, LbvhArrays a, uint32_t* __restrict__ survive) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) survive[i] = ploc_survives(a, i);
}

__global__ void __launch_bounds__(256)
k_gather_leaf_boxes(const F4* __restrict__ tri_lo, const F4* __restrict__ tri_hi, const uint32_t* __restrict__ order, uint32_t n,
                    F4* __restrict__ leaf_lo, F4* __restrict__ leaf_hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    leaf_lo[i] = tri_lo[order[i]]; leaf_hi[i] = tri_hi[order[i]];
}

__global__ void __launch_bounds__(256)
k_lbvh_survive(int n, LbvhArrays a, uint32_t* __restrict__ survive) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) survive[i] = lbvh_survives(a, i);
}

__global__ void __launch_bounds__(256)
k_lbvh_mark_records(int n, LbvhArrays a, const uint32_t* __restrict__ survive, uint32_t* __restrict__ is_record) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1) is_record[i] = lbvh_is_record(a, survive, i);
}

__global__ void __launch_bounds__(256)
k_lbvh_emit(int n, LbvhArrays a, const F4* __restrict__ leaf_lo, const F4* __restrict__ leaf_hi,
            const uint32_t* __restrict__ survive, const uint32_t* __restrict__ is_record, const uint32_t* __restrict__ new_index,
            F4* __restrict__ nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n - 1 && is_record[i]) lbvh_emit_node(a, leaf_lo, leaf_hi, survive, new_index, i, nodes);
}

__global__ void k_lbvh_emit_single(uint32_t n, const BuildBounds* gb, F4* nodes) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = float_unflip(gb->scene_lo[a]); hi[a] = float_unflip(gb->scene_hi[a]); }
    lbvh_emit_single(n, lo, hi, nodes);
}

__global__ void __launch_bounds__(256)
k_gather_tris(const float* __restrict__ pos, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ order, uint32_t n,
              const MeshData* __restrict__ meshes, uint32_t n_meshes, F4* __restrict__ tris) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lbvh_gather_tri(pos, idx, order, i, meshes, n_meshes, tris);
}


// ---- BVH8q collapse (body in ftn_bvh8_build.cuh) ------------------------------------------------------------------
// One launch per level of the wide tree.  ctr[l] = number of wide nodes of level l (ctr[0] = 1, the root); the nodes of
// level l are records [sum ctr[0..l), +ctr[l]) and brefs[] holds the binary subtree each of them covers.  A node reserves
// the records of its interior children with one atomicAdd on ctr[l + 1] and its triangles with one on ctr[BVH8_CTR_TRIS]:
// record and triangle ADDRESSES depend on the order of those atomics, the tree and every traversal result do not.
#define BVH8_MAX_LEVELS 256
#define BVH8_CTR_TRIS (BVH8_MAX_LEVELS + 1)
#define BVH8_CTR_ERROR (BVH8_MAX_LEVELS + 2)
#define BVH8_CTR_COUNT (BVH8_MAX_LEVELS + 3)
struct Bvh8DeviceAlloc {
    uint32_t next_begin, max_nodes; uint32_t* ctr; uint32_t level;
    __device__ void operator()(uint32_t n_inner, uint32_t n_tris, uint32_t* child_base, uint32_t* tri_base) const {
        uint32_t cb = next_begin + (n_inner ? atomicAdd(&ctr[level + 1], n_inner) : 0u);
        if (cb + n_inner > max_nodes) { atomicExch(&ctr[BVH8_CTR_ERROR], 1u); cb = 0u; }   // cannot happen (bvh8_max_nodes); reported, not trusted
        *child_base = cb;
        *tri_base = n_tris ? atomicAdd(&ctr[BVH8_CTR_TRIS], n_tris) : 0u;
    }
};
__global__ void __launch_bounds__(128)
k_bvh8_level(LbvhArrays a, const F4* __restrict__ leaf_lo, const F4* __restrict__ leaf_hi, uint32_t level, uint32_t* __restrict__ ctr,
             F4* __restrict__ nodes, uint32_t* __restrict__ brefs, const uint32_t* __restrict__ order_in, uint32_t* __restrict__ order_out,
             uint32_t max_nodes, uint32_t single_count, const BuildBounds* __restrict__ gb) {
    __shared__ uint32_t s_begin;
    if (threadIdx.x == 0) { uint32_t b = 0; for (uint32_t l = 0; l < level; ++l) b += ctr[l]; s_begin = b; }
    __syncthreads();
    const uint32_t begin = s_begin, cnt = ctr[level];
    if (level + 1u >= (uint32_t)BVH8_MAX_LEVELS) { if (cnt != 0u && threadIdx.x == 0 && blockIdx.x == 0) atomicExch(&ctr[BVH8_CTR_ERROR], 2u); return; }
    Bvh8DeviceAlloc alloc; alloc.next_begin = begin + cnt; alloc.max_nodes = max_nodes; alloc.ctr = ctr; alloc.level = level;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const uint32_t w = begin + i, bref = brefs[w];
        F4 lo, hi;
        if (single_count != 0u) {
            lo.x = float_unflip(gb->scene_lo[0]); lo.y = float_unflip(gb->scene_lo[1]); lo.z = float_unflip(gb->scene_lo[2]); lo.w = 0.0f;
            hi.x = float_unflip(gb->scene_hi[0]); hi.y = float_unflip(gb->scene_hi[1]); hi.z = float_unflip(gb->scene_hi[2]); hi.w = 0.0f;
        } else bvh8_ref_box(a, leaf_lo, leaf_hi, bref, &lo, &hi);
        bvh8_collapse_node(a, leaf_lo, leaf_hi, bref, single_count, lo, hi, w, alloc, nodes, brefs, order_in, order_out);
    }
}

// ---- env-map tables on the device -------------------------------------------------------------------------
// MIPMap level-0 lookups (mipmap.rs:245-312); shared with the shading kernels via ftn_shade.cuh.
}  // namespace ftn
#include "ftn_shade.cuh"
namespace ftn {

__global__ void __launch_bounds__(256)
k_env_func(EnvLightData env, float* __restrict__ func) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < env.nu * env.nv) func[k] = env_func_value(env, k);
}
__global__ void __launch_bounds__(128)
k_env_row_cdf(const float* __restrict__ func, int nu, int nv, float* __restrict__ cdf, float* __restrict__ integral) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) dist_row_build(func + (size_t)v * nu, nu, cdf + (size_t)v * (nu + 1), integral + v);
}

// The conditional rows of an environment map (2048 rows of 1024 for C4): the running sum of a row is sequential by
// definition (dist_row_build), so one LANE owns one row, and the warp moves 32x32 tiles through shared memory so that
// the global loads and stores are coalesced (one thread per row with strided accesses took 0.9 ms for 8 MB).  Same
// operations in the same order as dist_row_build: the tables are bit-identical.
__global__ void __launch_bounds__(128)
k_env_rows_running_sum(const float* __restrict__ func, int nu, int nv, float* __restrict__ cdf, float* __restrict__ integral) {
    __shared__ float tile[4][32][33];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * 4 + warp) * 32;
    if (row0 >= nv) return;   // warp-uniform
    float (*t)[33] = tile[warp];
    const float nf = (float)nu;
    float run = 0.0f;
    for (int k0 = 0; k0 < nu; k0 += 32) {
        const int col = k0 + lane;
        for (int r = 0; r < 32; ++r) t[r][lane] = (row0 + r < nv && col < nu) ? func[(size_t)(row0 + r) * nu + col] : 0.0f;
        __syncwarp();
        const int n_cols = min(32, nu - k0);
        for (int j = 0; j < n_cols; ++j) { run = rn_add(run, rn_div(t[lane][j], nf)); t[lane][j] = run; }
        __syncwarp();
        for (int r = 0; r < 32; ++r) if (row0 + r < nv && col < nu) cdf[(size_t)(row0 + r) * (nu + 1) + col + 1] = t[r][lane];
        __syncwarp();
    }
    if (row0 + lane < nv) { integral[row0 + lane] = run; cdf[(size_t)(row0 + lane) * (nu + 1)] = 0.0f; }
}
__global__ void __launch_bounds__(256)
k_env_rows_normalise(int nu, int nv, const float* __restrict__ integral, float* __restrict__ cdf) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (size_t)nv * (nu + 1)) return;
    const int row = (int)(k / (size_t)(nu + 1)), i = (int)(k % (size_t)(nu + 1));
    if (i == 0) return;
    const float total = integral[row];
    cdf[k] = (total == 0.0f) ? rn_div((float)i, (float)nu) : rn_div(cdf[k], total);
}

static DeviceArena g_arena[FTN_MAX_DEVICES];
DeviceArena& device_arena(int device) { return g_arena[(device >= 0 && device < FTN_MAX_DEVICES) ? device : 0]; }
int DeviceArena::reserve(int which, size_t need, const char* what, void** out) {
    if (bytes[which] < need) {
        if (p[which]) { cudaFree(p[which]); p[which] = nullptr; bytes[which] = 0; }
        cudaError_t e = cudaMalloc(&p[which], need);
        if (e != cudaSuccess) { p[which] = nullptr; return cuda_fail(e, what, __FILE__, __LINE__); }
        bytes[which] = need;
    }
    *out = p[which];
    return FTN_OK;
}
int DeviceArena::pinned_counts(uint32_t** out, size_t need) {
    if (h_pinned_bytes < need) {
        if (h_pinned) { cudaFreeHost(h_pinned); h_pinned = nullptr; h_pinned_bytes = 0; }
        cudaError_t e = cudaHostAlloc(&h_pinned, need < 256 ? 256 : need, cudaHostAllocDefault);
        if (e != cudaSuccess) { h_pinned = nullptr; return cuda_fail(e, "cudaHostAlloc (counter read-back)", __FILE__, __LINE__); }
        h_pinned_bytes = need < 256 ? 256 : need;
    }
    *out = (uint32_t*)h_pinned;
    return FTN_OK;
}
// Scene buffers come from the device's stream-ordered pool (cudaMallocAsync on the legacy stream), kept mapped between
// scenes (release threshold = max): creating and destroying a scene per frame, as the end-to-end path does, costs no
// page mapping after the first one.  ftn_release_cached_memory trims the pool.
static bool g_pool_ready[FTN_MAX_DEVICES] = {};
static cudaError_t scene_malloc(void** p, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < FTN_MAX_DEVICES && !g_pool_ready[dev]) {
        cudaMemPool_t pool;
        if ((e = cudaDeviceGetDefaultMemPool(&pool, dev)) != cudaSuccess) return e;
        unsigned long long keep = ~0ull;
        if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
        g_pool_ready[dev] = true;
    }
    return cudaMallocAsync(p, bytes ? bytes : 1, 0);
}
template <class T> static cudaError_t scene_malloc(T** p, size_t bytes) { return scene_malloc((void**)p, bytes); }
static void scene_free(void* p) { if (p) cudaFreeAsync(p, 0); }

int release_cached_memory() {
    for (int d = 0; d < FTN_MAX_DEVICES; ++d) {
        DeviceArena& a = g_arena[d];
        std::lock_guard<std::recursive_mutex> lock(a.m);
        if (!a.p[0] && !a.p[1] && !a.p[2] && !a.p[3] && !a.h_pinned) continue;
        if (cudaSetDevice(d) != cudaSuccess) continue;
        for (int i = 0; i < 4; ++i) { cudaFree(a.p[i]); a.p[i] = nullptr; a.bytes[i] = 0; }
        if (a.h_pinned) { cudaFreeHost(a.h_pinned); a.h_pinned = nullptr; a.h_pinned_bytes = 0; }
    }
    for (int d = 0; d < FTN_MAX_DEVICES; ++d) {
        if (!g_pool_ready[d]) continue;
        cudaMemPool_t pool;
        if (cudaSetDevice(d) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) continue;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    return FTN_OK;
}

SceneView make_view(const FtnScene& s) {
    SceneView v;
    v.bvh.nodes = s.d_nodes; v.bvh.tris = s.d_tris; v.bvh.n_nodes = s.n_nodes; v.bvh.n_tris = s.n_tris; v.bvh.wide = s.wide ? 1u : 0u;
    v.pos = s.d_pos; v.nrm = s.d_nrm; v.uv = s.d_uv; v.idx = s.d_idx;
    v.meshes = s.d_meshes; v.materials = s.d_materials; v.textures = s.d_textures;
    v.spheres = s.d_spheres; v.n_spheres = s.n_spheres;
    v.lights = s.d_lights; v.n_lights = (uint32_t)s.h_lights.size();
    v.n_tris = s.n_tris;
    static const int thr = [] { const char* e = getenv("FTN_REFILL_THRESHOLD"); int t = e ? atoi(e) : FTN_REFILL_THRESHOLD_DEFAULT; return t < 1 ? 1 : (t > 32 ? 32 : t); }();
    v.refill_threshold = thr;
    // node step wins the per-step vote when 16 * #node lanes >= bias * #leaf lanes.  BVH8q (final code, 8 blocks per SM,
    // profiles/r02_ab_vote_bias.txt): 28 against 14 gives C4 +1.3 %, C3 coherent +4.4 %, incoherent diffuse +3.6 %
    static const int bias_env = [] { const char* e = getenv("FTN_VOTE_BIAS"); int t = e ? atoi(e) : 0; return t < 0 ? 0 : (t > 256 ? 256 : t); }();
    v.vote_bias = bias_env ? bias_env : (s.wide ? FTN_VOTE_BIAS_WIDE_DEFAULT : FTN_VOTE_BIAS_DEFAULT);
    // measured (profiles/r01_ab_vote_ldg256.txt): the vote pays on deep trees (1M triangles: +6..9 %) and costs
    // on shallow ones (4332 triangles: -9 %), where its per-step bookkeeping is not amortised
    static const int vote_env = [] { const char* e = getenv("FTN_TRAVERSE_VOTE"); return e ? atoi(e) : -1; }();
    v.vote = vote_env >= 0 ? vote_env != 0 : s.n_tris >= FTN_VOTE_MIN_TRIS;
    return v;
}

template <class T> static int upload(T** dst, const T* src, size_t count) {
    *dst = nullptr;
    if (count == 0) return FTN_OK;
    FTN_CUDA(scene_malloc((void**)dst, count * sizeof(T)));
    FTN_CUDA(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return FTN_OK;
}

static M4 to_m4(const float* f) { M4 m; std::memcpy(m.m, f, 64); return m; }

// microfacet.rs:40-45 (host, once per material: the table is scene data, not per-hit work)
static float roughness_to_alpha_host(float roughness) {
    float rough = std::fmax(roughness, 1.0e-3f);
    float x = std::log(rough);
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

__global__ void k_rgb_to_rgba(const float* __restrict__ rgb, size_t n, F4* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { F4 t; t.x = rgb[3 * i]; t.y = rgb[3 * i + 1]; t.z = rgb[3 * i + 2]; t.w = 0.0f; out[i] = t; }
}

// one thread per guide entry: a full binary search of its key in its row (paid once per scene, saves ~8 dependent loads
// per cdf search of every light sample)
__global__ void __launch_bounds__(256)
k_env_guide(const float* __restrict__ cdf, int n, int rows, uint32_t* __restrict__ guide) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (size_t)rows * (size_t)(n + 1)) return;
    const int row = (int)(k / (size_t)(n + 1)), g = (int)(k % (size_t)(n + 1));
    guide[k] = env_guide_entry(cdf + (size_t)row * (size_t)(n + 1), n, g);
}

static int build_env_light(FtnScene* s, const FtnLight& fl, LightData* out) {
    EnvLightData& e = out->env;
    e.w = fl.width; e.h = fl.height;
    e.nu = fl.height; e.nv = fl.width;   // infinite.rs:64: `let (height, width) = mipmap.resolution()`
    int mx = fl.width > fl.height ? fl.width : fl.height;
    int lv = 0; while ((1 << (lv + 1)) <= mx) ++lv;
    e.levels = 1 + lv;
    e.l2w = to_m4(fl.light_to_world); e.w2l = to_m4(fl.world_to_light);
    e.world_radius = 0.0f; e.world_center[0] = e.world_center[1] = e.world_center[2] = 0.0f;
    const size_t n = (size_t)fl.width * fl.height;
    // the caller's RGB floats go up as they are (25 MB for 2048x1024) and are widened to RGBA texels on the device
    F4* d_tex; float *d_rgb, *d_func, *d_cdf, *d_int, *d_mcdf;
    FTN_TRY(upload(&d_rgb, fl.texels, 3 * n));
    cudaError_t me = scene_malloc(&d_tex, n * sizeof(F4));
    if (me != cudaSuccess) { scene_free(d_rgb); return cuda_fail(me, "cudaMalloc env texels", __FILE__, __LINE__); }
    s->owned.push_back(d_tex);
    k_rgb_to_rgba<<<(unsigned)((n + 255) / 256), 256>>>(d_rgb, n, d_tex);
    FTN_LAUNCHED();
    scene_free(d_rgb);   // ordered after the kernel on the same (legacy) stream
    e.texels = d_tex;
    FTN_CUDA(scene_malloc(&d_func, n * sizeof(float))); s->owned.push_back(d_func);
    FTN_CUDA(scene_malloc(&d_cdf, (size_t)e.nv * (e.nu + 1) * sizeof(float))); s->owned.push_back(d_cdf);
    FTN_CUDA(scene_malloc(&d_int, (size_t)e.nv * sizeof(float))); s->owned.push_back(d_int);
    FTN_CUDA(scene_malloc(&d_mcdf, (size_t)(e.nv + 1) * sizeof(float))); s->owned.push_back(d_mcdf);
    e.cond_func = d_func; e.cond_cdf = d_cdf; e.cond_integral = d_int; e.marg_cdf = d_mcdf;
    k_env_func<<<(unsigned)((n + 255) / 256), 256>>>(e, d_func);
    FTN_LAUNCHED();
    k_env_rows_running_sum<<<(e.nv + 127) / 128, 128>>>(d_func, e.nu, e.nv, d_cdf, d_int);
    FTN_LAUNCHED();
    k_env_rows_normalise<<<(unsigned)(((size_t)e.nv * (e.nu + 1) + 255) / 256), 256>>>(e.nu, e.nv, d_int, d_cdf);
    FTN_LAUNCHED();
    // marginal = Distribution1D::new(row integrals): one more "row" of length nv
    float* d_mint;
    FTN_CUDA(scene_malloc(&d_mint, sizeof(float))); s->owned.push_back(d_mint);
    k_env_row_cdf<<<1, 128>>>(d_int, e.nv, 1, d_mcdf, d_mint);
    FTN_LAUNCHED();
    uint32_t *d_cguide, *d_mguide;
    FTN_CUDA(scene_malloc(&d_cguide, (size_t)e.nv * (e.nu + 1) * sizeof(uint32_t))); s->owned.push_back(d_cguide);
    FTN_CUDA(scene_malloc(&d_mguide, (size_t)(e.nv + 1) * sizeof(uint32_t))); s->owned.push_back(d_mguide);
    k_env_guide<<<(unsigned)(((size_t)e.nv * (e.nu + 1) + 255) / 256), 256>>>(d_cdf, e.nu, e.nv, d_cguide);
    FTN_LAUNCHED();
    k_env_guide<<<(unsigned)((e.nv + 1 + 255) / 256), 256>>>(d_mcdf, e.nv, 1, d_mguide);
    FTN_LAUNCHED();
    e.cond_guide = d_cguide; e.marg_guide = d_mguide;
    FTN_CUDA(cudaMemcpy(&e.marg_integral, d_mint, sizeof(float), cudaMemcpyDeviceToHost));
    return FTN_OK;
}

int scene_create(const FtnSceneDesc* d, FtnScene** out) {
    if (!d || !out) return set_error(FTN_ERR_INVALID_ARGUMENT, "null argument");
    if (d->abi_version != FTN_ABI_VERSION) return set_error(FTN_ERR_INVALID_ARGUMENT, "abi version mismatch");
    if (d->n_triangles && (!d->positions || !d->indices || !d->meshes)) return set_error(FTN_ERR_INVALID_ARGUMENT, "triangles without positions/indices/meshes");
    if (d->n_triangles >= (1u << 29)) return set_error(FTN_ERR_INVALID_ARGUMENT, "too many triangles (leaf refs hold 29 bits)");
    uint32_t covered = 0;
    for (uint32_t m = 0; m < d->n_meshes; ++m) {
        if (d->meshes[m].first_tri != covered) return set_error(FTN_ERR_INVALID_ARGUMENT, "meshes must tile the index buffer in order");
        if (d->meshes[m].material_id >= (int32_t)d->n_materials) return set_error(FTN_ERR_INVALID_ARGUMENT, "material id out of range");
        covered += d->meshes[m].n_tris;
        if (covered > d->n_triangles) return set_error(FTN_ERR_INVALID_ARGUMENT, "mesh range exceeds the index buffer");
    }
    if (covered != d->n_triangles) return set_error(FTN_ERR_INVALID_ARGUMENT, "meshes do not cover all triangles");
    for (size_t k = 0; k < 3 * (size_t)d->n_triangles; ++k)
        if (d->indices[k] >= d->n_vertices) return set_error(FTN_ERR_INVALID_ARGUMENT, "vertex index out of range");
    int dev = 0;
    FTN_CUDA(cudaGetDevice(&dev));
    FtnScene* s = new FtnScene();
    s->device = dev;
    s->n_verts = d->n_vertices; s->n_tris = d->n_triangles; s->n_meshes = d->n_meshes;
    s->n_spheres = d->n_spheres; s->n_materials = d->n_materials; s->n_lights = d->n_lights;
    int rc = FTN_OK;
    auto bail = [&](int code) { scene_destroy(s); return code; };
    if ((rc = upload(&s->d_pos, d->positions, 3 * (size_t)d->n_vertices)) != FTN_OK) return bail(rc);
    if (d->normals && (rc = upload(&s->d_nrm, d->normals, 3 * (size_t)d->n_vertices)) != FTN_OK) return bail(rc);
    if (d->uvs && (rc = upload(&s->d_uv, d->uvs, 2 * (size_t)d->n_vertices)) != FTN_OK) return bail(rc);
    if ((rc = upload(&s->d_idx, d->indices, 3 * (size_t)d->n_triangles)) != FTN_OK) return bail(rc);
    // the texture table (FtnSceneDesc::textures): image pyramids are copied as RGBA texels
    std::vector<TextureData> texs(d->n_textures);
    if (d->n_textures && !d->textures) return bail(set_error(FTN_ERR_INVALID_ARGUMENT, "n_textures != 0 but textures is null"));
    for (uint32_t t = 0; t < d->n_textures; ++t) {
        std::vector<F4> texels;
        if (texture_from_abi(d->textures[t], &texs[t], &texels) != FTN_OK) return bail(set_error(FTN_ERR_INVALID_ARGUMENT, "texture table: bad texture type or image pyramid"));
        if (texs[t].type == FTN_TEXTURE_IMAGE) {
            F4* d_img = nullptr;
            if ((rc = upload(&d_img, texels.data(), texels.size())) != FTN_OK) return bail(rc);
            s->owned.push_back(d_img);
            texs[t].image = d_img;
        }
    }
    if ((rc = upload(&s->d_textures, texs.data(), texs.size())) != FTN_OK) return bail(rc);
    std::vector<MaterialData> mats(d->n_materials);
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const FtnMaterial fm = fold_constant_param_textures(d->materials[m], d->textures, d->n_textures);
        MaterialData& md = mats[m];
        std::memset(&md, 0, sizeof(md));
        md.type = fm.type;
        if (fm.type < FTN_MATERIAL_MATTE || fm.type > FTN_MATERIAL_GLASS) return bail(set_error(FTN_ERR_INVALID_ARGUMENT, "unknown material type"));
        for (int c = 0; c < 3; ++c) { md.kd[c] = fm.kd[c]; md.ks[c] = fm.ks[c]; md.eta[c] = fm.eta[c]; md.k[c] = fm.k[c]; }
        float ur = fm.u_roughness, vr = fm.v_roughness;
        if (fm.type == FTN_MATERIAL_MIRROR) for (int c = 0; c < 3; ++c) md.kd[c] = fm.kr[c];   // Kr travels in the kd slot
        if (fm.type == FTN_MATERIAL_GLASS) for (int c = 0; c < 3; ++c) { md.kd[c] = fm.kr[c]; md.ks[c] = fm.kt[c]; }   // Kr, Kt; eta[0] = index
        md.kd_texture = (fm.type == FTN_MATERIAL_MATTE || fm.type == FTN_MATERIAL_PLASTIC || fm.type == FTN_MATERIAL_MIRROR) ? fm.kd_texture : 0;
        if (fm.type == FTN_MATERIAL_MATTE) {   // matte.rs:42-49: sigma clamped to [0, 90] degrees; != 0 -> OrenNayar::new (reflection/mod.rs:259-267)
            const float sigma = std::fmin(std::fmax(fm.sigma, 0.0f), 90.0f);
            if (sigma != 0.0f) {
                const float sr = sigma * (float)(3.14159265358979323846 / 180.0), s2 = sr * sr;
                md.type = FTN_CLASS_OREN_NAYAR;
                md.alpha_x = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));   // a
                md.alpha_y = 0.45f * s2 / (s2 + 0.09f);              // b
            }
        }
        for (int c = 0; c < 3; ++c) { md.tex1[c] = fm.tex1[c]; md.tex2[c] = fm.tex2[c]; }
        for (int c = 0; c < 2; ++c) { md.uv_scale[c] = fm.uv_scale[c]; md.uv_delta[c] = fm.uv_delta[c]; }
        md.image = nullptr; md.img_w = md.img_h = md.img_levels = md.img_wrap = 0;
        if (md.kd_texture < FTN_TEXTURE_CONSTANT || md.kd_texture > FTN_TEXTURE_IMAGE) return bail(set_error(FTN_ERR_INVALID_ARGUMENT, "unknown texture type"));
        if (md.kd_texture == FTN_TEXTURE_IMAGE) {   // the host's MIPMap pyramid (mipmap.rs:78-143), copied as RGBA texels
            std::vector<F4> texels;
            if (!pack_image_pyramid(fm, &texels)) return bail(set_error(FTN_ERR_INVALID_ARGUMENT, "image texture: bad pyramid description"));
            F4* d_img = nullptr;
            if ((rc = upload(&d_img, texels.data(), texels.size())) != FTN_OK) return bail(rc);
            s->owned.push_back(d_img);
            s->has_image_texture = true;
            md.image = d_img; md.img_w = fm.image_width; md.img_h = fm.image_height; md.img_levels = fm.image_levels; md.img_wrap = fm.image_wrap;
        }
        if (fm.type == FTN_MATERIAL_PLASTIC) vr = ur;
        if (fm.remap_roughness) { ur = roughness_to_alpha_host(ur); vr = roughness_to_alpha_host(vr); }
        if (md.type != FTN_CLASS_OREN_NAYAR) { md.alpha_x = ur; md.alpha_y = vr; }   // Oren-Nayar keeps (a, b) there
        if (material_params_from_abi(fm, d->n_textures, [&](uint32_t id) { return texs[id - 1].type == FTN_TEXTURE_IMAGE; }, &md) != FTN_OK)
            return bail(set_error(FTN_ERR_INVALID_ARGUMENT, "material: param_texture id out of range"));
        if (md.uses_image) s->has_image_texture = true;
        const bool rough_textured = md.ptex[FTN_PARAM_UROUGHNESS] || md.ptex[FTN_PARAM_VROUGHNESS];
        if (fm.type == FTN_MATERIAL_GLASS && !rough_textured && ur == 0.0f && vr == 0.0f)   // glass.rs:64-67: FresnelSpecular is todo!() in the reference
            return bail(set_error(FTN_ERR_UNSUPPORTED, "smooth glass (both alphas 0) is todo!() in the reference (glass.rs:66)"));
        s->material_present[md.type] = true;
    }
    for (uint32_t m = 0; m < d->n_meshes; ++m) if (d->meshes[m].material_id < 0 && d->meshes[m].n_tris) s->has_null_material = true;
    for (uint32_t i = 0; i < d->n_spheres; ++i) if (d->spheres[i].material_id < 0) s->has_null_material = true;
    if ((rc = upload(&s->d_materials, mats.data(), mats.size())) != FTN_OK) return bail(rc);
    // explicit lights first, then the primitives' area lights in primitive order (scene/mod.rs:32-49; the reference lists them
    // in ITS BVH's order -- the same set, and uniform_sample_one_light picks uniformly, so the estimator is the same)
    for (uint32_t l = 0; l < d->n_lights; ++l) {
        const FtnLight& fl = d->lights[l];
        if (fl.type == FTN_LIGHT_POINT || fl.type == FTN_LIGHT_DISTANT) {   // light/point.rs, light/distant.rs
            LightData ld; std::memset(&ld, 0, sizeof(ld));
            ld.type = fl.type == FTN_LIGHT_POINT ? 2 : 3; ld.sphere = -1;
            for (int c = 0; c < 3; ++c) { ld.emit[c] = fl.intensity[c]; ld.vec[c] = fl.type == FTN_LIGHT_POINT ? fl.point[c] : fl.direction[c]; }
            s->h_lights.push_back(ld);
            continue;
        }
        if (fl.type != FTN_LIGHT_INFINITE || fl.width < 1 || fl.height < 1 || !fl.texels) return bail(set_error(FTN_ERR_INVALID_ARGUMENT, "bad light"));
        LightData ld; std::memset(&ld, 0, sizeof(ld));
        ld.type = 0; ld.sphere = -1;
        if ((rc = build_env_light(s, fl, &ld)) != FTN_OK) return bail(rc);
        s->h_lights.push_back(ld);
    }
    // ... the area lights of emissive meshes (one per triangle) ...
    std::vector<MeshData> meshes;
    build_mesh_table(d, &meshes, &s->h_lights);
    if ((rc = upload(&s->d_meshes, meshes.data(), meshes.size())) != FTN_OK) return bail(rc);
    // ... then those of emissive spheres
    s->h_spheres.resize(d->n_spheres);
    for (uint32_t i = 0; i < d->n_spheres; ++i) {
        const FtnSphere& fs = d->spheres[i];
        SphereData& sd = s->h_spheres[i];
        if (fs.material_id >= (int32_t)d->n_materials) return bail(set_error(FTN_ERR_INVALID_ARGUMENT, "material id out of range"));
        sd.o2w = to_m4(fs.object_to_world); sd.w2o = to_m4(fs.world_to_object);
        // Sphere::new, sphere.rs:30-49
        const float r = fs.radius;
        sd.radius = r;
        sd.z_min = std::fmin(std::fmax(std::fmin(fs.z_min, fs.z_max), -r), r);
        sd.z_max = std::fmin(std::fmax(std::fmax(fs.z_min, fs.z_max), -r), r);
        sd.theta_min = std::acos(std::fmin(std::fmax(fs.z_min / r, -1.0f), 1.0f));
        sd.theta_max = std::acos(std::fmin(std::fmax(fs.z_max / r, -1.0f), 1.0f));
        sd.phi_max = std::fmin(std::fmax(fs.phi_max_deg, 0.0f), 360.0f) * (3.14159265358979323846f / 180.0f);
        sd.reverse_orientation = fs.reverse_orientation;
        sd.material = fs.material_id;
        sd.light = -1;
        sd.emit[0] = fs.emit[0]; sd.emit[1] = fs.emit[1]; sd.emit[2] = fs.emit[2];
        sd.area = sd.phi_max * sd.radius * (sd.z_max - sd.z_min);   // sphere.rs:77-79
        if (fs.emissive) {
            LightData ld; std::memset(&ld, 0, sizeof(ld));
            ld.type = 1; ld.sphere = (int32_t)i;
            ld.emit[0] = fs.emit[0]; ld.emit[1] = fs.emit[1]; ld.emit[2] = fs.emit[2];
            sd.light = (int)s->h_lights.size();
            s->h_lights.push_back(ld);
        }
    }
    if ((rc = upload(&s->d_spheres, s->h_spheres.data(), s->h_spheres.size())) != FTN_OK) return bail(rc);
    if ((rc = upload(&s->d_lights, s->h_lights.data(), s->h_lights.size())) != FTN_OK) return bail(rc);
    cudaError_t e = scene_malloc(&s->d_work, FTN_MAX_QUERIES_IN_FLIGHT * sizeof(unsigned long long));
    if (e != cudaSuccess) return bail(cuda_fail(e, "cudaMalloc work counter", __FILE__, __LINE__));
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return bail(cuda_fail(e, "scene_create sync", __FILE__, __LINE__));
    *out = s;
    return FTN_OK;
}

int scene_destroy(FtnScene* s) {
    if (!s) return FTN_OK;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != s->device) cudaSetDevice(s->device);
    cudaDeviceSynchronize();   // the frees below are ordered on the legacy stream only; nothing may still read the scene
    scene_free(s->d_pos); scene_free(s->d_nrm); scene_free(s->d_uv); scene_free(s->d_idx);
    scene_free(s->d_meshes); scene_free(s->d_materials); scene_free(s->d_textures); scene_free(s->d_spheres); scene_free(s->d_lights);
    scene_free(s->d_nodes); scene_free(s->d_tris); scene_free(s->d_codes); scene_free(s->d_order); scene_free(s->d_work);
    for (void* p : s->owned) scene_free(p);
    if (cur != s->device) cudaSetDevice(cur);
    delete s;
    return FTN_OK;
}

// Bounds3::join of the sphere bounds on the host: Shape::world_bound = o2w(object_bound),
// shapes/mod.rs:17-19, transform.rs:276-283, corner order bounds.rs:184-196 (8 points, min/max only).
static void sphere_world_bound(const SphereData& sd, float lo[3], float hi[3]) {
    const float omin[3] = {-sd.radius, -sd.radius, sd.z_min}, omax[3] = {sd.radius, sd.radius, sd.z_max};
    for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
    for (int c = 0; c < 8; ++c) {
        const float p[3] = {(c & 4) ? omax[0] : omin[0], (c & 2) ? omax[1] : omin[1], (c & 1) ? omax[2] : omin[2]};
        const float* m = sd.o2w.m;
        float q[4];
        for (int r = 0; r < 4; ++r) q[r] = ((m[r] * p[0] + m[4 + r] * p[1]) + m[8 + r] * p[2]) + m[12 + r] * 1.0f;
        const float iw = 1.0f / q[3];
        for (int a = 0; a < 3; ++a) { const float v = q[a] * iw; lo[a] = std::fmin(lo[a], v); hi[a] = std::fmax(hi[a], v); }
    }
}

int bvh_build(FtnScene* s) {
    if (!s) return set_error(FTN_ERR_INVALID_ARGUMENT, "null scene");
    if (s->built) return FTN_OK;
    FTN_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = 0;
    cudaEvent_t ev0, ev1, ev_s0, ev_s1;
    FTN_CUDA(cudaEventCreate(&ev0)); FTN_CUDA(cudaEventCreate(&ev1)); FTN_CUDA(cudaEventCreate(&ev_s0)); FTN_CUDA(cudaEventCreate(&ev_s1));
    bool sort_timed = false;
    FTN_CUDA(cudaEventRecord(ev0, st));
    const uint32_t n = s->n_tris;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int rc = FTN_OK;
    if (n > 0) {
        BuildBounds* d_gb = nullptr;
        F4 *tri_lo = nullptr, *tri_hi = nullptr, *leaf_lo = nullptr, *leaf_hi = nullptr;
        uint32_t *keys = nullptr, *survive = nullptr, *is_record = nullptr;
        LbvhArrays a; std::memset(&a, 0, sizeof(a));
        // Topology builder: PLOC for scenes of >= FTN_PLOC_MIN_TRIS triangles, the Karras radix tree below that
        // (and as the fallback for degenerate depth); FTN_BVH_BUILDER=ploc|lbvh overrides.  Measured on B200
        // (profiles/r01_ab_ploc.txt): PLOC +2..6 % rays/s on the 1M-triangle sphere, +15 % on the gear-ring scene,
        // +2.5 % on the 4332-triangle cube -- where its ~15 merge rounds (one host read-back each) cost 1.3 ms
        // of build time against 0.15 ms saved per render, hence the size threshold.
        const char* builder_env = getenv("FTN_BVH_BUILDER");
        const bool use_ploc = builder_env ? std::string(builder_env) == "ploc" : n >= FTN_PLOC_MIN_TRIS;
        // Node layout: BVH8q (compressed 8-wide, ftn_bvh8.cuh) for scenes of >= FTN_WIDE_MIN_TRIS triangles, BVH2x64 records
        // below (FTN_BVH_LAYOUT=bvh8|bvh2 overrides).  Measured on B200 (profiles/r02_ab_bvh8.txt): 1M triangles, incoherent
        // rays +4 % (diffuse bounce) / +23 % (interior); 147 k-triangle C4 render +2 %; 4332-triangle C2 render -5 %
        // (three levels of wide nodes: the 230-instruction node test is not amortised).
        const char* layout_env = getenv("FTN_BVH_LAYOUT");
        const bool wide = layout_env ? std::string(layout_env) != "bvh2" : n >= FTN_WIDE_MIN_TRIS;
        const size_t max_nodes8 = bvh8_max_nodes(n);
        // all temporaries come from the device's build arena: one (cached) allocation, no cudaFree
        // (each of which would synchronise the device) per build
        const size_t ni_max = n > 1 ? n - 1 : 1;
        auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t tmp_bytes = al(sizeof(BuildBounds)) + 4 * al((size_t)n * sizeof(F4)) + al((size_t)n * 4)
                               + 7 * al(ni_max * 4) + al((2 * (size_t)n - 1) * 4) + 2 * al(ni_max * sizeof(F4))
                               + al(radix_sort_scratch_bytes(n)) + al(scan_scratch_elems(ni_max) * 4) + 4096
                               + (use_ploc ? 9 * al((size_t)n * 4) + 2 * al((size_t)n * sizeof(F4)) + al(scan_scratch_elems(n) * 4) + 4096 : 0)
                               + (wide ? al(max_nodes8 * FTN_NODE8_BYTES) + al(max_nodes8 * 4) + al((size_t)n * 4) + al(BVH8_CTR_COUNT * 4) + 4096 : 0);
        DeviceArena& arena = device_arena(s->device);
        std::lock_guard<std::recursive_mutex> arena_lock(arena.m);
        char* tmp_base = nullptr; size_t tmp_off = 0;
        if ((rc = arena.reserve(DeviceArena::BUILD, tmp_bytes, "cudaMalloc (bvh build temporaries)", (void**)&tmp_base)) != FTN_OK) {
            cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev_s0); cudaEventDestroy(ev_s1); return rc;
        }
        auto dalloc = [&](void** p, size_t bytes) -> int {
            tmp_off = al(tmp_off);
            if (tmp_off + bytes > tmp_bytes) return set_error(FTN_ERR_OUT_OF_MEMORY, "bvh build arena under-sized");
            *p = tmp_base + tmp_off; tmp_off += bytes;
            return FTN_OK;
        };
        auto cleanup = [&]() {};
        void* sort_scratch = nullptr; uint32_t* scan_scratch = nullptr;
        const unsigned gb256 = (n + 255) / 256;
        do {
            if ((rc = dalloc((void**)&d_gb, sizeof(BuildBounds))) != FTN_OK) break;
            if ((rc = dalloc((void**)&tri_lo, (size_t)n * sizeof(F4))) != FTN_OK) break;
            if ((rc = dalloc((void**)&tri_hi, (size_t)n * sizeof(F4))) != FTN_OK) break;
            if ((rc = dalloc((void**)&keys, (size_t)n * 4)) != FTN_OK) break;
            cudaError_t e;
            if (!s->d_codes && (e = scene_malloc(&s->d_codes, (size_t)n * 4)) != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc codes", __FILE__, __LINE__); break; }
            if (!s->d_order && (e = scene_malloc(&s->d_order, (size_t)n * 4)) != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc order", __FILE__, __LINE__); break; }
            k_init_bounds<<<1, 32, 0, st>>>(d_gb); count_launch();
            k_tri_bounds<<<gb256, 256, 0, st>>>(s->d_pos, s->d_idx, n, tri_lo, tri_hi, d_gb); count_launch();
            if ((rc = dalloc(&sort_scratch, radix_sort_scratch_bytes(n))) != FTN_OK) break;
            cudaEventRecord(ev_s0, st);
            k_morton<<<gb256, 256, 0, st>>>(tri_lo, tri_hi, n, d_gb, s->d_codes, keys, s->d_order); count_launch();
            if ((e = cudaGetLastError()) != cudaSuccess) { rc = cuda_fail(e, "bounds/morton kernels", __FILE__, __LINE__); break; }
            if ((rc = radix_sort_pairs(keys, s->d_order, n, 30, sort_scratch, st)) != FTN_OK) break;
            cudaEventRecord(ev_s1, st);
            sort_timed = true;
            if ((rc = dalloc((void**)&leaf_lo, (size_t)n * sizeof(F4))) != FTN_OK) break;
            if ((rc = dalloc((void**)&leaf_hi, (size_t)n * sizeof(F4))) != FTN_OK) break;
            k_gather_leaf_boxes<<<gb256, 256, 0, st>>>(tri_lo, tri_hi, s->d_order, n, leaf_lo, leaf_hi); count_launch();
            if (!s->d_tris && (e = scene_malloc(&s->d_tris, (size_t)n * FTN_TRI_F4 * sizeof(F4))) != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc tris", __FILE__, __LINE__); break; }
            const uint32_t* final_order = s->d_order;   // leaf order of the emitted tree; PLOC: depth-first order of its tree
            // BVH8q: collapse the binary tree `a` (or, single_count != 0, wrap the whole scene into one leaf child) level by
            // level, then move the records into an exact-size allocation; sets final_order to the wide tree's triangle order
            auto collapse_wide = [&](const LbvhArrays& a8, const F4* llo, const F4* lhi, uint32_t single_count) -> int {
                F4* nodes8 = nullptr; uint32_t *brefs = nullptr, *order8 = nullptr, *ctr = nullptr;
                int r;
                if ((r = dalloc((void**)&nodes8, max_nodes8 * FTN_NODE8_BYTES)) != FTN_OK || (r = dalloc((void**)&brefs, max_nodes8 * 4)) != FTN_OK ||
                    (r = dalloc((void**)&order8, (size_t)n * 4)) != FTN_OK || (r = dalloc((void**)&ctr, BVH8_CTR_COUNT * 4)) != FTN_OK) return r;
                cudaError_t ce;
                const uint32_t one = 1u;
                if ((ce = cudaMemsetAsync(ctr, 0, BVH8_CTR_COUNT * 4, st)) != cudaSuccess || (ce = cudaMemsetAsync(brefs, 0, 4, st)) != cudaSuccess ||
                    (ce = cudaMemcpyAsync(ctr, &one, 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return cuda_fail(ce, "bvh8 counters", __FILE__, __LINE__);
                const unsigned g8 = (unsigned)std::min<size_t>((max_nodes8 + 127) / 128, (size_t)148 * 16);
                std::vector<uint32_t> h_ctr(BVH8_CTR_COUNT, 0u);
                uint32_t level = 0;
                for (;;) {
                    const uint32_t group = level == 0 ? 6u : 8u;    // levels launched per read-back of the counters
                    for (uint32_t g = 0; g < group; ++g, ++level) {
                        k_bvh8_level<<<g8, 128, 0, st>>>(a8, llo, lhi, level, ctr, nodes8, brefs, final_order, order8, (uint32_t)max_nodes8, single_count, d_gb); count_launch();
                    }
                    if ((ce = cudaMemcpyAsync(h_ctr.data(), ctr, BVH8_CTR_COUNT * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess || (ce = cudaStreamSynchronize(st)) != cudaSuccess)
                        return cuda_fail(ce, "bvh8 collapse", __FILE__, __LINE__);
                    if (h_ctr[BVH8_CTR_ERROR] != 0u) return set_error(FTN_ERR_CUDA, h_ctr[BVH8_CTR_ERROR] == 1u ? "bvh8 collapse: node bound exceeded" : "bvh8 collapse: tree deeper than the traversal stack");
                    if (h_ctr[level] == 0u) break;                  // the last level launched produced no children
                }
                uint32_t total = 0, levels = 0;
                for (uint32_t l = 0; l < (uint32_t)BVH8_MAX_LEVELS && h_ctr[l] != 0u; ++l) { total += h_ctr[l]; ++levels; }
                if (h_ctr[BVH8_CTR_TRIS] != n) return set_error(FTN_ERR_CUDA, "bvh8 collapse: triangle count mismatch");
                if (levels + 2u > (uint32_t)FTN_STACK8_SIZE) return set_error(FTN_ERR_CUDA, "bvh8 collapse: tree deeper than the traversal stack");
                if (getenv("FTN_DEBUG_BUILD")) fprintf(stderr, "[ftn] BVH8q: %u triangles, %u records in %u levels\n", n, total, levels);
                if (s->d_nodes) { scene_free(s->d_nodes); s->d_nodes = nullptr; }
                if ((ce = scene_malloc(&s->d_nodes, (size_t)total * FTN_NODE8_BYTES)) != cudaSuccess) return cuda_fail(ce, "cudaMalloc nodes", __FILE__, __LINE__);
                if ((ce = cudaMemcpyAsync(s->d_nodes, nodes8, (size_t)total * FTN_NODE8_BYTES, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return cuda_fail(ce, "copy nodes", __FILE__, __LINE__);
                s->n_nodes = total; s->bvh_levels = levels; s->wide = true;
                final_order = order8;
                return FTN_OK;
            };
            if (wide && n <= (uint32_t)FTN_LEAF8_MAX) {
                LbvhArrays none; std::memset(&none, 0, sizeof(none));
                if ((rc = collapse_wide(none, leaf_lo, leaf_hi, n)) != FTN_OK) break;
                k_gather_tris<<<gb256, 256, 0, st>>>(s->d_pos, s->d_idx, final_order, n, s->d_meshes, s->n_meshes, s->d_tris); count_launch();
            } else if (!wide && n <= (uint32_t)FTN_LEAF_MAX) {
                k_gather_tris<<<gb256, 256, 0, st>>>(s->d_pos, s->d_idx, final_order, n, s->d_meshes, s->n_meshes, s->d_tris); count_launch();
                if (!s->d_nodes && (e = scene_malloc(&s->d_nodes, FTN_NODE_F4 * sizeof(F4))) != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc nodes", __FILE__, __LINE__); break; }
                k_lbvh_emit_single<<<1, 32, 0, st>>>(n, d_gb, s->d_nodes); count_launch();
                s->n_nodes = 1;
            } else {
                const size_t ni = n - 1;
                if ((rc = dalloc((void**)&a.left, ni * 4)) != FTN_OK) break;
                if ((rc = dalloc((void**)&a.right, ni * 4)) != FTN_OK) break;
                if ((rc = dalloc((void**)&a.first, ni * 4)) != FTN_OK) break;
                if ((rc = dalloc((void**)&a.last, ni * 4)) != FTN_OK) break;
                if ((rc = dalloc((void**)&a.parent, (2 * (size_t)n - 1) * 4)) != FTN_OK) break;
                if ((rc = dalloc((void**)&a.arrive, ni * 4)) != FTN_OK) break;
                if ((rc = dalloc((void**)&a.node_lo, ni * sizeof(F4))) != FTN_OK) break;
                if ((rc = dalloc((void**)&a.node_hi, ni * sizeof(F4))) != FTN_OK) break;
                if ((rc = dalloc((void**)&survive, ni * 4)) != FTN_OK) break;
                if ((rc = dalloc((void**)&is_record, ni * 4)) != FTN_OK) break;
                const unsigned gi = (unsigned)((ni + 255) / 256);
                bool ploc_done = false;
                if (use_ploc) {
                    // ---- PLOC: merge mutual nearest neighbours of the Morton order until one cluster is left ----
                    uint32_t *cl0 = nullptr, *cl1 = nullptr, *nn = nullptr, *mg = nullptr, *va = nullptr, *ms = nullptr, *vs = nullptr, *newpos = nullptr, *order2 = nullptr, *d_small = nullptr;
                    uint32_t* pscan = nullptr; F4 *lo2 = nullptr, *hi2 = nullptr;
                    if ((rc = dalloc((void**)&cl0, (size_t)n * 4)) != FTN_OK || (rc = dalloc((void**)&cl1, (size_t)n * 4)) != FTN_OK ||
                        (rc = dalloc((void**)&nn, (size_t)n * 4)) != FTN_OK || (rc = dalloc((void**)&mg, (size_t)n * 4)) != FTN_OK ||
                        (rc = dalloc((void**)&va, (size_t)n * 4)) != FTN_OK || (rc = dalloc((void**)&ms, (size_t)n * 4)) != FTN_OK ||
                        (rc = dalloc((void**)&vs, (size_t)n * 4)) != FTN_OK || (rc = dalloc((void**)&newpos, (size_t)n * 4)) != FTN_OK ||
                        (rc = dalloc((void**)&order2, (size_t)n * 4)) != FTN_OK || (rc = dalloc((void**)&d_small, 64)) != FTN_OK ||
                        (rc = dalloc((void**)&pscan, scan_scratch_elems(n) * 4)) != FTN_OK ||
                        (rc = dalloc((void**)&lo2, (size_t)n * sizeof(F4))) != FTN_OK || (rc = dalloc((void**)&hi2, (size_t)n * sizeof(F4))) != FTN_OK) break;
                    if ((e = cudaMemsetAsync(a.parent, 0xFF, (2 * (size_t)n - 1) * 4, st)) != cudaSuccess) { rc = cuda_fail(e, "memset parent", __FILE__, __LINE__); break; }
                    if ((e = cudaMemsetAsync(d_small, 0, 64, st)) != cudaSuccess) { rc = cuda_fail(e, "memset", __FILE__, __LINE__); break; }
                    k_ploc_init<<<gb256, 256, 0, st>>>(n, cl0); count_launch();
                    uint32_t c = n, created = 0;
                    uint32_t *cin = cl0, *cout = cl1;
                    uint32_t* d_state = d_small + 8;            // {c, created} x 2 (double buffer)
                    const uint32_t h_state0[2] = {n, 0u};
                    if ((e = cudaMemcpyAsync(d_state, h_state0, 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) { rc = cuda_fail(e, "PLOC state", __FILE__, __LINE__); break; }
                    int parity = 0;
                    const int group = 4;                        // rounds per host read-back
                    while (c > (uint32_t)PLOC_FINISH_MAX && rc == FTN_OK) {
                        const unsigned gc = (c + 255) / 256;    // c: exact at the start of the group, an upper bound inside it
                        for (int r = 0; r < group && rc == FTN_OK; ++r) {
                            const uint32_t* st_in = d_state + 2 * parity;
                            uint32_t* st_out = d_state + 2 * (parity ^ 1);
                            k_ploc_nearest<<<gc, 256, 0, st>>>(a, leaf_lo, leaf_hi, cin, st_in, nn); count_launch();
                            k_ploc_flags<<<gc, 256, 0, st>>>(nn, st_in, c, mg, va); count_launch();
                            if ((rc = exclusive_scan_u32(mg, ms, c, pscan, st)) != FTN_OK) break;
                            if ((rc = exclusive_scan_u32(va, vs, c, pscan, st)) != FTN_OK) break;
                            k_ploc_merge<<<gc, 256, 0, st>>>(a, leaf_lo, leaf_hi, cin, cout, nn, mg, va, ms, vs, n, st_in, st_out); count_launch();
                            std::swap(cin, cout);
                            parity ^= 1;
                        }
                        if (rc != FTN_OK) break;
                        uint32_t h_state[2] = {0, 0};
                        if ((e = cudaMemcpyAsync(h_state, d_state + 2 * parity, 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess || (e = cudaStreamSynchronize(st)) != cudaSuccess) { rc = cuda_fail(e, "PLOC round", __FILE__, __LINE__); break; }
                        if (h_state[0] == 0u || h_state[0] >= c || h_state[0] + h_state[1] != n) { rc = set_error(FTN_ERR_CUDA, "PLOC rounds without progress"); break; }
                        c = h_state[0]; created = h_state[1];
                    }
                    if (rc != FTN_OK) break;
                    if (c > 1u) { k_ploc_finish<<<1, PLOC_FINISH_MAX, 0, st>>>(a, leaf_lo, leaf_hi, cin, c, n, created, d_small + 4); count_launch(); }
                    k_ploc_leaf_positions<<<gb256, 256, 0, st>>>(a, n, newpos, d_small + 1); count_launch();
                    uint32_t h_small[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    if ((e = cudaMemcpyAsync(h_small, d_small, sizeof(h_small), cudaMemcpyDeviceToHost, st)) != cudaSuccess || (e = cudaStreamSynchronize(st)) != cudaSuccess) { rc = cuda_fail(e, "PLOC depth", __FILE__, __LINE__); break; }
                    if (c > 1u && (h_small[4] != 0u || h_small[5] != (uint32_t)ni || h_small[6] != 0u)) { rc = set_error(FTN_ERR_CUDA, "PLOC finish kernel did not end at root 0"); break; }
                    const uint32_t max_depth = h_small[1];
                    const char* depth_env = getenv("FTN_PLOC_MAX_DEPTH");   // test hook: force the fallback
                    const uint32_t depth_limit = depth_env ? (uint32_t)atoi(depth_env) : (uint32_t)FTN_STACK_SIZE - 4u;
                    if (getenv("FTN_DEBUG_BUILD")) fprintf(stderr, "[ftn] PLOC: %u triangles, tree depth %u (limit %u)%s\n", n, max_depth, depth_limit, max_depth <= depth_limit ? "" : " -> radix-tree fallback");
                    if (max_depth <= depth_limit) {   // else: a degenerate chain; the radix tree below is depth-bounded
                        k_ploc_node_ranges<<<gi, 256, 0, st>>>(a, n); count_launch();
                        k_ploc_permute_leaves<<<gb256, 256, 0, st>>>(n, newpos, leaf_lo, leaf_hi, s->d_order, lo2, hi2, order2); count_launch();
                        k_ploc_rewrite_refs<<<gi, 256, 0, st>>>(a, n, newpos); count_launch();
                        leaf_lo = lo2; leaf_hi = hi2; final_order = order2;
                        k_ploc_survive<<<gi, 256, 0, st>>>((int)n, a, survive); count_launch();
                        ploc_done = true;
                    }
                }
                if (!ploc_done) {
                    if ((e = cudaMemsetAsync(a.arrive, 0, ni * 4, st)) != cudaSuccess) { rc = cuda_fail(e, "memset arrive", __FILE__, __LINE__); break; }
                    k_lbvh_topology<<<gi, 256, 0, st>>>(keys, (int)n, a); count_launch();
                    k_lbvh_refit<<<gb256, 256, 0, st>>>((int)n, a, leaf_lo, leaf_hi); count_launch();
                    k_lbvh_survive<<<gi, 256, 0, st>>>((int)n, a, survive); count_launch();
                }
                if ((e = cudaGetLastError()) != cudaSuccess) { rc = cuda_fail(e, "lbvh kernels", __FILE__, __LINE__); break; }
                if (wide) {
                    if ((rc = collapse_wide(a, leaf_lo, leaf_hi, 0u)) != FTN_OK) break;
                    k_gather_tris<<<gb256, 256, 0, st>>>(s->d_pos, s->d_idx, final_order, n, s->d_meshes, s->n_meshes, s->d_tris); count_launch();
                } else {
                k_gather_tris<<<gb256, 256, 0, st>>>(s->d_pos, s->d_idx, final_order, n, s->d_meshes, s->n_meshes, s->d_tris); count_launch();
                uint32_t last_flag = 0, last_idx = 0;
                k_lbvh_mark_records<<<gi, 256, 0, st>>>((int)n, a, survive, is_record); count_launch();
                if ((e = cudaMemcpyAsync(&last_flag, is_record + ni - 1, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { rc = cuda_fail(e, "read record flags", __FILE__, __LINE__); break; }
                uint32_t* new_index = a.arrive;   // reuse: arrival counters are dead after the refit
                if ((rc = dalloc((void**)&scan_scratch, scan_scratch_elems(ni) * 4)) != FTN_OK) break;
                if ((rc = exclusive_scan_u32(is_record, new_index, ni, scan_scratch, st)) != FTN_OK) break;
                if ((e = cudaMemcpyAsync(&last_idx, new_index + ni - 1, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) { rc = cuda_fail(e, "read scan", __FILE__, __LINE__); break; }
                if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { rc = cuda_fail(e, "lbvh sync", __FILE__, __LINE__); break; }
                s->n_nodes = last_idx + last_flag;
                if (s->d_nodes) { scene_free(s->d_nodes); s->d_nodes = nullptr; }
                if ((e = scene_malloc(&s->d_nodes, (size_t)s->n_nodes * FTN_NODE_F4 * sizeof(F4))) != cudaSuccess) { rc = cuda_fail(e, "cudaMalloc nodes", __FILE__, __LINE__); break; }
                k_lbvh_emit<<<gi, 256, 0, st>>>((int)n, a, leaf_lo, leaf_hi, survive, is_record, new_index, s->d_nodes); count_launch();
                }
            }
            if ((e = cudaGetLastError()) != cudaSuccess) { rc = cuda_fail(e, "emit kernels", __FILE__, __LINE__); break; }
            BuildBounds hb;
            if ((e = cudaMemcpyAsync(&hb, d_gb, sizeof(hb), cudaMemcpyDeviceToHost, st)) != cudaSuccess) { rc = cuda_fail(e, "read bounds", __FILE__, __LINE__); break; }
            if ((e = cudaStreamSynchronize(st)) != cudaSuccess) { rc = cuda_fail(e, "bvh build sync", __FILE__, __LINE__); break; }
            for (int c = 0; c < 3; ++c) {
                uint32_t ul = hb.scene_lo[c], uh = hb.scene_hi[c];
                uint32_t bl = (ul & 0x80000000u) ? (ul & 0x7FFFFFFFu) : ~ul, bh = (uh & 0x80000000u) ? (uh & 0x7FFFFFFFu) : ~uh;
                std::memcpy(&lo[c], &bl, 4); std::memcpy(&hi[c], &bh, 4);
            }
        } while (0);
        cleanup();
        if (rc != FTN_OK) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev_s0); cudaEventDestroy(ev_s1); return rc; }
    }
    for (const SphereData& sd : s->h_spheres) {
        float slo[3], shi[3];
        sphere_world_bound(sd, slo, shi);
        for (int c = 0; c < 3; ++c) { lo[c] = std::fmin(lo[c], slo[c]); hi[c] = std::fmax(hi[c], shi[c]); }
    }
    for (int c = 0; c < 3; ++c) { s->bounds[c] = lo[c]; s->bounds[3 + c] = hi[c]; }
    // Scene::new -> Light::preprocess (infinite.rs:93-97): bounding sphere of the world bound (bounds.rs:208-212)
    bool lights_changed = false;
    for (LightData& ld : s->h_lights) {
        if (ld.type != 0 && ld.type != 3) continue;   // infinite.rs:93-97, distant.rs:46-50
        float c[3], r2 = 0.0f;
        for (int a = 0; a < 3; ++a) c[a] = (lo[a] + hi[a]) / 2.0f;
        const float dx = hi[0] - c[0], dy = hi[1] - c[1], dz = hi[2] - c[2];
        r2 = (dx * dx + dy * dy) + dz * dz;
        ld.env.world_radius = std::sqrt(r2);
        ld.env.world_center[0] = c[0]; ld.env.world_center[1] = c[1]; ld.env.world_center[2] = c[2];
        lights_changed = true;
    }
    if (lights_changed) FTN_CUDA(cudaMemcpy(s->d_lights, s->h_lights.data(), s->h_lights.size() * sizeof(LightData), cudaMemcpyHostToDevice));
    FTN_CUDA(cudaEventRecord(ev1, st));
    FTN_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.0f;
    FTN_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    if (sort_timed) { float sms = 0.0f; if (cudaEventElapsedTime(&sms, ev_s0, ev_s1) == cudaSuccess) s->sort_seconds = sms * 1e-3; }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev_s0); cudaEventDestroy(ev_s1);
    s->build_seconds = ms * 1e-3;
    s->built = true;
    return FTN_OK;
}

}  // namespace ftn

ftn::SceneView FtnScene::view() const { return ftn::make_view(*this); }
