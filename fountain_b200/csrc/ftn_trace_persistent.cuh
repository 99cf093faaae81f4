// Persistent-warp traversal with per-lane ray replenishment (device only).
//
// ncu on the first version (one ray per thread, the warp refetching only when all 32 lanes were
// done) showed the traversal kernels ISSUE-bound, not memory-bound, with 6.8 (incoherent, 1M
// triangles) to 10.1 (C2) active threads per executed instruction out of 32
// (smsp__thread_inst_executed_per_inst_executed, profiles/r01_*): lanes whose ray had finished
// idled until the slowest lane of the warp was done.  Here every lane owns its ray state
// (node cursor, short stack in local memory, t_max, best hit); the warp runs
//     flush finished lanes -> refill idle lanes from a global work counter (one atomicAdd for all
//     idle lanes of the warp, __ballot_sync/__popc ranks) -> traverse
// and leaves the traverse loop as soon as fewer than `refill_threshold` lanes are still active.
// Inside the traverse loop interior-node steps and leaf steps are separate, warp-uniform phases:
// a lane that reaches a leaf waits for a leaf step instead of serialising against lanes that are
// still testing boxes, and the warp picks the phase most of its lanes are waiting for.
//
// The arithmetic per (ray, node) and (ray, triangle) is exactly that of ftn_bvh.cuh /
// ftn_geom.cuh (the single-ray bvh2_traverse stays as the host-testable statement of it).
#pragma once
#include "ftn_trace.cuh"

namespace ftn {

// Source:  __device__ bool load(uint32_t item, RayF* ray)       -- false: nothing to trace for this item
// Sink:    __device__ void store(bool valid, uint32_t item, const RayF& ray, const SceneHit& hit)
//          called by ALL 32 lanes together (valid = this lane has a finished ray), so it may use
//          warp collectives (queue_push).
// SPHERES = false compiles the EFloat sphere side list out of the kernel (scenes without spheres).
// MODE selects the layout and the form of the traverse loop (trace_mode(sc) picks it per scene):
//   FTN_MODE_WHILE  BVH2x64, while-while      FTN_MODE_VOTE  BVH2x64, per-step vote      FTN_MODE_WIDE  BVH8q (below)
#define FTN_MODE_WHILE 0
#define FTN_MODE_VOTE 1
#define FTN_MODE_WIDE 2
inline int trace_mode(const SceneView& sc) { return sc.bvh.wide ? FTN_MODE_WIDE : (sc.vote ? FTN_MODE_VOTE : FTN_MODE_WHILE); }
// runtime mode -> template argument M
#define FTN_MODE3(mode, CALL)                                                   \
    do {                                                                        \
        if ((mode) == FTN_MODE_WIDE) { constexpr int M = FTN_MODE_WIDE; CALL; } \
        else if ((mode) == FTN_MODE_VOTE) { constexpr int M = FTN_MODE_VOTE; CALL; } \
        else { constexpr int M = FTN_MODE_WHILE; CALL; }                        \
    } while (0)

template <bool ANY, bool COUNT, bool SPHERES, class Source, class Sink>
__device__ __forceinline__ void trace_persistent8(const SceneView& sc, uint32_t n_items, uint32_t* work_counter,
                                                  Source& src, Sink& sink, TraceCounters& tc);

template <bool ANY, bool COUNT, bool SPHERES, int MODE, class Source, class Sink>
__device__ __forceinline__ void trace_persistent(const SceneView& sc, uint32_t n_items, uint32_t* work_counter,
                                                 Source& src, Sink& sink, TraceCounters& tc) {
    if constexpr (MODE == FTN_MODE_WIDE) { trace_persistent8<ANY, COUNT, SPHERES>(sc, n_items, work_counter, src, sink, tc); return; }
    else {
    constexpr bool VOTE = MODE == FTN_MODE_VOTE;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const BvhView bvh = sc.bvh;
    bool has_ray = false, finished = false, exhausted = false;
    uint32_t item = 0;
    RayF ray; ray.o = v3s(0.0f); ray.d = v3s(0.0f); ray.t_max = 0.0f; ray.time = 0.0f;
    RaySlab slab; slab.o = v3s(0.0f); slab.inv_d = v3s(0.0f); slab.widen = 1.0f;
    RayShear shear; shear.kx = 0; shear.ky = 1; shear.kz = 2; shear.sx = shear.sy = shear.sz = 0.0f;
    SceneHit hit; hit.slot = FTN_NO_HIT_SLOT; hit.t = 0.0f; hit.tri.t = hit.tri.b0 = hit.tri.b1 = hit.tri.b2 = 0.0f;
    float t_max = 0.0f;
    int stack[FTN_STACK_SIZE];
    int sp = 0, cur = FTN_TRAVERSAL_DONE, leaf = 0;   // leaf < 0: a postponed leaf reference

    for (;;) {
        // ---- flush ----
        hit.t = t_max;
        sink.store(finished, item, ray, hit);
        if (finished) { has_ray = false; finished = false; }
        // ---- refill ----
        const unsigned idle = __ballot_sync(0xffffffffu, !has_ray);
        if (idle != 0u && !exhausted) {
            const int n_idle = __popc(idle), leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(work_counter, (uint32_t)n_idle);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!has_ray) {
                const uint32_t k = base + (uint32_t)__popc(idle & lt);
                if (k < n_items) {
                    item = k;
                    has_ray = true;
                    hit.slot = FTN_NO_HIT_SLOT;
                    if (!src.load(k, &ray)) { finished = true; t_max = ray.t_max; cur = FTN_TRAVERSAL_DONE; }
                    else {
                        t_max = ray.t_max;
                        // analytic spheres first, with the ray's own t_max (see ftn_trace.cuh)
                        if (SPHERES) {
                            for (uint32_t i = 0; i < sc.n_spheres; ++i) {
                                RayF r = ray; r.t_max = t_max;
                                SphereHit sh;
                                if (COUNT) tc.tris++;
                                if (sphere_intersect(sc.spheres[i], r, &sh)) { t_max = sh.t; hit.slot = FTN_SPHERE_SLOT_FLAG | i; if (ANY) break; }
                            }
                        }
                        if ((ANY && hit.slot != FTN_NO_HIT_SLOT) || bvh.n_nodes == 0u) { finished = true; cur = FTN_TRAVERSAL_DONE; }
                        else { slab = make_ray_slab(ray.o, ray.d); shear = make_ray_shear(ray.d); sp = 0; cur = 0; leaf = 0; }
                    }
                }
            }
            if (base + (uint32_t)n_idle >= n_items) exhausted = true;   // warp-uniform
        }
        if (__ballot_sync(0xffffffffu, has_ray) == 0u) break;
        // ---- traverse ----
        // One loop, one warp-uniform decision per step: the warp runs an interior-node step or a
        // leaf step, whichever more of its lanes are waiting for (vote by __ballot_sync/__popc).
        // A lane that reaches a leaf parks it in `leaf` and keeps walking nodes (speculative
        // traversal) until it meets a second one.  The SIMT model of this loop
        // (tests/hostsim sim_warp_model, scripts/warp_model.py) predicts 12.9 -> 20 lanes per
        // node step and 10.4 -> 13 per triangle test on the C3 incoherent batch against the
        // earlier "all lanes finish their nodes, then all leaves" form.
        const int thresh = exhausted ? 1 : sc.refill_threshold;
        if (VOTE) {
        const int bias = sc.vote_bias;                     // node step wins when 16 * #node lanes >= bias * #leaf lanes
        // Lane state is carried by (cur, leaf) alone: a lane without a ray or with a finished one has
        // cur == DONE and leaf == 0 and takes no part.
        for (;;) {
            if (leaf >= 0 && cur < 0 && cur != FTN_TRAVERSAL_DONE) {             // park the leaf, pop the next node
                leaf = cur;
                cur = (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE;
            }
            const unsigned m_node = __ballot_sync(0xffffffffu, cur >= 0), m_leaf = __ballot_sync(0xffffffffu, leaf < 0);
            // every unfinished lane wants one or the other; leave when the warp is too empty
            if (__popc(m_node | m_leaf) < thresh) break;
            if (16 * __popc(m_node) >= bias * __popc(m_leaf)) {
                if (cur >= 0) {
                    if (COUNT) tc.nodes++;
                    cur = node_step(bvh, cur, slab, t_max, stack, sp);
                }
            } else if (leaf < 0) {
                const bool stop = leaf_step<ANY, COUNT>(bvh, leaf, ray.o, shear, &t_max, &hit.slot, &hit.tri, &tc);
                leaf = 0;
                if (stop) cur = FTN_TRAVERSAL_DONE;
            }
        }
        if (has_ray && cur == FTN_TRAVERSAL_DONE && leaf >= 0) finished = true;
        } else {
        // while-while form: all lanes finish their node walk (parking one leaf), then all leaves
        for (;;) {
            const bool act = has_ray && !finished;
            while (act && cur >= 0) {
                if (COUNT) tc.nodes++;
                cur = node_step(bvh, cur, slab, t_max, stack, sp);
                if (cur < 0 && cur != FTN_TRAVERSAL_DONE && leaf >= 0) {
                    leaf = cur;
                    cur = (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE;
                }
            }
            while (act && leaf < 0) {
                const bool stop = leaf_step<ANY, COUNT>(bvh, leaf, ray.o, shear, &t_max, &hit.slot, &hit.tri, &tc);
                leaf = 0;
                if (stop) cur = FTN_TRAVERSAL_DONE;
                else if (cur < 0 && cur != FTN_TRAVERSAL_DONE) { leaf = cur; cur = (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE; }
            }
            if (act && cur == FTN_TRAVERSAL_DONE) finished = true;
            if (__popc(__ballot_sync(0xffffffffu, has_ray && !finished)) < thresh) break;
        }
        }
    }
    }
}

// ---- the same loop over the BVH8q layout (ftn_bvh8.cuh) -----------------------------------------------------------
// Lane state: a NODE GROUP (child_base, permuted hit bits | imask) = the interior children of one node still to visit,
// and a TRIANGLE GROUP (tri_base, permuted hit bits | counts) = its hit leaf children still to test.  A lane with a
// triangle group waits for a leaf step, any other unfinished lane for a node step; the warp votes per step as above.
// The stack holds node groups only: at most one entry per level of the tree (<= FTN_STACK8_SIZE, checked by the
// build), the first FTN_STACK8_SHARED of them in shared memory ([entry][thread]: conflict-free for any mix of depths),
// the rest in local memory.
template <bool ANY, bool COUNT, bool SPHERES, class Source, class Sink>
__device__ __forceinline__ void trace_persistent8(const SceneView& sc, uint32_t n_items, uint32_t* work_counter,
                                                  Source& src, Sink& sink, TraceCounters& tc) {
    __shared__ uint2 s_stack[FTN_STACK8_SHARED][FTN_TRACE_THREADS];
    uint2 l_stack[FTN_STACK8_SIZE - FTN_STACK8_SHARED];
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const BvhView bvh = sc.bvh;
    bool has_ray = false, finished = false, exhausted = false;
    uint32_t item = 0;
    RayF ray; ray.o = v3s(0.0f); ray.d = v3s(0.0f); ray.t_max = 0.0f; ray.time = 0.0f;
    Ray8 r8; r8.o = v3s(0.0f); r8.idir = v3s(1.0f); r8.octinv = 7u;
    RayShear shear; shear.kx = 0; shear.ky = 1; shear.kz = 2; shear.sx = shear.sy = shear.sz = 0.0f;
    SceneHit hit; hit.slot = FTN_NO_HIT_SLOT; hit.t = 0.0f; hit.tri.t = hit.tri.b0 = hit.tri.b1 = hit.tri.b2 = 0.0f;
    float t_max = 0.0f;
    uint32_t ng_base = 0u, ng_bits = 0u, tg_base = 0u, tg_bits = 0u;   // no hit bits in either group = nothing to do
    uint32_t tri_next = 0u, tri_left = 0u;                             // cursor into the leaf child being tested
    int sp = 0;

    for (;;) {
        // ---- flush ----
        hit.t = t_max;
        sink.store(finished, item, ray, hit);
        if (finished) { has_ray = false; finished = false; }
        // ---- refill ----
        const unsigned idle = __ballot_sync(0xffffffffu, !has_ray);
        if (idle != 0u && !exhausted) {
            const int n_idle = __popc(idle), leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(work_counter, (uint32_t)n_idle);
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!has_ray) {
                const uint32_t k = base + (uint32_t)__popc(idle & lt);
                if (k < n_items) {
                    item = k;
                    has_ray = true;
                    hit.slot = FTN_NO_HIT_SLOT;
                    ng_bits = 0u; tg_bits = 0u; tri_left = 0u; sp = 0;
                    if (!src.load(k, &ray)) { finished = true; t_max = ray.t_max; }
                    else {
                        t_max = ray.t_max;
                        if (SPHERES) {
                            for (uint32_t i = 0; i < sc.n_spheres; ++i) {
                                RayF r = ray; r.t_max = t_max;
                                SphereHit sh;
                                if (COUNT) tc.tris++;
                                if (sphere_intersect(sc.spheres[i], r, &sh)) { t_max = sh.t; hit.slot = FTN_SPHERE_SLOT_FLAG | i; if (ANY) break; }
                            }
                        }
                        if ((ANY && hit.slot != FTN_NO_HIT_SLOT) || bvh.n_nodes == 0u) finished = true;
                        else {
                            r8 = make_ray8(ray.o, ray.d); shear = make_ray_shear(ray.d);
                            ng_base = 0u; ng_bits = (1u << r8.octinv) | (1u << 8);   // the root as the only child of a virtual group
                        }
                    }
                }
            }
            if (base + (uint32_t)n_idle >= n_items) exhausted = true;   // warp-uniform
        }
        if (__ballot_sync(0xffffffffu, has_ray) == 0u) break;
        // ---- traverse ----
        const int thresh = exhausted ? 1 : sc.refill_threshold;
        const int bias = sc.vote_bias;
        for (;;) {
            const bool want_leaf = ((tg_bits & 0xFFu) | tri_left) != 0u;
            const bool want_node = !want_leaf && (ng_bits & 0xFFu) != 0u;
            const unsigned m_node = __ballot_sync(0xffffffffu, want_node), m_leaf = __ballot_sync(0xffffffffu, want_leaf);
            if (__popc(m_node | m_leaf) < thresh) break;
            if (16 * __popc(m_node) >= bias * __popc(m_leaf)) {
                if (want_node) {
                    const uint32_t node = node8_pop_child(ng_base, ng_bits, r8.octinv);
                    if (ng_bits & 0xFFu) {                                        // siblings left: keep the group for later
                        const uint2 e = make_uint2(ng_base, ng_bits);
                        if (sp < FTN_STACK8_SHARED) s_stack[sp][threadIdx.x] = e; else l_stack[sp - FTN_STACK8_SHARED] = e;
                        ++sp;
                    }
                    if (COUNT) tc.nodes++;
                    const Node8Hits h = node8_test(bvh.nodes, node, r8, t_max);
                    ng_base = h.child_base; ng_bits = h.ng_bits; tg_base = h.tri_base; tg_bits = h.tg_bits;
                }
            } else if (want_leaf) {
                // ONE triangle per leaf step: a lane walks through its leaf child (<= 3 triangles) with a cursor, so lanes
                // with short leaves do not idle behind long ones (SIMT model, scripts/warp_model8.py: 7.7 -> 12.7 lanes per
                // triangle test on the C3 incoherent batch, -11 % issue slots per ray)
                if (tri_left == 0u) node8_pop_leaf(tg_base, tg_bits, r8.octinv, &tri_next, &tri_left);
                const bool stop = tris8_test<ANY, COUNT>(bvh, tri_next, 1u, ray.o, shear, &t_max, &hit.slot, &hit.tri, &tc);
                ++tri_next; --tri_left;
                if (stop) { tg_bits = 0u; ng_bits = 0u; tri_left = 0u; sp = 0; }
            }
            // a lane with nothing left in either group takes the next group from its stack
            if (!(((tg_bits | ng_bits) & 0xFFu) | tri_left) && sp > 0) {
                --sp;
                const uint2 e = (sp < FTN_STACK8_SHARED) ? s_stack[sp][threadIdx.x] : l_stack[sp - FTN_STACK8_SHARED];
                ng_base = e.x; ng_bits = e.y;
            }
        }
        if (has_ray && !(((tg_bits | ng_bits) & 0xFFu) | tri_left) && sp == 0) finished = true;
    }
}

}  // namespace ftn
