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
// and leaves the traverse loop as soon as fewer than FTN_REFILL_THRESHOLD lanes are still active.
// Inside the traverse loop interior-node steps and leaf steps are separate phases (while-while):
// a lane that reaches a leaf waits for the leaf phase instead of serialising against lanes
// that are still testing boxes.
//
// The arithmetic per (ray, node) and (ray, triangle) is exactly that of ftn_bvh.cuh /
// ftn_geom.cuh (the single-ray bvh2_traverse stays as the host-testable statement of it).
#pragma once
#include "ftn_trace.cuh"

namespace ftn {

#ifndef FTN_REFILL_THRESHOLD
#define FTN_REFILL_THRESHOLD 20
#endif

// Source:  __device__ bool load(uint32_t item, RayF* ray)       -- false: nothing to trace for this item
// Sink:    __device__ void store(bool valid, uint32_t item, const RayF& ray, const SceneHit& hit)
//          called by ALL 32 lanes together (valid = this lane has a finished ray), so it may use
//          warp collectives (queue_push).
// SPHERES = false compiles the EFloat sphere side list out of the kernel (scenes without spheres).
template <bool ANY, bool COUNT, bool SPHERES, class Source, class Sink>
__device__ __forceinline__ void trace_persistent(const SceneView& sc, uint32_t n_items, uint32_t* work_counter,
                                                 Source& src, Sink& sink, TraceCounters& tc) {
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
        const int thresh = exhausted ? 1 : sc.refill_threshold;
        for (;;) {
            const bool act = has_ray && !finished;
            // phase 1: interior nodes.  The first leaf a lane meets is POSTPONED (speculative
            // traversal): the lane keeps walking until it meets a second leaf or runs out of nodes, so
            // lanes wait for each other only every other leaf.
            while (act && cur >= 0) {
                if (COUNT) tc.nodes++;
                cur = node_step(bvh, cur, slab, t_max, stack, sp);
                if (cur < 0 && cur != FTN_TRAVERSAL_DONE && leaf >= 0) {
                    leaf = cur;
                    cur = (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE;
                }
            }
            // phase 2: the postponed leaf, then the second one if the lane stopped on it
            while (act && leaf < 0) {
                const bool stop = leaf_step<ANY, COUNT>(bvh, leaf, ray.o, shear, &t_max, &hit.slot, &hit.tri, &tc);
                leaf = 0;
                if (stop) cur = FTN_TRAVERSAL_DONE;
                else if (cur < 0 && cur != FTN_TRAVERSAL_DONE) { leaf = cur; cur = (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE; }
            }
            if (act && cur == FTN_TRAVERSAL_DONE) finished = true;
            // phase 3: leave when the warp is too empty (idle lanes then flush + refill)
            if (__popc(__ballot_sync(0xffffffffu, has_ray && !finished)) < thresh) break;
        }
    }
}

}  // namespace ftn
