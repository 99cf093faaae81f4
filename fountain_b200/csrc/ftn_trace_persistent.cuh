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
#define FTN_TRAVERSAL_DONE ((int)0x80000000)

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
                const F4* nd = bvh.nodes + 4 * (size_t)cur;
                const F4 n0 = ld4(nd), n1 = ld4(nd + 1), nz = ld4(nd + 2), ci = ld4(nd + 3);
                if (COUNT) tc.nodes++;
                float e0, e1;
                const bool h0 = slab_test(slab, n0.x, n0.y, n0.z, n0.w, nz.x, nz.y, t_max, &e0);
                const bool h1 = slab_test(slab, n1.x, n1.y, n1.z, n1.w, nz.z, nz.w, t_max, &e1);
                const int c0 = (int)f2u(ci.x), c1 = (int)f2u(ci.y);
                if (h0 && h1) {
                    const bool swap = e1 < e0;
                    stack[sp++] = swap ? c0 : c1;
                    cur = swap ? c1 : c0;
                } else if (h0) cur = c0;
                else if (h1) cur = c1;
                else cur = (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE;
                if (cur < 0 && cur != FTN_TRAVERSAL_DONE && leaf >= 0) {
                    leaf = cur;
                    cur = (sp > 0) ? stack[--sp] : FTN_TRAVERSAL_DONE;
                }
            }
            // phase 2: the postponed leaf, then the second one if the lane stopped on it
            while (act && leaf < 0) {
                const uint32_t ref = ~(uint32_t)leaf;
                const uint32_t first = ref >> 2, count = (ref & 3u) + 1u;
                bool stop = false;
                for (uint32_t i = 0; i < count; ++i) {
                    const F4* t = bvh.tris + 3 * (size_t)(first + i);
                    const F4 a = ld4(t), b = ld4(t + 1), c = ld4(t + 2);
                    if (COUNT) tc.tris++;
                    TriHit h;
                    if (triangle_intersect(V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), ray.o, shear, t_max, &h)) {
                        t_max = h.t; hit.slot = first + i; hit.tri = h;
                        if (ANY) { stop = true; break; }
                    }
                }
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
