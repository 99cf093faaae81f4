// Scene-level ray query: triangle BVH + the analytic sphere side list
// (Scene::intersect / intersect_test, scene/mod.rs:51-57).
#pragma once
#include "ftn_scene.h"
#include "ftn_geom.cuh"
#include "ftn_bvh.cuh"
#include "ftn_bvh8.cuh"

namespace ftn {

#define FTN_TRACE_THREADS 128
#ifndef FTN_TRACE_BLOCKS_PER_SM
#define FTN_TRACE_BLOCKS_PER_SM 8     /* 64 registers/thread: 8 x 128 threads fill the register file */
#endif
#ifndef FTN_TRACE8_BLOCKS_PER_SM
#define FTN_TRACE8_BLOCKS_PER_SM 7    /* BVH8q kernels: 72 registers/thread */
#endif
#ifdef FTN_TRACE_MIN_BLOCKS              /* A/B: force a register budget for more resident warps */
#define FTN_TRACE_LAUNCH_BOUNDS __launch_bounds__(FTN_TRACE_THREADS, FTN_TRACE_MIN_BLOCKS)
#else
#define FTN_TRACE_LAUNCH_BOUNDS __launch_bounds__(FTN_TRACE_THREADS)
#endif

struct SceneHit {
    uint32_t slot;     // FTN_NO_HIT_SLOT | triangle leaf-order slot | FTN_SPHERE_SLOT_FLAG + sphere index
    float t;
    TriHit tri;        // valid for triangle hits
};

// Spheres are tested first with the ray's own t_max, then the triangle BVH with t_max shrunk to
// the sphere hit: the same accept rule as the reference's single BVH over all primitives
// (a primitive wins iff its t is <= the best so far, primitive.rs:48-54).
template <bool ANY, bool COUNT>
FTN_HD void scene_intersect(const SceneView& sc, const RayF& ray, SceneHit* out, TraceCounters* ctr) {
    out->slot = FTN_NO_HIT_SLOT;
    float t_max = ray.t_max;
    for (uint32_t i = 0; i < sc.n_spheres; ++i) {
        RayF r = ray; r.t_max = t_max;
        SphereHit sh;
        if (COUNT) ctr->tris++;
        if (sphere_intersect(sc.spheres[i], r, &sh)) {
            t_max = sh.t; out->slot = FTN_SPHERE_SLOT_FLAG | i;
            if (ANY) { out->t = t_max; return; }
        }
    }
    TriHit th;
    const uint32_t slot = sc.bvh.wide ? bvh8_traverse<ANY, COUNT>(sc.bvh, ray.o, ray.d, &t_max, &th, ctr)
                                      : bvh_traverse<ANY, COUNT>(sc.bvh, ray.o, ray.d, &t_max, &th, ctr);
    if (slot != FTN_NO_HIT_SLOT) { out->slot = slot; out->tri = th; }
    out->t = t_max;
}

}  // namespace ftn
