// Batch ray queries against the aggregate: Scene::intersect / Scene::intersect_test
// (scene/mod.rs:51-57 -> bvh.rs:160-266) for whole ray batches.  One thread per ray, rays
// fetched with two 128-bit loads, dynamic warp-granular work fetch so long traversals do not
// leave SMs idle at the tail.
#include "ftn_scene.h"
#include "ftn_shade.cuh"
#include "ftn_trace.cuh"

namespace ftn {

template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(FTN_TRACE_THREADS)
k_intersect_batch(SceneView sc, const FtnRay* __restrict__ rays, FtnHit* __restrict__ hits, uint8_t* __restrict__ any_out,
                  size_t n, unsigned long long* __restrict__ work_counter, unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    unsigned long long local_nodes = 0, local_tris = 0;
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(work_counter, 32ull);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const size_t i = base + lane;
        if (i < n) {
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(rays + i));
            const float4 r1 = __ldg(reinterpret_cast<const float4*>(rays + i) + 1);
            RayF ray; ray.o = V3(r0.x, r0.y, r0.z); ray.d = V3(r0.w, r1.x, r1.y); ray.t_max = r1.z; ray.time = r1.w;
            SceneHit h;
            TraceCounters tc; tc.nodes = 0; tc.tris = 0;
            scene_intersect<ANY, COUNT>(sc, ray, &h, &tc);
            if (COUNT) { local_nodes += tc.nodes; local_tris += tc.tris; }
            if (ANY) any_out[i] = (h.slot != FTN_NO_HIT_SLOT) ? 1 : 0;
            else {
                FtnHit out;
                if (h.slot == FTN_NO_HIT_SLOT) { out.prim = FTN_NO_HIT; out.t = ray.t_max; out.b1 = 0.0f; out.b2 = 0.0f; }
                else if (h.slot & FTN_SPHERE_SLOT_FLAG) { out.prim = sc.n_tris + (h.slot & ~FTN_SPHERE_SLOT_FLAG); out.t = h.t; out.b1 = 0.0f; out.b2 = 0.0f; }
                else { out.prim = f2u(ld4(sc.bvh.tris + 3 * (size_t)h.slot).w); out.t = h.t; out.b1 = h.tri.b1; out.b2 = h.tri.b2; }
                reinterpret_cast<float4*>(hits)[i] = make_float4(__uint_as_float(out.prim), out.t, out.b1, out.b2);
            }
        }
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { local_nodes += __shfl_xor_sync(0xffffffffu, local_nodes, o); local_tris += __shfl_xor_sync(0xffffffffu, local_tris, o); }
        if (lane == 0) { atomicAdd(&counters[0], local_nodes); atomicAdd(&counters[1], local_tris); }
    }
}

static int g_sm_count = 0;
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

// grid: a multiple of the SM count (148 on B200), capped by the work available
unsigned trace_grid(size_t n, int blocks_per_sm) {
    const size_t need = (n + FTN_TRACE_THREADS - 1) / FTN_TRACE_THREADS;
    const size_t full = (size_t)sm_count() * blocks_per_sm;
    return (unsigned)(need < full ? (need ? need : 1) : full);
}

int intersect_device(const FtnScene* s, size_t n, const FtnRay* d_rays, FtnHit* d_hits, uint8_t* d_any,
                     bool any, unsigned long long* d_counters, cudaStream_t st) {
    if (!s) return set_error(FTN_ERR_INVALID_ARGUMENT, "null scene");
    if (!s->built) return set_error(FTN_ERR_INVALID_ARGUMENT, "ftn_bvh_build has not been called");
    if (n == 0) return FTN_OK;
    unsigned long long* d_work = s->d_work;   // one query at a time per scene
    FTN_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), st));
    const SceneView sc = s->view();
    const unsigned grid = trace_grid(n, FTN_TRACE_BLOCKS_PER_SM);
    if (any) k_intersect_batch<true, false><<<grid, FTN_TRACE_THREADS, 0, st>>>(sc, d_rays, nullptr, d_any, n, d_work, nullptr);
    else if (d_counters) k_intersect_batch<false, true><<<grid, FTN_TRACE_THREADS, 0, st>>>(sc, d_rays, d_hits, nullptr, n, d_work, d_counters);
    else k_intersect_batch<false, false><<<grid, FTN_TRACE_THREADS, 0, st>>>(sc, d_rays, d_hits, nullptr, n, d_work, nullptr);
    FTN_LAUNCHED();
    return FTN_OK;
}

}  // namespace ftn
