// Batch ray queries against the aggregate: Scene::intersect / Scene::intersect_test
// (scene/mod.rs:51-57 -> bvh.rs:160-266) for whole ray batches.  One thread per ray, rays
// fetched with two 128-bit loads, dynamic warp-granular work fetch so long traversals do not
// leave SMs idle at the tail.
#include "ftn_scene.h"
#include <algorithm>
#include "ftn_shade.cuh"
#include "ftn_trace_persistent.cuh"

namespace ftn {

struct BatchSource {
    const FtnRay* rays;
    __device__ __forceinline__ bool load(uint32_t i, RayF* ray) const {
        // rays and hits stream through once: evict-first (ld.global.cs / st.global.cs) so that they
        // do not push BVH nodes and triangles out of the L2
        const float4 r0 = __ldcs(reinterpret_cast<const float4*>(rays + i));
        const float4 r1 = __ldcs(reinterpret_cast<const float4*>(rays + i) + 1);
        ray->o = V3(r0.x, r0.y, r0.z); ray->d = V3(r0.w, r1.x, r1.y); ray->t_max = r1.z; ray->time = r1.w;
        return true;
    }
};
template <bool ANY>
struct BatchSink {
    SceneView sc; FtnHit* hits; uint8_t* any_out;
    __device__ __forceinline__ void store(bool valid, uint32_t i, const RayF& ray, const SceneHit& h) const {
        if (!valid) return;
        if (ANY) { any_out[i] = (h.slot != FTN_NO_HIT_SLOT) ? 1 : 0; return; }
        uint32_t prim; float t = h.t, b1 = 0.0f, b2 = 0.0f;
        if (h.slot == FTN_NO_HIT_SLOT) { prim = FTN_NO_HIT; t = ray.t_max; }
        else if (h.slot & FTN_SPHERE_SLOT_FLAG) prim = sc.n_tris + (h.slot & ~FTN_SPHERE_SLOT_FLAG);
        else { prim = f2u(ld4(sc.bvh.tris + (size_t)FTN_TRI_F4 * (size_t)h.slot).w); b1 = h.tri.b1; b2 = h.tri.b2; }
        __stcs(reinterpret_cast<float4*>(hits) + i, make_float4(__uint_as_float(prim), t, b1, b2));
    }
};

template <bool ANY, bool COUNT, bool SPH, int MODE>
__global__ void FTN_TRACE_LAUNCH_BOUNDS
k_intersect_batch(SceneView sc, const FtnRay* __restrict__ rays, FtnHit* __restrict__ hits, uint8_t* __restrict__ any_out,
                  uint32_t n, uint32_t* __restrict__ work_counter, unsigned long long* __restrict__ counters) {
    BatchSource src; src.rays = rays;
    BatchSink<ANY> sink; sink.sc = sc; sink.hits = hits; sink.any_out = any_out;
    TraceCounters tc; tc.nodes = 0; tc.tris = 0;
    trace_persistent<ANY, COUNT, SPH, MODE>(sc, n, work_counter, src, sink, tc);
    if (COUNT) {
        unsigned long long nn = tc.nodes, tt = tc.tris;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { nn += __shfl_xor_sync(0xffffffffu, nn, o); tt += __shfl_xor_sync(0xffffffffu, tt, o); }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&counters[0], nn); atomicAdd(&counters[1], tt); }
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
    const SceneView sc = s->view();
    const size_t chunk = (size_t)1 << 30;   // the work counter is 32-bit
    for (size_t off = 0; off < n; off += chunk) {
        const uint32_t m = (uint32_t)std::min(chunk, n - off);
        // this call's own work counter: the next slot of the scene's ring (calls on different streams may overlap)
        uint32_t* d_work = reinterpret_cast<uint32_t*>(s->d_work + (s->work_slot.fetch_add(1u, std::memory_order_relaxed) % FTN_MAX_QUERIES_IN_FLIGHT));
        FTN_CUDA(cudaMemsetAsync(d_work, 0, sizeof(unsigned long long), st));
        const unsigned grid = trace_grid(m, sc.bvh.wide ? FTN_TRACE8_BLOCKS_PER_SM : FTN_TRACE_BLOCKS_PER_SM);
        const FtnRay* r = d_rays + off;
        FtnHit* h = d_hits ? d_hits + off : nullptr;
        uint8_t* a = d_any ? d_any + off : nullptr;
        const bool sph = s->n_spheres != 0, count = d_counters != nullptr;
        const int mode = trace_mode(sc);
        if (any) { FTN_MODE3(mode, FTN_BOOL1(sph, (k_intersect_batch<true, false, B0, M><<<grid, FTN_TRACE_THREADS, 0, st>>>(sc, r, nullptr, a, m, d_work, nullptr)))); }
        else if (count) { FTN_MODE3(mode, FTN_BOOL1(sph, (k_intersect_batch<false, true, B0, M><<<grid, FTN_TRACE_THREADS, 0, st>>>(sc, r, h, nullptr, m, d_work, d_counters)))); }
        else { FTN_MODE3(mode, FTN_BOOL1(sph, (k_intersect_batch<false, false, B0, M><<<grid, FTN_TRACE_THREADS, 0, st>>>(sc, r, h, nullptr, m, d_work, nullptr)))); }
        if (off + chunk < n) { FTN_LAUNCHED(); }
    }
    FTN_LAUNCHED();
    return FTN_OK;
}

}  // namespace ftn
