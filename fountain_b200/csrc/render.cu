// Wavefront path tracer: replaces SamplerIntegrator::render_parallel (integrator/mod.rs:218-283)
// and PathIntegrator::incident_radiance (integrator/path.rs:25-95) with one kernel per stage:
//
//   k_raygen        camera samples -> primary rays            (sampler/mod.rs:43-51, camera/mod.rs:117-143)
//   k_extend        closest hit for the active queue, paths binned into per-material queues
//   k_shade_miss    environment radiance for escaped paths    (path.rs:45-51, scene/mod.rs:58-64)
//   k_shade_null    surfaces without a BSDF: respawn           (path.rs:76-80)
//   k_shade<MAT>    one launch per material class over its queue (material-sorted shading):
//                   emission, light sample + BSDF eval, MIS BSDF sample, continuation sample, RR
//                   (path.rs:54-92, integrator/mod.rs:289-395)
//   k_shadow        any-hit for the light-sample rays          (light/mod.rs:82-84)
//   k_mis           closest/any-hit for the BSDF-sample rays + light lookup (integrator/mod.rs:364-389)
//   k_film_accumulate / k_film_resolve   box-filter footprint rule of Film::add_sample_to_tile and
//                   merge_film_tile (film.rs:121-172), as a deterministic per-pixel gather
//
// Queues hold path ids and are compacted with warp-aggregated atomics (__match_any_sync groups
// the lanes by destination queue, one atomicAdd per group, __shfl_sync broadcasts the base).
#include "ftn_scene.h"
#include "ftn_path.cuh"
#include "ftn_trace_persistent.cuh"
#include <algorithm>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>

namespace ftn {

int sm_count();
unsigned trace_grid(size_t n, int blocks_per_sm);
// Persistent grid of one traversal-kernel instantiation: as many blocks per SM as the kernel's registers / shared memory
// really allow (64 registers -> 8 blocks of 128 threads, 72 -> 7; asked from the runtime once per instantiation), capped by
// the work.  A fixed 7 left the 64-register BVH8q kernels one block per SM short (profiles/r02_ab_trace_grid.txt).
template <class Kernel>
static unsigned persistent_grid(Kernel kernel, size_t n) {
    // keyed by the kernel's ADDRESS (instantiations of one template share a function type); every device of a box is the same GPU
    static std::mutex m;
    static std::map<const void*, int> cache;
    int blocks;
    {
        std::lock_guard<std::mutex> lock(m);
        auto it = cache.find((const void*)kernel);
        if (it == cache.end()) {
            int b = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, FTN_TRACE_THREADS, 0) != cudaSuccess || b < 1) { cudaGetLastError(); b = FTN_TRACE8_BLOCKS_PER_SM; }
            const char* cap = getenv("FTN_TRACE_MAX_BLOCKS_PER_SM");   // A/B knob
            if (cap && atoi(cap) > 0 && b > atoi(cap)) b = atoi(cap);
            it = cache.emplace((const void*)kernel, b).first;
        }
        blocks = it->second;
    }
    return trace_grid(n, blocks);
}

enum { Q_ACTIVE_OUT = 0, Q_MISS, Q_NULL, Q_MAT0, Q_MAT1, Q_MAT2, Q_MAT3, Q_MAT4, Q_MAT5, Q_SHADOW, Q_MIS, Q_COUNT };
static_assert(Q_MAT5 - Q_MAT0 + 1 == FTN_N_CLASSES, "one shade queue per material class");
enum { W_EXTEND = Q_COUNT, W_SHADOW, W_MIS, CTR_COUNT };

struct PathArrays {
    float4* ray_o;      // origin.xyz, time
    float4* ray_d;      // dir.xyz, t_max
    uint32_t* hit;      // slot of the closest hit (FTN_NO_HIT_SLOT on a miss)
    float4* beta;       // throughput rgb, -
    float4* L;          // radiance rgb, -
    uint32_t* state;    // bits 0..15 bounces, bit 16 specular_bounce
    float4* sh_o;       // shadow ray origin.xyz, -
    float4* sh_d;       // shadow ray dir.xyz (target - origin), -
    float4* sh_L;       // contribution added when unoccluded
    float4* mis_o;      // MIS ray origin
    float4* mis_d;      // MIS ray dir
    float4* mis_w;      // weight rgb, light index (bits)
    float2* p_film;     // CameraSample.p_film
    float4* diff[4];    // rx_origin, rx_dir, ry_origin, ry_dir of the ray differential a specular reflection left on the path
                        // (FTN_STATE_HAS_DIFF); allocated only for the direct-lighting integrator on scenes with image textures
    uint8_t* spill;     // per sample pixel: some sample's footprint is not exactly its own pixel
};

__device__ __forceinline__ V3 ld3(const float4* p, size_t i) { const float4 v = p[i]; return V3(v.x, v.y, v.z); }
__device__ __forceinline__ void st3(float4* p, size_t i, V3 v, float w = 0.0f) { p[i] = make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ RayF load_ray(const PathArrays& pa, uint32_t path) {
    const float4 o = pa.ray_o[path], d = pa.ray_d[path];
    RayF ray; ray.o = V3(o.x, o.y, o.z); ray.d = V3(d.x, d.y, d.z); ray.t_max = d.w; ray.time = o.w;
    return ray;
}

// Warp-aggregated queue append: __match_any_sync groups the lanes by destination queue, the
// lowest lane of each group does ONE atomicAdd for the group, __shfl_sync broadcasts the base.
// All 32 lanes call; lanes with target < 0 push nothing.
__device__ __forceinline__ void queue_push(uint32_t* const* queues, uint32_t* counts, int target, uint32_t value) {
    const unsigned peers = __match_any_sync(0xffffffffu, target);
    if (target >= 0) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&counts[target], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        queues[target][base + rank] = value;
    }
}

struct Queues { uint32_t* q[Q_COUNT]; };

// The three appends of the shade stage (next-bounce, shadow and MIS queues) with ONE atomic round trip:
// lanes 0..2 reserve the three ranges in a single atomic instruction (the three sequential queue_push calls
// left 11 % of k_shade's stall samples waiting on their atomics).  All 32 lanes call.
__device__ __forceinline__ void queue_push3(const Queues& qs, uint32_t* counts, bool to_active, bool to_shadow, bool to_mis, uint32_t path) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned m_a = __ballot_sync(0xffffffffu, to_active), m_s = __ballot_sync(0xffffffffu, to_shadow), m_m = __ballot_sync(0xffffffffu, to_mis);
    uint32_t base = 0;
    if (lane < 3) {
        const unsigned m = lane == 0 ? m_a : (lane == 1 ? m_s : m_m);
        const int q = lane == 0 ? Q_ACTIVE_OUT : (lane == 1 ? Q_SHADOW : Q_MIS);
        if (m) base = atomicAdd(&counts[q], (uint32_t)__popc(m));
    }
    const uint32_t b_a = __shfl_sync(0xffffffffu, base, 0), b_s = __shfl_sync(0xffffffffu, base, 1), b_m = __shfl_sync(0xffffffffu, base, 2);
    if (to_active) qs.q[Q_ACTIVE_OUT][b_a + __popc(m_a & lt)] = path;
    if (to_shadow) qs.q[Q_SHADOW][b_s + __popc(m_s & lt)] = path;
    if (to_mis) qs.q[Q_MIS][b_m + __popc(m_m & lt)] = path;
}

// node-visit / triangle-test totals of the counting variants (the bytes-per-ray roofline input)
__device__ __forceinline__ void flush_trace_counters(const TraceCounters& tc, unsigned long long* trav) {
    unsigned long long n = tc.nodes, t = tc.tris;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); t += __shfl_xor_sync(0xffffffffu, t, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&trav[0], n); atomicAdd(&trav[1], t); }
}

// ---- raygen ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_raygen(PassParams pp, PathArrays pa) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pp.n_paths) return;
    float fx, fy; bool spills;
    const RayF ray = raygen_path(pp, i, &fx, &fy, &spills);
    if (spills) pa.spill[i / (uint32_t)pp.s_count] = 1;   // benign race: every writer stores 1
    st3(pa.ray_o, i, ray.o, ray.time);
    st3(pa.ray_d, i, ray.d, ray.t_max);
    st3(pa.beta, i, v3s(1.0f));
    st3(pa.L, i, v3s(0.0f));
    pa.state[i] = 0u;
    pa.p_film[i] = make_float2(fx, fy);
}

// ---- extend: closest hit + binning by material class ----------------------------------------------------------
// queue_in == nullptr: the identity queue (first bounce of a pass)
struct PathRaySource {
    PathArrays pa; const uint32_t* queue;
    __device__ __forceinline__ bool load(uint32_t k, RayF* ray) const {
        const uint32_t path = queue ? queue[k] : k;
        *ray = load_ray(pa, path);
        return true;
    }
};
struct ExtendSink {
    SceneView sc; PathArrays pa; const uint32_t* queue; Queues qs; uint32_t* counts; bool miss_always;
    __device__ __forceinline__ void store(bool valid, uint32_t k, const RayF&, const SceneHit& h) const {
        int target = -1; uint32_t path = 0;
        if (valid) {
            path = queue ? queue[k] : k;
            pa.hit[path] = h.slot;
            if (h.slot == FTN_NO_HIT_SLOT) {
                // an escaped path adds environment radiance only at bounce 0 or after a specular bounce
                // (path.rs:45-51); otherwise it simply ends here and needs no miss kernel
                const uint32_t st = pa.state[path];
                if (miss_always || (st & FTN_STATE_BOUNCES) == 0u || (st & FTN_STATE_SPECULAR)) target = Q_MISS;
            } else {
                const int material = hit_material(sc, h.slot);
                target = (material < 0) ? Q_NULL : (Q_MAT0 + sc.materials[material].type);
            }
        }
        queue_push(qs.q, counts, target, path);   // all 32 lanes arrive here together
    }
};
template <bool COUNT, bool SPH, int MODE>
__global__ void FTN_TRACE_LAUNCH_BOUNDS
k_extend(SceneView sc, PathArrays pa, const uint32_t* __restrict__ queue_in, uint32_t n_in, Queues qs, uint32_t* __restrict__ counts,
         unsigned long long* __restrict__ trav, bool miss_always) {
    PathRaySource src; src.pa = pa; src.queue = queue_in;
    ExtendSink sink; sink.sc = sc; sink.pa = pa; sink.queue = queue_in; sink.qs = qs; sink.counts = counts; sink.miss_always = miss_always;
    TraceCounters tc; tc.nodes = 0; tc.tris = 0;
    trace_persistent<false, COUNT, SPH, MODE>(sc, n_in, &counts[W_EXTEND], src, sink, tc);
    if (COUNT) flush_trace_counters(tc, trav);
}

// ---- miss: Scene::environment_emitted_radiance (path.rs:45-51, scene/mod.rs:58-64) -----------------------------
__global__ void __launch_bounds__(256)
k_shade_miss(SceneView sc, PassParams pp, PathArrays pa, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ counts) {
    const uint32_t n = counts[Q_MISS];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t path = queue[k];
        const uint32_t st = pa.state[path];
        const bool add = (pp.integrator == FTN_INTEGRATOR_DIRECT_LIGHTING) || ((st & FTN_STATE_BOUNCES) == 0u) || (st & FTN_STATE_SPECULAR);
        if (!add) continue;
        const V3 le = scene_env_radiance(sc, ld3(pa.ray_d, path));
        st3(pa.L, path, ld3(pa.L, path) + ld3(pa.beta, path) * le);
    }
}

// ---- material-sorted shading: one launch per material class (QUEUE = Q_NULL, Q_MAT0..2) -------------------------
// 4 blocks of 128 threads per SM = a cap of 128 registers: the conductor / plastic shaders would take 162-166
// (3 blocks, 18 % occupancy); capped they spill a little and run 4.5 % faster on C4 (3550 -> 3710 Mrays/s).
// Caps of 5 / 6 blocks were slower (profiles/r01_ab_vote_ldg256.txt).
#ifndef FTN_SHADE_MIN_BLOCKS
#define FTN_SHADE_MIN_BLOCKS 4
#endif
#define FTN_SHADE_LAUNCH_BOUNDS __launch_bounds__(128, FTN_SHADE_MIN_BLOCKS)
// IMG: the scene holds an image texture -- only then is the mip lookup (and its call frame) compiled into the
// matte / plastic / Oren-Nayar shaders, so scenes without one run the kernels they always ran
// The head of the shader's dependent-load chain.  The shaders are latency-bound at 4 blocks per SM: ncu's source page
// (profiles/r02_ncu_c4_shade_source.txt) puts 35-40 % of k_shade<metal>'s stall samples on queue[k] -> path -> the scattered
// state arrays -> the triangle record, long-scoreboard waits on a chain that does not depend on the shading itself.
//   * prefetch.global.L1 / .L2 of the next paths' state (round-2 experiment, removed): -8 % on C4 -- a prefetch moves whole
//     lines where the loads move 32-byte sectors and the kernel already draws 24 % of the DRAM peak
//     (profiles/r02_ab_shade_prefetch.txt);
//   * FTN_SHADE_STAGE = 1 (A/B, measured, left OFF): the queue entry is loaded two iterations ahead (one register) and the
//     NEXT path's state (ray origin / direction, beta, hit slot, state word: 56 B) is copied global -> shared with cp.async
//     while the current path is shaded -- sector-granular like a load, no registers held across the shading body, no warp
//     stalled on it; two 7-KB buffers per block of 128 threads (each thread reads and refills only its own slots: no
//     barrier).  B200 (profiles/r02_ab_shade_prefetch.txt): +7 % on C2's Lambert shader (0.251 -> 0.234 ms per launch), -4.5 %
//     on C4 (3.75 -> 3.92 ms): with the image light the shader's time goes to scattered L2 / DRAM sectors of the 48 MB of
//     environment tables, and the 56 KB of shared memory per SM come out of the L1 that caches them.
#ifndef FTN_SHADE_STAGE
#define FTN_SHADE_STAGE 0
#endif
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
struct ShadeStage { float4 ray_o[128], ray_d[128], beta[128]; uint32_t hit[128], state[128]; };
__device__ __forceinline__ void stage_path_state(ShadeStage& st, const PathArrays& pa, uint32_t path) {
    const int t = threadIdx.x;
    cp_async16(&st.ray_o[t], pa.ray_o + path); cp_async16(&st.ray_d[t], pa.ray_d + path); cp_async16(&st.beta[t], pa.beta + path);
    cp_async4(&st.hit[t], pa.hit + path); cp_async4(&st.state[t], pa.state + path);
}
// The kernels' sink of estimate_direct's two rays (ftn_path.cuh DirectOut): straight into the path's shadow / MIS slots
struct PathDirectSink {
    PathArrays pa; uint32_t path; bool has_shadow, has_mis;
    __device__ __forceinline__ void reset() { has_shadow = false; has_mis = false; }
    // the hit's own emission (beta * Le): the path's radiance is read and written back only here, i.e. only where a hit emits
    __device__ __forceinline__ void emitted(V3 e) { if (!is_black(e)) st3(pa.L, path, ld3(pa.L, path) + e); }
    __device__ __forceinline__ void shadow(V3 o, V3 d, V3 L) { has_shadow = true; st3(pa.sh_o, path, o); st3(pa.sh_d, path, d); st3(pa.sh_L, path, L); }
    __device__ __forceinline__ void mis(V3 o, V3 d, V3 w, int light) {
        has_mis = true; st3(pa.mis_o, path, o); st3(pa.mis_d, path, d); st3(pa.mis_w, path, w, u2f((uint32_t)light));
    }
};
template <int QUEUE, bool IMG = false>
__global__ void FTN_SHADE_LAUNCH_BOUNDS
k_shade(SceneView sc, PassParams pp, PathArrays pa, const uint32_t* __restrict__ queue, Queues qs, uint32_t* __restrict__ counts, uint32_t* __restrict__ err) {
    const uint32_t n = counts[QUEUE];
    const uint32_t n32 = (n + 31u) & ~31u;
    const uint32_t stride = gridDim.x * blockDim.x;   // n < 2^31 and stride < 2^23: k + 2 stride does not wrap
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
#if FTN_SHADE_STAGE
    __shared__ ShadeStage stage[2];
    uint32_t p0 = k < n ? queue[k] : 0u;                               // this iteration's path
    uint32_t p1 = k + stride < n ? queue[k + stride] : 0u;            // the next one
    int buf = 0;
    if (k < n) stage_path_state(stage[0], pa, p0);
    cp_async_commit();
#endif
    for (; k < n32; k += stride) {
        int t_active = -1, t_shadow = -1, t_mis = -1;
        uint32_t path = 0;
#if FTN_SHADE_STAGE
        const uint32_t k1 = k + stride, k2 = k + 2u * stride;
        const uint32_t p2 = k2 < n ? queue[k2] : 0u;                     // lands while this path is shaded
        cp_async_wait_all();                                             // this path's state (issued one iteration ago)
        const float4 s_o = stage[buf].ray_o[threadIdx.x], s_d = stage[buf].ray_d[threadIdx.x], s_b = stage[buf].beta[threadIdx.x];
        const uint32_t s_hit = stage[buf].hit[threadIdx.x], s_state = stage[buf].state[threadIdx.x];
        if (k1 < n) stage_path_state(stage[buf ^ 1], pa, p1);           // the next path's, into the other buffer
        cp_async_commit();
#endif
        if (k < n) {
#if FTN_SHADE_STAGE
            path = p0;
            RayF ray; ray.o = V3(s_o.x, s_o.y, s_o.z); ray.d = V3(s_d.x, s_d.y, s_d.z); ray.t_max = s_d.w; ray.time = s_o.w;
            const uint32_t slot = s_hit, state = s_state;
            const V3 beta = V3(s_b.x, s_b.y, s_b.z);
#else
            path = queue[k];
            const RayF ray = load_ray(pa, path);
            const uint32_t slot = pa.hit[path], state = pa.state[path];
            const V3 beta = ld3(pa.beta, path);
#endif
            ShadeOut o;
            RayDiff cd; const RayDiff* carried = nullptr;
            if (IMG && pa.diff[0] && (state & FTN_STATE_HAS_DIFF)) {
                cd.rx_o = ld3(pa.diff[0], path); cd.rx_d = ld3(pa.diff[1], path); cd.ry_o = ld3(pa.diff[2], path); cd.ry_d = ld3(pa.diff[3], path);
                carried = &cd;
            }
            // L enters as 0: the stage only ADDS the hit's own emission (beta * Le), which the sink adds to the path's radiance
            PathDirectSink ds; ds.pa = pa; ds.path = path;
            shade_surface_to<QUEUE == Q_NULL ? -1 : QUEUE - Q_MAT0, IMG>(sc, pp, path, ray, slot, state, beta, v3s(0.0f), &o, &ds, err, carried);
            if (IMG && o.has_diff) {
                if (pa.diff[0]) { st3(pa.diff[0], path, o.diff.rx_o); st3(pa.diff[1], path, o.diff.rx_d); st3(pa.diff[2], path, o.diff.ry_o); st3(pa.diff[3], path, o.diff.ry_d); }
                else o.state &= ~FTN_STATE_HAS_DIFF;
            }
            if (ds.has_shadow) t_shadow = Q_SHADOW;
            if (ds.has_mis) t_mis = Q_MIS;
            if (o.alive) {
                st3(pa.ray_o, path, o.next_o, ray.time);
                st3(pa.ray_d, path, o.next_d, FTN_INF);
                st3(pa.beta, path, o.beta);
                pa.state[path] = o.state;
                t_active = Q_ACTIVE_OUT;
            }
        }
        queue_push3(qs, counts, t_active >= 0, t_shadow >= 0, t_mis >= 0, path);
#if FTN_SHADE_STAGE
        p0 = p1; p1 = p2; buf ^= 1;
#endif
    }
}

// ---- shadow rays: VisibilityTester::unoccluded (light/mod.rs:82-84) ---------------------------------------------
struct ShadowSource {
    PathArrays pa; const uint32_t* queue;
    __device__ __forceinline__ bool load(uint32_t k, RayF* ray) const {
        const uint32_t path = queue[k];
        ray->o = ld3(pa.sh_o, path); ray->d = ld3(pa.sh_d, path);
        ray->t_max = rn_sub(1.0f, 0.0001f);   // 1 - SHADOW_EPSILON, interaction.rs:10,55
        ray->time = 0.0f;                     // carried by the reference's Ray, read by nothing on this path (no motion blur): not worth a sector
        return true;
    }
};
struct ShadowSink {
    PathArrays pa; const uint32_t* queue;
    __device__ __forceinline__ void store(bool valid, uint32_t k, const RayF&, const SceneHit& h) const {
        if (!valid || h.slot != FTN_NO_HIT_SLOT) return;
        const uint32_t path = queue[k];
        st3(pa.L, path, ld3(pa.L, path) + ld3(pa.sh_L, path));
    }
};
template <bool COUNT, bool SPH, int MODE>
__global__ void FTN_TRACE_LAUNCH_BOUNDS
k_shadow(SceneView sc, PathArrays pa, const uint32_t* __restrict__ queue, uint32_t* __restrict__ counts, unsigned long long* __restrict__ trav) {
    ShadowSource src; src.pa = pa; src.queue = queue;
    ShadowSink sink; sink.pa = pa; sink.queue = queue;
    TraceCounters tc; tc.nodes = 0; tc.tris = 0;
    trace_persistent<true, COUNT, SPH, MODE>(sc, counts[Q_SHADOW], &counts[W_SHADOW], src, sink, tc);
    if (COUNT) flush_trace_counters(tc, trav);
}

// ---- MIS (BSDF-sampled) rays: integrator/mod.rs:364-389 ---------------------------------------------------------------
struct MisSource {
    PathArrays pa; const uint32_t* queue;
    __device__ __forceinline__ bool load(uint32_t k, RayF* ray) const {
        const uint32_t path = queue[k];
        ray->o = ld3(pa.mis_o, path); ray->d = ld3(pa.mis_d, path); ray->t_max = FTN_INF; ray->time = 0.0f;
        return true;
    }
};
// FTN_MIS_RESOLVE = 1 (default): the traversal kernel only records the slot the MIS ray found (in PathArrays::hit, which the
// shade stage of this bounce has already consumed) and k_mis_resolve looks the radiance up afterwards with full warps.
// Inlined into the sink (= 0, round 1) the environment lookup (acos / atan2 / four texel fetches) ran on the few lanes that
// had just finished, inside an issue-bound kernel, and its registers cost the persistent loop two resident blocks per SM
// (94-96 registers -> 5 blocks instead of 7; profiles/r02_ab_mis_resolve.txt).
#ifndef FTN_MIS_RESOLVE
#define FTN_MIS_RESOLVE 1
#endif
// FTN_MIS_MIN_BLOCKS (A/B, with FTN_MIS_RESOLVE = 0): cap k_mis's registers for that many resident blocks instead
template <bool ENV_ONLY>
struct MisSink {
    SceneView sc; PathArrays pa; const uint32_t* queue;
    __device__ __forceinline__ void store(bool valid, uint32_t k, const RayF& ray, const SceneHit& h) const {
        if (!valid) return;
        const uint32_t path = queue[k];
#if FTN_MIS_RESOLVE
        // env lights only: a hit contributes nothing, and hit[path] still holds the (non-miss) slot this bounce was shaded at --
        // only the misses are recorded
        if (!ENV_ONLY || h.slot == FTN_NO_HIT_SLOT) pa.hit[path] = h.slot;
#else
        const float4 w4 = pa.mis_w[path];
        const V3 incident = mis_incident(sc, sc.lights[f2u(w4.w)], ray, h.slot);
        if (!is_black(incident)) st3(pa.L, path, ld3(pa.L, path) + V3(w4.x, w4.y, w4.z) * incident);
#endif
    }
};
#if defined(FTN_MIS_MIN_BLOCKS)
#define FTN_MIS_LAUNCH_BOUNDS __launch_bounds__(FTN_TRACE_THREADS, FTN_MIS_MIN_BLOCKS)
#else
#define FTN_MIS_LAUNCH_BOUNDS FTN_TRACE_LAUNCH_BOUNDS
#endif
// ENV_ONLY: with only infinite lights a hit contributes nothing whatever it is, so any-hit suffices
template <bool ENV_ONLY, bool COUNT, bool SPH, int MODE>
__global__ void FTN_MIS_LAUNCH_BOUNDS
k_mis(SceneView sc, PathArrays pa, const uint32_t* __restrict__ queue, uint32_t* __restrict__ counts, unsigned long long* __restrict__ trav) {
    MisSource src; src.pa = pa; src.queue = queue;
    MisSink<ENV_ONLY> sink; sink.sc = sc; sink.pa = pa; sink.queue = queue;
    TraceCounters tc; tc.nodes = 0; tc.tris = 0;
    trace_persistent<ENV_ONLY, COUNT, SPH, MODE>(sc, counts[Q_MIS], &counts[W_MIS], src, sink, tc);
    if (COUNT) flush_trace_counters(tc, trav);
}
// Radiance arriving along the MIS rays of this bounce (integrator/mod.rs:364-389), one thread per ray of the MIS queue
__global__ void __launch_bounds__(256)
k_mis_resolve(SceneView sc, PathArrays pa, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ counts) {
    const uint32_t n = counts[Q_MIS];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t path = queue[k];
        const uint32_t slot = pa.hit[path];
        const float4 w4 = pa.mis_w[path];
        const LightData& light = sc.lights[f2u(w4.w)];
        V3 incident;
        if (light.type == 0) {                         // infinite light: only a miss sees it, and only the direction matters
            if (slot != FTN_NO_HIT_SLOT) continue;
            incident = env_emitted(light.env, ld3(pa.mis_d, path));
        } else {
            RayF ray; ray.o = ld3(pa.mis_o, path); ray.d = ld3(pa.mis_d, path); ray.t_max = FTN_INF; ray.time = 0.0f;
            incident = mis_incident(sc, light, ray, slot);
        }
        if (!is_black(incident)) st3(pa.L, path, ld3(pa.L, path) + V3(w4.x, w4.y, w4.z) * incident);
    }
}

// ---- film ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_film_accumulate(PassParams pp, PathArrays pa, float4* __restrict__ accum, int reach, uint32_t* __restrict__ err) {
    const int fw = pp.film.crop_max[0] - pp.film.crop_min[0], fh = pp.film.crop_max[1] - pp.film.crop_min[1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fw * fh) return;
    float4 acc = accum[i];
    film_gather_pixel(pp, pa.p_film, pa.L, pa.spill, i, reach, &acc, err);
    accum[i] = acc;
}

__global__ void __launch_bounds__(256)
k_film_resolve(const float4* __restrict__ accum, FtnPixel* __restrict__ pixels, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4* out = reinterpret_cast<float4*>(pixels) + i;
    float4 p = *out;
    film_resolve_pixel(accum[i], &p);
    *out = p;
}

// Film::into_spectrum_buffer, film.rs:195-210 + xyz_to_rgb spectrum/mod.rs:28-35
__global__ void __launch_bounds__(256)
k_film_to_rgb(const FtnPixel* __restrict__ pixels, float* __restrict__ rgb, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = reinterpret_cast<const float4*>(pixels)[i];
    float r = rn_sub(rn_sub(rn_mul(3.240479f, p.x), rn_mul(1.537150f, p.y)), rn_mul(0.498535f, p.z));
    float g = rn_add(rn_add(rn_mul(-0.969256f, p.x), rn_mul(1.875991f, p.y)), rn_mul(0.041556f, p.z));
    float b = rn_add(rn_sub(rn_mul(0.055648f, p.x), rn_mul(0.204043f, p.y)), rn_mul(1.057311f, p.z));
    if (p.w != 0.0f) {
        const float inv = rn_div(1.0f, p.w);
        r = fmaxf(0.0f, rn_mul(r, inv)); g = fmaxf(0.0f, rn_mul(g, inv)); b = fmaxf(0.0f, rn_mul(b, inv));
    }
    rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
}

// ---- host driver -------------------------------------------------------------------------------------------------------------
int film_pixel_count(const FtnFilm* f, int32_t* w, int32_t* h) {
    if (!f) return set_error(FTN_ERR_INVALID_ARGUMENT, "null film");
    FilmGeom g;
    if (film_geometry(f, &g) != FTN_OK) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad film (resolution, filter radius or crop window)");
    if (w) *w = g.crop_max[0] - g.crop_min[0];
    if (h) *h = g.crop_max[1] - g.crop_min[1];
    return FTN_OK;
}

// Device workspace of the wavefront state: the process-wide grow-only arena of the device
// (ftn_scene.h), so that neither a second render nor a render of a NEW scene (the end-to-end path:
// upload, build, render, read back) pays cudaMalloc/cudaFree of the path state.  Renders on one
// device serialise on the arena's mutex (they would contend for the SMs anyway).
struct Carver {
    char* p; size_t off = 0;
    template <class T> T* take(size_t count) { off = (off + 255) & ~(size_t)255; T* r = reinterpret_cast<T*>(p + off); off += count * sizeof(T); return r; }
};

// Paths per wavefront pass.  Every launch ends in a tail with idle SMs and every bounce in one host read-back, so
// fewer, larger passes win: C2 (profiles/r01_ab_pass_size.txt) 4Mi -> 7.6 ms, 8Mi -> 6.7 ms, 16Mi -> 6.0 ms per
// 16.8M-path step; C4 at 64 spp (profiles/r02_ab_pass_size.txt) 16Mi -> 4076, 32Mi -> 4252, 64Mi -> 4331 Mrays/s.
// 64Mi paths take 14.8 GB of the 180 GB (220 B per path); a pass is halved until its workspace fits in what the
// device has free (render_device), so a GPU that is shared or holds a 50M-triangle scene still renders.
static size_t g_max_paths_per_pass = 64u << 20;
// bytes of wavefront state for P paths over a film of fw x fh pixels and n_spix sample pixels
static size_t pass_workspace_bytes(size_t P, size_t film_px, size_t n_spix, bool with_diff) {
    return (with_diff ? 14 : 10) * (P * sizeof(float4) + 256) + 2 * (P * 4 + 256) + P * sizeof(float2) + 256 +
           film_px * sizeof(float4) + 256 + (size_t)(Q_COUNT + 1) * (P * 4 + 256) + n_spix + 4096;
}

// CUDA-event pairs around every traversal / shading launch, per kernel class (0 extend, 1 shadow, 2 mis, 3 shade): the
// live per-kernel durations bench.py's roofline uses.  Created only when the caller asks (FTN_STATS_TIME_KERNELS).
struct TraceTimer {
    std::vector<cudaEvent_t> ev; std::vector<int> cls;
    cudaStream_t st; bool on = false;
    void begin(int c) { if (!on) return; cudaEvent_t a, b; if (cudaEventCreate(&a) != cudaSuccess) { on = false; return; } if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); on = false; return; } ev.push_back(a); ev.push_back(b); cls.push_back(c); cudaEventRecord(a, st); }
    void end() { if (on && !ev.empty()) cudaEventRecord(ev.back(), st); }
    void collect(double secs[4], uint64_t launches[4]) {
        for (size_t i = 0; i < cls.size(); ++i) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]) == cudaSuccess) { secs[cls[i]] += ms * 1e-3; launches[cls[i]]++; }
        }
    }
    ~TraceTimer() { for (cudaEvent_t e : ev) cudaEventDestroy(e); }
};
// owns an event for the duration of a scope: early error returns do not leak it
struct EventGuard {
    cudaEvent_t e = nullptr;
    cudaError_t create(unsigned flags = cudaEventDefault) { return cudaEventCreateWithFlags(&e, flags); }
    ~EventGuard() { if (e) cudaEventDestroy(e); }
};

int render_device(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                  const FtnIntegrator* integ, FtnPixel* d_pixels, FtnStats* stats, cudaStream_t st) {
    if (!s || !cam || !film || !smp || !integ || !d_pixels) return set_error(FTN_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->built) return set_error(FTN_ERR_INVALID_ARGUMENT, "ftn_bvh_build has not been called");
    if (smp->mode != FTN_SAMPLER_COUNTER) return set_error(FTN_ERR_UNSUPPORTED, "the GPU renders with the counter sampler; the reference's sequential per-tile stream cannot be parallelised per sample");
    if (smp->samples_per_pixel < 1 || smp->sample_stride < 1 || smp->sample_begin < 0) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad sampler");
    if (integ->type != FTN_INTEGRATOR_PATH && integ->type != FTN_INTEGRATOR_DIRECT_LIGHTING) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad integrator");
    if (integ->max_depth < 0 || integ->max_depth > 60000) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad max_depth");
    const uint32_t stat_flags = stats ? stats->flags : 0u;   // the one INPUT field of FtnStats
    const bool count_traversal = (stat_flags & FTN_STATS_COUNT_TRAVERSAL) != 0u;
    FTN_CUDA(cudaSetDevice(s->device));
    FilmGeom fg;
    if (film_geometry(film, &fg) != FTN_OK) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad film (resolution, filter radius or crop window)");
    const int fw = fg.crop_max[0] - fg.crop_min[0], fh = fg.crop_max[1] - fg.crop_min[1];
    const int sbw = fg.sb_max[0] - fg.sb_min[0], sbh = fg.sb_max[1] - fg.sb_min[1];
    const size_t n_spix = (size_t)sbw * sbh;
    // samples owned by this call: s = begin + i*stride < spp
    const int n_samples = (smp->sample_begin < smp->samples_per_pixel) ? (smp->samples_per_pixel - smp->sample_begin + smp->sample_stride - 1) / smp->sample_stride : 0;
    const char* env_pp = getenv("FTN_PATHS_PER_PASS");
    size_t max_paths = env_pp ? (size_t)strtoull(env_pp, nullptr, 10) : g_max_paths_per_pass;
    int s_per_pass = (int)std::max<size_t>(1, max_paths / std::max<size_t>(1, n_spix));
    s_per_pass = std::min(s_per_pass, std::max(1, n_samples));
    if (s->device < 0 || s->device >= FTN_MAX_DEVICES) return set_error(FTN_ERR_INVALID_ARGUMENT, "device index out of range");
    DeviceArena& arena = device_arena(s->device);
    std::lock_guard<std::recursive_mutex> arena_lock(arena.m);
    // ray differentials behind mirrors (integrator/mod.rs:59-83) only matter where a texture is filtered with them
    const bool with_diff = integ->type == FTN_INTEGRATOR_DIRECT_LIGHTING && s->has_image_texture;
    while (s_per_pass > 1 && n_spix * (size_t)s_per_pass >= (1ull << 31)) s_per_pass = (s_per_pass + 1) / 2;
    if (pass_workspace_bytes(n_spix * (size_t)s_per_pass, (size_t)fw * fh, n_spix, with_diff) > arena.bytes[DeviceArena::PATHS]) {
        // the arena has to grow: shrink the pass until its workspace fits in what the arena already holds + 3/4 of the free
        // memory (cudaMemGetInfo costs 0.7-6 ms on a process that holds GBs, so it is asked only here, not per render)
        size_t free_b = 0, total_b = 0;
        FTN_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = arena.bytes[DeviceArena::PATHS] + free_b / 4 * 3;
        while (s_per_pass > 1 && pass_workspace_bytes(n_spix * (size_t)s_per_pass, (size_t)fw * fh, n_spix, with_diff) > budget)
            s_per_pass = (s_per_pass + 1) / 2;
    }
    const size_t P = n_spix * (size_t)s_per_pass;
    if (P >= (1ull << 31)) return set_error(FTN_ERR_INVALID_ARGUMENT, "film too large for one pass");

    EventGuard g0, g1, gc;
    FTN_CUDA(g0.create()); FTN_CUDA(g1.create()); FTN_CUDA(gc.create(cudaEventDisableTiming));
    const cudaEvent_t ev0 = g0.e, ev1 = g1.e, ev_counts = gc.e;
    FTN_CUDA(cudaEventRecord(ev0, st));

    const size_t ws_bytes = pass_workspace_bytes(P, (size_t)fw * fh, n_spix, with_diff);
    Carver cv;
    FTN_TRY(arena.reserve(DeviceArena::PATHS, ws_bytes, "cudaMalloc (render workspace)", (void**)&cv.p));
    PathArrays pa;
    pa.ray_o = cv.take<float4>(P); pa.ray_d = cv.take<float4>(P); pa.beta = cv.take<float4>(P); pa.L = cv.take<float4>(P);
    pa.sh_o = cv.take<float4>(P); pa.sh_d = cv.take<float4>(P); pa.sh_L = cv.take<float4>(P);
    pa.mis_o = cv.take<float4>(P); pa.mis_d = cv.take<float4>(P); pa.mis_w = cv.take<float4>(P);
    pa.hit = cv.take<uint32_t>(P); pa.state = cv.take<uint32_t>(P); pa.p_film = cv.take<float2>(P);
    for (int i = 0; i < 4; ++i) pa.diff[i] = with_diff ? cv.take<float4>(P) : nullptr;
    pa.spill = cv.take<uint8_t>(n_spix);
    float4* accum = cv.take<float4>((size_t)fw * fh);
    Queues qs;
    for (int q = 0; q < Q_COUNT; ++q) qs.q[q] = cv.take<uint32_t>(P);
    uint32_t* active_in = cv.take<uint32_t>(P);   // ping-pong partner of Q_ACTIVE_OUT
    uint32_t* counts = cv.take<uint32_t>(CTR_COUNT);
    uint32_t* d_err = cv.take<uint32_t>(1);
    unsigned long long* d_trav = cv.take<unsigned long long>(6);   // [class][nodes, tris]
    FTN_CUDA(cudaMemsetAsync(accum, 0, (size_t)fw * fh * sizeof(float4), st));
    FTN_CUDA(cudaMemsetAsync(d_err, 0, 4, st));
    FTN_CUDA(cudaMemsetAsync(d_trav, 0, 6 * sizeof(unsigned long long), st));

    const SceneView sc = s->view();
    const bool sph = s->n_spheres > 0;
    bool has_area = false;
    for (const LightData& l : s->h_lights) if (l.type == 1 || l.type == FTN_LIGHT_TYPE_TRIANGLE) has_area = true;
    uint64_t camera_samples = 0;
    uint64_t class_rays[3] = {0, 0, 0};
    // shade kernels: grid-stride loops over the class queues; FTN_SHADE_BLOCKS_PER_SM (A/B knob, default 8 = two waves of
    // the 4 resident blocks)
    const char* env_sb = getenv("FTN_SHADE_BLOCKS_PER_SM");
    const unsigned shade_grid = (unsigned)(sm_count() * std::max(1, std::min(32, env_sb ? atoi(env_sb) : 8)));
    const int reach = (int)std::ceil(std::max(fg.radius[0], fg.radius[1]) + 0.5f);
    TraceTimer timer; timer.st = st; timer.on = (stat_flags & FTN_STATS_TIME_KERNELS) != 0u;
    uint32_t* h_counts = nullptr;           // pinned read-back slot of the queue counters, one per device
    FTN_TRY(arena.pinned_counts(&h_counts, CTR_COUNT * sizeof(uint32_t)));

    PassParams pp; std::memset(&pp, 0, sizeof(pp));
    pp.film = fg; pp.cam = *cam; pp.seed_key = sampler_seed_key(smp->seed);
    pp.spp = smp->samples_per_pixel; pp.s_stride = smp->sample_stride;
    pp.integrator = integ->type; pp.max_depth = integ->max_depth; pp.rr_threshold = integ->rr_threshold;

    const int tmode = trace_mode(sc);
    for (int done = 0; done < n_samples; done += s_per_pass) {
        const int sc_n = std::min(s_per_pass, n_samples - done);
        pp.s_first = smp->sample_begin + done * smp->sample_stride;
        pp.s_count = sc_n;
        pp.n_paths = (uint32_t)(n_spix * (size_t)sc_n);
        FTN_CUDA(cudaMemsetAsync(pa.spill, 0, n_spix, st));
        k_raygen<<<(pp.n_paths + 255) / 256, 256, 0, st>>>(pp, pa);
        FTN_LAUNCHED();
        camera_samples += pp.n_paths;
        uint32_t n_active = pp.n_paths;
        const uint32_t* q_in = nullptr;   // identity queue for the first extend
        uint32_t* q_out = qs.q[Q_ACTIVE_OUT];
        uint32_t* q_spare = active_in;
        const int iter_cap = integ->max_depth + 2 + 4096;   // null-BSDF surfaces do not count as bounces
        for (int it = 0; it < iter_cap && n_active > 0; ++it) {
            FTN_CUDA(cudaMemsetAsync(counts, 0, CTR_COUNT * 4, st));
            Queues q = qs; q.q[Q_ACTIVE_OUT] = q_out;
            timer.begin(0);
            FTN_MODE3(tmode, FTN_BOOL2(count_traversal, sph, (k_extend<B0, B1, M><<<persistent_grid(k_extend<B0, B1, M>, n_active), FTN_TRACE_THREADS, 0, st>>>(sc, pa, q_in, n_active, q, counts, d_trav, integ->type == FTN_INTEGRATOR_DIRECT_LIGHTING))));
            timer.end();
            FTN_LAUNCHED();
            class_rays[0] += n_active;
            timer.begin(3);
            k_shade_miss<<<shade_grid, 256, 0, st>>>(sc, pp, pa, q.q[Q_MISS], counts);
            FTN_LAUNCHED();
            if (s->has_null_material) { k_shade<Q_NULL><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_NULL], q, counts, d_err); FTN_LAUNCHED(); }
            if (s->material_present[0]) {
                if (s->has_image_texture) k_shade<Q_MAT0, true><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT0], q, counts, d_err);
                else k_shade<Q_MAT0><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT0], q, counts, d_err);
                FTN_LAUNCHED();
            }
            if (s->material_present[1]) {
                if (s->has_image_texture) k_shade<Q_MAT1, true><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT1], q, counts, d_err);
                else k_shade<Q_MAT1><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT1], q, counts, d_err);
                FTN_LAUNCHED();
            }
            if (s->material_present[2]) {
                if (s->has_image_texture) k_shade<Q_MAT2, true><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT2], q, counts, d_err);
                else k_shade<Q_MAT2><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT2], q, counts, d_err);
                FTN_LAUNCHED();
            }
            if (s->material_present[3]) {
                if (s->has_image_texture) k_shade<Q_MAT3, true><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT3], q, counts, d_err);
                else k_shade<Q_MAT3><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT3], q, counts, d_err);
                FTN_LAUNCHED();
            }
            if (s->material_present[4]) {   // rough glass
                if (s->has_image_texture) k_shade<Q_MAT4, true><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT4], q, counts, d_err);
                else k_shade<Q_MAT4><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT4], q, counts, d_err);
                FTN_LAUNCHED();
            }
            if (s->material_present[5]) {
                if (s->has_image_texture) k_shade<Q_MAT5, true><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT5], q, counts, d_err);
                else k_shade<Q_MAT5><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT5], q, counts, d_err);
                FTN_LAUNCHED();
            }
            timer.end();
            // The shadow and MIS queues cannot be longer than this iteration's input queue, and their kernels
            // read the exact lengths on the device: launch them sized by that bound BEFORE waiting for the
            // counts, so that the host round trip (needed to size the next iteration and to stop) is hidden
            // behind them instead of leaving the GPU idle once per bounce.
            // (the copy goes to PINNED memory and is waited for through an event recorded right behind it: the
            //  shadow / MIS kernels are already queued when the host wakes up)
            FTN_CUDA(cudaMemcpyAsync(h_counts, counts, CTR_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            FTN_CUDA(cudaEventRecord(ev_counts, st));
            timer.begin(1);
            FTN_MODE3(tmode, FTN_BOOL2(count_traversal, sph, (k_shadow<B0, B1, M><<<persistent_grid(k_shadow<B0, B1, M>, n_active), FTN_TRACE_THREADS, 0, st>>>(sc, pa, q.q[Q_SHADOW], counts, d_trav + 2))));
            timer.end();
            FTN_LAUNCHED();
            timer.begin(2);
            if (has_area) { FTN_MODE3(tmode, FTN_BOOL2(count_traversal, sph, (k_mis<false, B0, B1, M><<<persistent_grid(k_mis<false, B0, B1, M>, n_active), FTN_TRACE_THREADS, 0, st>>>(sc, pa, q.q[Q_MIS], counts, d_trav + 4)))); }
            else { FTN_MODE3(tmode, FTN_BOOL2(count_traversal, sph, (k_mis<true, B0, B1, M><<<persistent_grid(k_mis<true, B0, B1, M>, n_active), FTN_TRACE_THREADS, 0, st>>>(sc, pa, q.q[Q_MIS], counts, d_trav + 4)))); }
            FTN_LAUNCHED();
#if FTN_MIS_RESOLVE
            k_mis_resolve<<<shade_grid, 256, 0, st>>>(sc, pa, q.q[Q_MIS], counts);
            FTN_LAUNCHED();
#endif
            timer.end();
            FTN_CUDA(cudaEventSynchronize(ev_counts));
            const uint32_t* hc = h_counts;
            class_rays[1] += hc[Q_SHADOW];
            class_rays[2] += hc[Q_MIS];
            n_active = hc[Q_ACTIVE_OUT];
            q_in = q_out;
            std::swap(q_out, q_spare);
        }
        k_film_accumulate<<<(fw * fh + 127) / 128, 128, 0, st>>>(pp, pa, accum, reach, d_err);
        FTN_LAUNCHED();
    }
    k_film_resolve<<<(fw * fh + 255) / 256, 256, 0, st>>>(accum, d_pixels, fw * fh);
    FTN_LAUNCHED();
    uint32_t h_err = 0;
    unsigned long long h_trav[6] = {0, 0, 0, 0, 0, 0};
    FTN_CUDA(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
    FTN_CUDA(cudaMemcpyAsync(h_trav, d_trav, sizeof(h_trav), cudaMemcpyDeviceToHost, st));
    FTN_CUDA(cudaEventRecord(ev1, st));
    FTN_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.0f;
    FTN_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->camera_samples = camera_samples;
        stats->rays_closest = class_rays[0] + class_rays[2]; stats->rays_any = class_rays[1];   // Scene::intersect / intersect_test calls
        stats->device_seconds = ms * 1e-3; stats->bvh_build_seconds = s->build_seconds; stats->morton_sort_seconds = s->sort_seconds;
        stats->bvh_nodes = s->n_nodes; stats->bvh_node_bytes = s->wide ? FTN_NODE8_BYTES : FTN_NODE_BYTES; stats->bvh_tri_bytes = FTN_TRI_BYTES;
        double tsec[4] = {0, 0, 0, 0}; uint64_t tl[4] = {0, 0, 0, 0};
        timer.collect(tsec, tl);
        for (int c = 0; c < 3; ++c) { stats->trace_seconds[c] = tsec[c]; stats->trace_launches[c] = tl[c]; }
        stats->shade_seconds = tsec[3]; stats->shade_launches = tl[3];
        stats->flags = stat_flags;
        for (int c = 0; c < 3; ++c) { stats->trace_rays[c] = class_rays[c]; stats->trace_nodes[c] = h_trav[2 * c]; stats->trace_tris[c] = h_trav[2 * c + 1]; }
        stats->node_visits = h_trav[0] + h_trav[2] + h_trav[4]; stats->tri_tests = h_trav[1] + h_trav[3] + h_trav[5];
    }
    if (h_err & ERR_NAN) return set_error(FTN_ERR_NAN_RADIANCE, "NaN radiance value (check_radiance, integrator/mod.rs:285)");
    if (h_err & ERR_UNSUPPORTED) return set_error(FTN_ERR_UNSUPPORTED, "the reference hits unimplemented!() on this input (env map_pdf == 0 or a null BSDF under the direct-lighting integrator)");
    return FTN_OK;
}

// ftn_render: device film from the arena (no per-call cudaMalloc/cudaFree), then one D2H copy.
int render_host(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                const FtnIntegrator* integ, FtnPixel* out_pixels, FtnStats* stats) {
    if (!s || !film || !out_pixels) return set_error(FTN_ERR_INVALID_ARGUMENT, "null argument");
    int32_t w = 0, h = 0;
    FTN_TRY(film_pixel_count(film, &w, &h));
    if (s->device < 0 || s->device >= FTN_MAX_DEVICES) return set_error(FTN_ERR_INVALID_ARGUMENT, "device index out of range");
    FTN_CUDA(cudaSetDevice(s->device));
    const size_t bytes = (size_t)w * h * sizeof(FtnPixel);
    DeviceArena& arena = device_arena(s->device);
    FtnPixel* d_px = nullptr;
    // held from the reservation of the film slot to the end of the read-back: a concurrent ftn_render with a larger
    // film (or ftn_release_cached_memory) would otherwise re-allocate the slot under this render
    std::lock_guard<std::recursive_mutex> lock(arena.m);
    FTN_TRY(arena.reserve(DeviceArena::FILM, bytes, "cudaMalloc (film)", (void**)&d_px));
    FTN_CUDA(cudaMemsetAsync(d_px, 0, bytes, nullptr));
    const int render_rc = render_device(s, cam, film, smp, integ, d_px, stats, nullptr);
    // like the reference's panic, a NaN / unsupported render still leaves the film readable
    if (render_rc == FTN_OK || render_rc == FTN_ERR_NAN_RADIANCE || render_rc == FTN_ERR_UNSUPPORTED) {
        const std::string keep = last_error_string();
        cudaError_t e = cudaMemcpy(out_pixels, d_px, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return cuda_fail(e, "D2H film", __FILE__, __LINE__);
        restore_error_string(keep);
    }
    return render_rc;
}

int film_to_rgb_device(size_t n, const FtnPixel* d_pixels, float* d_rgb, cudaStream_t st) {
    if (n == 0) return FTN_OK;
    k_film_to_rgb<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_pixels, d_rgb, n);
    FTN_LAUNCHED();
    return FTN_OK;
}

}  // namespace ftn
