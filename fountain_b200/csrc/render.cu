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
#include <algorithm>
#include <cstring>
#include <cmath>

namespace ftn {

int sm_count();
unsigned trace_grid(size_t n, int blocks_per_sm);

enum { Q_ACTIVE_OUT = 0, Q_MISS, Q_NULL, Q_MAT0, Q_MAT1, Q_MAT2, Q_SHADOW, Q_MIS, Q_COUNT };
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

// ---- raygen ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_raygen(PassParams pp, PathArrays pa) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pp.n_paths) return;
    float fx, fy;
    const RayF ray = raygen_path(pp, i, &fx, &fy);
    st3(pa.ray_o, i, ray.o, ray.time);
    st3(pa.ray_d, i, ray.d, ray.t_max);
    st3(pa.beta, i, v3s(1.0f));
    st3(pa.L, i, v3s(0.0f));
    pa.state[i] = 0u;
    pa.p_film[i] = make_float2(fx, fy);
}

// ---- extend: closest hit + binning by material class ----------------------------------------------------------
// queue_in == nullptr: the identity queue (first bounce of a pass)
__global__ void __launch_bounds__(FTN_TRACE_THREADS)
k_extend(SceneView sc, PathArrays pa, const uint32_t* __restrict__ queue_in, uint32_t n_in, Queues qs, uint32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&counts[W_EXTEND], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_in) break;
        const uint32_t k = base + lane;
        int target = -1;
        uint32_t path = 0;
        if (k < n_in) {
            path = queue_in ? queue_in[k] : k;
            const RayF ray = load_ray(pa, path);
            SceneHit h; TraceCounters tc;
            scene_intersect<false, false>(sc, ray, &h, &tc);
            pa.hit[path] = h.slot;
            if (h.slot == FTN_NO_HIT_SLOT) target = Q_MISS;
            else {
                const int material = hit_material(sc, h.slot);
                target = (material < 0) ? Q_NULL : (Q_MAT0 + sc.materials[material].type);
            }
        }
        queue_push(qs.q, counts, target, path);
    }
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
template <int QUEUE>
__global__ void __launch_bounds__(128)
k_shade(SceneView sc, PassParams pp, PathArrays pa, const uint32_t* __restrict__ queue, Queues qs, uint32_t* __restrict__ counts, uint32_t* __restrict__ err) {
    const uint32_t n = counts[QUEUE];
    const uint32_t n32 = (n + 31u) & ~31u;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n32; k += gridDim.x * blockDim.x) {
        int t_active = -1, t_shadow = -1, t_mis = -1;
        uint32_t path = 0;
        if (k < n) {
            path = queue[k];
            const RayF ray = load_ray(pa, path);
            ShadeOut o;
            shade_surface(sc, pp, path, ray, pa.hit[path], pa.state[path], ld3(pa.beta, path), ld3(pa.L, path), &o, err);
            st3(pa.L, path, o.L);
            if (o.direct.has_shadow) { st3(pa.sh_o, path, o.direct.sh_o); st3(pa.sh_d, path, o.direct.sh_d); st3(pa.sh_L, path, o.direct.sh_L); t_shadow = Q_SHADOW; }
            if (o.direct.has_mis) {
                st3(pa.mis_o, path, o.direct.mis_o); st3(pa.mis_d, path, o.direct.mis_d);
                st3(pa.mis_w, path, o.direct.mis_w, u2f((uint32_t)o.direct.mis_light)); t_mis = Q_MIS;
            }
            if (o.alive) {
                st3(pa.ray_o, path, o.next_o, ray.time);
                st3(pa.ray_d, path, o.next_d, FTN_INF);
                st3(pa.beta, path, o.beta);
                pa.state[path] = o.state;
                t_active = Q_ACTIVE_OUT;
            }
        }
        queue_push(qs.q, counts, t_active, path);
        queue_push(qs.q, counts, t_shadow, path);
        queue_push(qs.q, counts, t_mis, path);
    }
}

// ---- shadow rays: VisibilityTester::unoccluded (light/mod.rs:82-84) ---------------------------------------------
__global__ void __launch_bounds__(FTN_TRACE_THREADS)
k_shadow(SceneView sc, PathArrays pa, const uint32_t* __restrict__ queue, uint32_t* __restrict__ counts) {
    const uint32_t n = counts[Q_SHADOW];
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&counts[W_SHADOW], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t k = base + lane;
        if (k < n) {
            const uint32_t path = queue[k];
            RayF ray; ray.o = ld3(pa.sh_o, path); ray.d = ld3(pa.sh_d, path);
            ray.t_max = rn_sub(1.0f, 0.0001f);   // 1 - SHADOW_EPSILON, interaction.rs:10,55
            ray.time = pa.ray_o[path].w;
            SceneHit h; TraceCounters tc;
            scene_intersect<true, false>(sc, ray, &h, &tc);
            if (h.slot == FTN_NO_HIT_SLOT) st3(pa.L, path, ld3(pa.L, path) + ld3(pa.sh_L, path));
        }
    }
}

// ---- MIS (BSDF-sampled) rays: integrator/mod.rs:364-389 ---------------------------------------------------------------
template <bool ENV_ONLY>
__global__ void __launch_bounds__(FTN_TRACE_THREADS)
k_mis(SceneView sc, PathArrays pa, const uint32_t* __restrict__ queue, uint32_t* __restrict__ counts) {
    const uint32_t n = counts[Q_MIS];
    const int lane = threadIdx.x & 31;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&counts[W_MIS], 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t k = base + lane;
        if (k < n) {
            const uint32_t path = queue[k];
            const float4 w4 = pa.mis_w[path];
            const LightData& light = sc.lights[f2u(w4.w)];
            RayF ray; ray.o = ld3(pa.mis_o, path); ray.d = ld3(pa.mis_d, path); ray.t_max = FTN_INF; ray.time = pa.ray_o[path].w;
            SceneHit h; TraceCounters tc;
            // with only infinite lights a hit contributes nothing whatever it is, so any-hit suffices
            if (ENV_ONLY) scene_intersect<true, false>(sc, ray, &h, &tc);
            else scene_intersect<false, false>(sc, ray, &h, &tc);
            const V3 incident = mis_incident(sc, light, ray, h.slot);
            if (!is_black(incident)) st3(pa.L, path, ld3(pa.L, path) + V3(w4.x, w4.y, w4.z) * incident);
        }
    }
}

// ---- film ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_film_accumulate(PassParams pp, PathArrays pa, float4* __restrict__ accum, int reach, uint32_t* __restrict__ err) {
    const int fw = pp.film.crop_max[0] - pp.film.crop_min[0], fh = pp.film.crop_max[1] - pp.film.crop_min[1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= fw * fh) return;
    float4 acc = accum[i];
    film_gather_pixel(pp, pa.p_film, pa.L, i, reach, &acc, err);
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

struct DeviceBuf {
    void* p = nullptr;
    ~DeviceBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc (render)", __FILE__, __LINE__); }
        return FTN_OK;
    }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

static size_t g_max_paths_per_pass = 4u << 20;

int render_device(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                  const FtnIntegrator* integ, FtnPixel* d_pixels, FtnStats* stats, cudaStream_t st) {
    if (!s || !cam || !film || !smp || !integ || !d_pixels) return set_error(FTN_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->built) return set_error(FTN_ERR_INVALID_ARGUMENT, "ftn_bvh_build has not been called");
    if (smp->mode != FTN_SAMPLER_COUNTER) return set_error(FTN_ERR_UNSUPPORTED, "the GPU renders with the counter sampler; the reference's sequential per-tile stream cannot be parallelised per sample");
    if (smp->samples_per_pixel < 1 || smp->sample_stride < 1 || smp->sample_begin < 0) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad sampler");
    if (integ->type != FTN_INTEGRATOR_PATH && integ->type != FTN_INTEGRATOR_DIRECT_LIGHTING) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad integrator");
    if (integ->max_depth < 0 || integ->max_depth > 60000) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad max_depth");
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
    const size_t P = n_spix * (size_t)s_per_pass;
    if (P >= (1ull << 31)) return set_error(FTN_ERR_INVALID_ARGUMENT, "film too large for one pass");

    cudaEvent_t ev0, ev1;
    FTN_CUDA(cudaEventCreate(&ev0)); FTN_CUDA(cudaEventCreate(&ev1));
    FTN_CUDA(cudaEventRecord(ev0, st));

    DeviceBuf b_f4[10], b_hit, b_state, b_pfilm, b_accum, b_queues, b_counts, b_err;
    for (int i = 0; i < 10; ++i) FTN_TRY(b_f4[i].alloc(P * sizeof(float4)));
    FTN_TRY(b_hit.alloc(P * 4)); FTN_TRY(b_state.alloc(P * 4)); FTN_TRY(b_pfilm.alloc(P * sizeof(float2)));
    FTN_TRY(b_accum.alloc((size_t)fw * fh * sizeof(float4)));
    FTN_TRY(b_queues.alloc((size_t)(Q_COUNT + 1) * P * 4));
    FTN_TRY(b_counts.alloc(CTR_COUNT * 4)); FTN_TRY(b_err.alloc(4));
    PathArrays pa;
    pa.ray_o = b_f4[0].as<float4>(); pa.ray_d = b_f4[1].as<float4>(); pa.beta = b_f4[2].as<float4>(); pa.L = b_f4[3].as<float4>();
    pa.sh_o = b_f4[4].as<float4>(); pa.sh_d = b_f4[5].as<float4>(); pa.sh_L = b_f4[6].as<float4>();
    pa.mis_o = b_f4[7].as<float4>(); pa.mis_d = b_f4[8].as<float4>(); pa.mis_w = b_f4[9].as<float4>();
    pa.hit = b_hit.as<uint32_t>(); pa.state = b_state.as<uint32_t>(); pa.p_film = b_pfilm.as<float2>();
    float4* accum = b_accum.as<float4>();
    uint32_t* counts = b_counts.as<uint32_t>();
    uint32_t* d_err = b_err.as<uint32_t>();
    uint32_t* qmem = b_queues.as<uint32_t>();
    uint32_t* active_in = qmem + (size_t)Q_COUNT * P;   // ping-pong partner of Q_ACTIVE_OUT
    Queues qs;
    for (int q = 0; q < Q_COUNT; ++q) qs.q[q] = qmem + (size_t)q * P;
    FTN_CUDA(cudaMemsetAsync(accum, 0, (size_t)fw * fh * sizeof(float4), st));
    FTN_CUDA(cudaMemsetAsync(d_err, 0, 4, st));

    const SceneView sc = s->view();
    bool has_area = false, mat_present[3] = {false, false, false};
    for (const LightData& l : s->h_lights) if (l.type == 1) has_area = true;
    {   // which material classes exist decides which shade kernels are launched at all
        std::vector<MaterialData> mats(s->n_materials);
        if (s->n_materials) FTN_CUDA(cudaMemcpy(mats.data(), s->d_materials, mats.size() * sizeof(MaterialData), cudaMemcpyDeviceToHost));
        for (const MaterialData& m : mats) mat_present[m.type] = true;
    }
    bool has_null = false;   // any primitive without a material (null BSDF, path.rs:76-80)
    {
        std::vector<MeshData> meshes(s->n_meshes);
        if (s->n_meshes) FTN_CUDA(cudaMemcpy(meshes.data(), s->d_meshes, meshes.size() * sizeof(MeshData), cudaMemcpyDeviceToHost));
        for (const MeshData& m : meshes) if (m.material < 0) has_null = true;
        for (const SphereData& sd : s->h_spheres) if (sd.material < 0) has_null = true;
    }
    uint64_t rays_closest = 0, rays_any = 0, camera_samples = 0;
    const unsigned shade_grid = (unsigned)(sm_count() * 8);
    const int reach = (int)std::ceil(std::max(fg.radius[0], fg.radius[1]) + 0.5f);

    PassParams pp; std::memset(&pp, 0, sizeof(pp));
    pp.film = fg; pp.cam = *cam; pp.seed_key = sampler_seed_key(smp->seed);
    pp.spp = smp->samples_per_pixel; pp.s_stride = smp->sample_stride;
    pp.integrator = integ->type; pp.max_depth = integ->max_depth; pp.rr_threshold = integ->rr_threshold;

    for (int done = 0; done < n_samples; done += s_per_pass) {
        const int sc_n = std::min(s_per_pass, n_samples - done);
        pp.s_first = smp->sample_begin + done * smp->sample_stride;
        pp.s_count = sc_n;
        pp.n_paths = (uint32_t)(n_spix * (size_t)sc_n);
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
            k_extend<<<trace_grid(n_active, FTN_TRACE_BLOCKS_PER_SM), FTN_TRACE_THREADS, 0, st>>>(sc, pa, q_in, n_active, q, counts);
            FTN_LAUNCHED();
            rays_closest += n_active;
            k_shade_miss<<<shade_grid, 256, 0, st>>>(sc, pp, pa, q.q[Q_MISS], counts);
            FTN_LAUNCHED();
            if (has_null) { k_shade<Q_NULL><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_NULL], q, counts, d_err); FTN_LAUNCHED(); }
            if (mat_present[0]) { k_shade<Q_MAT0><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT0], q, counts, d_err); FTN_LAUNCHED(); }
            if (mat_present[1]) { k_shade<Q_MAT1><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT1], q, counts, d_err); FTN_LAUNCHED(); }
            if (mat_present[2]) { k_shade<Q_MAT2><<<shade_grid, 128, 0, st>>>(sc, pp, pa, q.q[Q_MAT2], q, counts, d_err); FTN_LAUNCHED(); }
            uint32_t hc[CTR_COUNT];
            FTN_CUDA(cudaMemcpyAsync(hc, counts, sizeof(hc), cudaMemcpyDeviceToHost, st));
            FTN_CUDA(cudaStreamSynchronize(st));
            if (hc[Q_SHADOW]) {
                k_shadow<<<trace_grid(hc[Q_SHADOW], FTN_TRACE_BLOCKS_PER_SM), FTN_TRACE_THREADS, 0, st>>>(sc, pa, q.q[Q_SHADOW], counts);
                FTN_LAUNCHED();
                rays_any += hc[Q_SHADOW];
            }
            if (hc[Q_MIS]) {
                const unsigned g = trace_grid(hc[Q_MIS], FTN_TRACE_BLOCKS_PER_SM);
                if (has_area) k_mis<false><<<g, FTN_TRACE_THREADS, 0, st>>>(sc, pa, q.q[Q_MIS], counts);
                else k_mis<true><<<g, FTN_TRACE_THREADS, 0, st>>>(sc, pa, q.q[Q_MIS], counts);
                FTN_LAUNCHED();
                rays_closest += hc[Q_MIS];
            }
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
    FTN_CUDA(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
    FTN_CUDA(cudaEventRecord(ev1, st));
    FTN_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.0f;
    FTN_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->camera_samples = camera_samples; stats->rays_closest = rays_closest; stats->rays_any = rays_any;
        stats->device_seconds = ms * 1e-3; stats->bvh_build_seconds = s->build_seconds;
        stats->bvh_nodes = s->n_nodes; stats->bvh_node_bytes = 64; stats->bvh_tri_bytes = 48;
    }
    if (h_err & ERR_NAN) return set_error(FTN_ERR_NAN_RADIANCE, "NaN radiance value (check_radiance, integrator/mod.rs:285)");
    if (h_err & ERR_UNSUPPORTED) return set_error(FTN_ERR_UNSUPPORTED, "the reference hits unimplemented!() on this input (env map_pdf == 0 or a null BSDF under the direct-lighting integrator)");
    return FTN_OK;
}

int film_to_rgb_device(size_t n, const FtnPixel* d_pixels, float* d_rgb, cudaStream_t st) {
    if (n == 0) return FTN_OK;
    k_film_to_rgb<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_pixels, d_rgb, n);
    FTN_LAUNCHED();
    return FTN_OK;
}

}  // namespace ftn
