// extern "C" surface of libfountain_gpu.so (include/fountain_gpu.h).  No CPU fallback: every
// compute entry point needs a CUDA device and says so when there is none.
#include "ftn_scene.h"
#include <mutex>
#include <atomic>
#include <cstdio>
#include <cstring>

namespace ftn {
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const std::string& msg) { g_last_error = msg; return code; }
std::string last_error_string() { return g_last_error; }
void restore_error_string(const std::string& s) { g_last_error = s; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    cudaGetLastError();   // clear the sticky flag of non-fatal errors
    const int code = (e == cudaErrorMemoryAllocation) ? FTN_ERR_OUT_OF_MEMORY
                   : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? FTN_ERR_NO_DEVICE : FTN_ERR_CUDA;
    return set_error(code, buf);
}
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int intersect_device(const FtnScene* s, size_t n, const FtnRay* d_rays, FtnHit* d_hits, uint8_t* d_any,
                     bool any, unsigned long long* d_counters, cudaStream_t st);
int render_device(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                  const FtnIntegrator* integ, FtnPixel* d_pixels, FtnStats* stats, cudaStream_t st);
int film_to_rgb_device(size_t n, const FtnPixel* d_pixels, float* d_rgb, cudaStream_t st);
int film_pixel_count(const FtnFilm* f, int32_t* w, int32_t* h);
int render_host(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                const FtnIntegrator* integ, FtnPixel* out_pixels, FtnStats* stats);
int release_cached_memory();
int render_multi(FtnScene* const* scenes, int32_t n, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                 const FtnIntegrator* integ, FtnPixel* out_pixels, FtnStats* stats);
}  // namespace ftn

using namespace ftn;

extern "C" {

FTN_API uint32_t ftn_abi_version(void) { return FTN_ABI_VERSION; }
FTN_API const char* ftn_last_error(void) { return g_last_error.c_str(); }
FTN_API uint64_t ftn_kernel_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

FTN_API int ftn_device_count(int* out_count) {
    if (!out_count) return set_error(FTN_ERR_INVALID_ARGUMENT, "null argument");
    *out_count = 0;
    cudaError_t e = cudaGetDeviceCount(out_count);
    if (e != cudaSuccess) { *out_count = 0; return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__); }
    return FTN_OK;
}
FTN_API int ftn_set_device(int device) { FTN_CUDA(cudaSetDevice(device)); return FTN_OK; }

FTN_API int ftn_scene_create(const FtnSceneDesc* desc, FtnScene** out_scene) {
    int n = 0;
    FTN_TRY(ftn_device_count(&n));
    if (n < 1) return set_error(FTN_ERR_NO_DEVICE, "no CUDA device; this library has no CPU path");
    return scene_create(desc, out_scene);
}
FTN_API int ftn_scene_destroy(FtnScene* scene) { return scene_destroy(scene); }
FTN_API int ftn_bvh_build(FtnScene* scene) { return bvh_build(scene); }

FTN_API int ftn_bvh_debug_morton(const FtnScene* s, uint32_t* codes, uint32_t* order) {
    if (!s || !s->built) return set_error(FTN_ERR_INVALID_ARGUMENT, "scene not built");
    if (s->n_tris == 0) return FTN_OK;
    if (codes) FTN_CUDA(cudaMemcpy(codes, s->d_codes, (size_t)s->n_tris * 4, cudaMemcpyDeviceToHost));
    if (order) FTN_CUDA(cudaMemcpy(order, s->d_order, (size_t)s->n_tris * 4, cudaMemcpyDeviceToHost));
    return FTN_OK;
}
FTN_API int ftn_scene_world_bound(const FtnScene* s, float out[6]) {
    if (!s || !s->built || !out) return set_error(FTN_ERR_INVALID_ARGUMENT, "scene not built");
    std::memcpy(out, s->bounds, sizeof(float) * 6);
    return FTN_OK;
}
FTN_API int ftn_scene_stats(const FtnScene* s, FtnStats* st) {
    if (!s || !st) return set_error(FTN_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(st, 0, sizeof(*st));
    st->bvh_build_seconds = s->build_seconds; st->morton_sort_seconds = s->sort_seconds; st->bvh_nodes = s->n_nodes; st->bvh_node_bytes = s->wide ? FTN_NODE8_BYTES : FTN_NODE_BYTES; st->bvh_tri_bytes = FTN_TRI_BYTES;
    st->kernel_launches = ftn_kernel_launch_count();
    return FTN_OK;
}

FTN_API int ftn_intersect_device(const FtnScene* s, size_t n, const FtnRay* d_rays, FtnHit* d_hits, void* stream) {
    return intersect_device(s, n, d_rays, d_hits, nullptr, false, nullptr, (cudaStream_t)stream);
}
FTN_API int ftn_intersect_test_device(const FtnScene* s, size_t n, const FtnRay* d_rays, uint8_t* d_out, void* stream) {
    return intersect_device(s, n, d_rays, nullptr, d_out, true, nullptr, (cudaStream_t)stream);
}
FTN_API int ftn_intersect_count_device(const FtnScene* s, size_t n, const FtnRay* d_rays, FtnHit* d_hits, uint64_t* d_counters, void* stream) {
    if (!d_counters) return set_error(FTN_ERR_INVALID_ARGUMENT, "null counters");
    return intersect_device(s, n, d_rays, d_hits, nullptr, false, (unsigned long long*)d_counters, (cudaStream_t)stream);
}

// Host-buffer queries: device staging from the per-device arena (no allocation per call) and a three-stream
// pipeline over chunks of the batch -- upload of chunk i+1, traversal of chunk i and download of chunk i-1
// overlap (truly so when the caller's buffers are page-locked, e.g. from ftn_host_alloc; with pageable memory the
// copies degrade to staged ones and the result is the same).
static int intersect_host(const FtnScene* s, size_t n, const FtnRay* rays, FtnHit* hits, uint8_t* any_out, bool any) {
    if (!s) return set_error(FTN_ERR_INVALID_ARGUMENT, "null scene");
    if (n == 0) return FTN_OK;
    if (!rays || (!any && !hits) || (any && !any_out)) return set_error(FTN_ERR_INVALID_ARGUMENT, "null buffer");
    FTN_CUDA(cudaSetDevice(s->device));
    const size_t out_elem = any ? 1 : sizeof(FtnHit);
    const size_t in_bytes = (n * sizeof(FtnRay) + 255) & ~(size_t)255;
    DeviceArena& arena = device_arena(s->device);
    std::lock_guard<std::recursive_mutex> lock(arena.m);
    char* base = nullptr;
    FTN_TRY(arena.reserve(DeviceArena::BATCH, in_bytes + n * out_elem + 256, "cudaMalloc (ray batch staging)", (void**)&base));
    FtnRay* d_rays = (FtnRay*)base;
    char* d_out = base + in_bytes;
    cudaStream_t st_in = nullptr, st_k = nullptr, st_out = nullptr;
    const int RING = 4;
    cudaEvent_t ev_in[RING], ev_k[RING], ev_out[RING];
    int n_ev = 0;
    int rc = FTN_OK;
    cudaError_t e = cudaSuccess;
    auto fail = [&](const char* what) { rc = cuda_fail(e, what, __FILE__, __LINE__); };
    if ((e = cudaStreamCreateWithFlags(&st_in, cudaStreamNonBlocking)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&st_k, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&st_out, cudaStreamNonBlocking)) != cudaSuccess) fail("cudaStreamCreate");
    for (; rc == FTN_OK && n_ev < RING; ++n_ev)
        if ((e = cudaEventCreateWithFlags(&ev_in[n_ev], cudaEventDisableTiming)) != cudaSuccess || (e = cudaEventCreateWithFlags(&ev_k[n_ev], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&ev_out[n_ev], cudaEventDisableTiming)) != cudaSuccess) { fail("cudaEventCreate"); break; }
    const size_t chunk = (size_t)1 << 19;     // 16 MiB of rays per stage
    size_t ci = 0;
    for (size_t off = 0; off < n && rc == FTN_OK; off += chunk, ++ci) {
        const size_t m = n - off < chunk ? n - off : chunk;
        const int r = (int)(ci % RING);
        if (ci >= (size_t)RING && (e = cudaEventSynchronize(ev_out[r])) != cudaSuccess) { fail("event sync"); break; }   // ring slot free again
        if ((e = cudaMemcpyAsync(d_rays + off, rays + off, m * sizeof(FtnRay), cudaMemcpyHostToDevice, st_in)) != cudaSuccess) { fail("H2D rays"); break; }
        cudaEventRecord(ev_in[r], st_in);
        cudaStreamWaitEvent(st_k, ev_in[r], 0);
        rc = intersect_device(s, m, d_rays + off, any ? nullptr : (FtnHit*)d_out + off, any ? (uint8_t*)d_out + off : nullptr, any, nullptr, st_k);
        if (rc != FTN_OK) break;
        cudaEventRecord(ev_k[r], st_k);
        cudaStreamWaitEvent(st_out, ev_k[r], 0);
        void* dst = any ? (void*)(any_out + off) : (void*)(hits + off);
        if ((e = cudaMemcpyAsync(dst, d_out + off * out_elem, m * out_elem, cudaMemcpyDeviceToHost, st_out)) != cudaSuccess) { fail("D2H hits"); break; }
        cudaEventRecord(ev_out[r], st_out);
    }
    for (cudaStream_t st : {st_in, st_k, st_out}) if (st) { cudaError_t e2 = cudaStreamSynchronize(st); if (rc == FTN_OK && e2 != cudaSuccess) { e = e2; fail("batch sync"); } }
    for (int i = 0; i < n_ev; ++i) { cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_k[i]); cudaEventDestroy(ev_out[i]); }
    for (cudaStream_t st : {st_in, st_k, st_out}) if (st) cudaStreamDestroy(st);
    return rc;
}
FTN_API int ftn_intersect(const FtnScene* s, size_t n, const FtnRay* rays, FtnHit* hits) { return intersect_host(s, n, rays, hits, nullptr, false); }
FTN_API int ftn_intersect_test(const FtnScene* s, size_t n, const FtnRay* rays, uint8_t* out) { return intersect_host(s, n, rays, nullptr, out, true); }

FTN_API int ftn_render_device(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                              const FtnIntegrator* integ, FtnPixel* d_pixels, FtnStats* stats, void* stream) {
    const uint64_t l0 = ftn_kernel_launch_count();
    const int rc = render_device(s, cam, film, smp, integ, d_pixels, stats, (cudaStream_t)stream);
    if (stats) stats->kernel_launches = ftn_kernel_launch_count() - l0;
    return rc;
}
FTN_API int ftn_render(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                       const FtnIntegrator* integ, FtnPixel* out_pixels, FtnStats* stats) {
    return render_host(s, cam, film, smp, integ, out_pixels, stats);
}
FTN_API int ftn_render_multi(FtnScene* const* scenes, int32_t n_scenes, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                             const FtnIntegrator* integ, FtnPixel* out_pixels, FtnStats* stats) {
    const uint64_t l0 = ftn_kernel_launch_count();
    const int rc = render_multi(scenes, n_scenes, cam, film, smp, integ, out_pixels, stats);
    if (stats) stats->kernel_launches = ftn_kernel_launch_count() - l0;
    return rc;
}
FTN_API int ftn_release_cached_memory(void) { return release_cached_memory(); }
FTN_API int ftn_host_alloc(size_t bytes, void** out) {
    if (!out) return set_error(FTN_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    FTN_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return FTN_OK;
}
FTN_API int ftn_host_free(void* p) {
    if (p) FTN_CUDA(cudaFreeHost(p));
    return FTN_OK;
}
FTN_API int ftn_film_to_rgb_device(size_t n, const FtnPixel* d_pixels, float* d_rgb, void* stream) {
    return film_to_rgb_device(n, d_pixels, d_rgb, (cudaStream_t)stream);
}
FTN_API int ftn_film_pixel_count(const FtnFilm* film, int32_t* w, int32_t* h) { return film_pixel_count(film, w, h); }

}  // extern "C"
