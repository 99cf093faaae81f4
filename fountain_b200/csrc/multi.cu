// ftn_render_multi: ONE host process driving N GPUs of a box (what the reference's render binary,
// src/bin/render.rs:82-89, would call on a multi-GPU machine; bench.py's torchrun path is the
// one-process-per-GPU form of the same thing).  The scene is replicated (one FtnScene per device), the
// samples are sharded by index, and the only exchange is ONE ncclReduce of the partial films over
// NVLink -- the analogue of merge_film_tile's mutex merge (film.rs:121-132).
//
// NCCL is bound at run time (dlopen "libnccl.so.2" + dlsym): a host that already carries an NCCL
// (a Python process with torch's bundled copy) shares it instead of loading a second one, and the
// library itself has no link-time dependency on it.
#include "ftn_scene.h"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <map>
#include <thread>

namespace ftn {

int render_device(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                  const FtnIntegrator* integ, FtnPixel* d_pixels, FtnStats* stats, cudaStream_t st);
int render_host(const FtnScene* s, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                const FtnIntegrator* integ, FtnPixel* out_pixels, FtnStats* stats);
int film_pixel_count(const FtnFilm* f, int32_t* w, int32_t* h);

namespace {
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};
NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!api.lib) { const char* e = dlerror(); api.error = std::string("libnccl.so.2 could not be loaded: ") + (e ? e : "?"); return; }
        auto sym = [&](const char* name) -> void* {
            void* p = dlsym(api.lib, name);
            if (!p && api.error.empty()) api.error = std::string("libnccl.so.2 lacks ") + name;
            return p;
        };
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

// communicators + one stream per device, kept per device list (ncclCommInitAll costs ~a second for 8 GPUs)
struct MultiCtx {
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;
};
std::mutex g_multi_mutex;
std::map<std::vector<int>, MultiCtx*> g_multi;

int nccl_fail(ncclResult_t r, const char* what) {
    const NcclApi& api = nccl_api();
    return set_error(FTN_ERR_CUDA, std::string(what) + " failed: " + (api.GetErrorString ? api.GetErrorString(r) : "NCCL error"));
}

int multi_ctx(const std::vector<int>& devices, MultiCtx** out) {
    NcclApi& api = nccl_api();
    if (!api.error.empty()) return set_error(FTN_ERR_UNSUPPORTED, api.error);
    auto it = g_multi.find(devices);
    if (it != g_multi.end()) { *out = it->second; return FTN_OK; }
    MultiCtx* c = new MultiCtx();
    c->devices = devices;
    c->comms.resize(devices.size());
    ncclResult_t r = api.CommInitAll(c->comms.data(), (int)devices.size(), devices.data());
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitAll"); }
    c->streams.resize(devices.size(), nullptr);
    for (size_t i = 0; i < devices.size(); ++i) {
        cudaError_t e = cudaSetDevice(devices[i]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) return cuda_fail(e, "stream for ftn_render_multi", __FILE__, __LINE__);   // the context stays cached half-built only on a broken device
    }
    g_multi[devices] = c;
    *out = c;
    return FTN_OK;
}
}  // namespace

int render_multi(FtnScene* const* scenes, int32_t n, const FtnCamera* cam, const FtnFilm* film, const FtnSampler* smp,
                 const FtnIntegrator* integ, FtnPixel* out_pixels, FtnStats* stats) {
    if (!scenes || n < 1 || !cam || !film || !smp || !integ || !out_pixels) return set_error(FTN_ERR_INVALID_ARGUMENT, "null argument");
    for (int i = 0; i < n; ++i) if (!scenes[i] || !scenes[i]->built) return set_error(FTN_ERR_INVALID_ARGUMENT, "every scene must be created and built on its device");
    if (n == 1) return render_host(scenes[0], cam, film, smp, integ, out_pixels, stats);
    if (smp->sample_stride < 1 || smp->sample_begin < 0) return set_error(FTN_ERR_INVALID_ARGUMENT, "bad sampler");
    std::vector<int> devices(n);
    for (int i = 0; i < n; ++i) {
        devices[i] = scenes[i]->device;
        for (int j = 0; j < i; ++j) if (devices[j] == devices[i]) return set_error(FTN_ERR_INVALID_ARGUMENT, "ftn_render_multi needs one scene per DISTINCT device");
    }
    int32_t w = 0, h = 0;
    FTN_TRY(film_pixel_count(film, &w, &h));
    const size_t n_px = (size_t)w * h, bytes = n_px * sizeof(FtnPixel);
    int prev_device = 0;
    cudaGetDevice(&prev_device);
    std::lock_guard<std::mutex> lock(g_multi_mutex);   // one multi-GPU render at a time per process (they would share every SM)
    MultiCtx* ctx = nullptr;
    FTN_TRY(multi_ctx(devices, &ctx));
    const uint32_t flags = stats ? stats->flags : 0u;

    std::vector<FtnPixel*> d_film(n, nullptr);
    std::vector<FtnStats> st(n);
    std::vector<int> rc(n, FTN_OK);
    std::vector<std::string> err(n);
    std::vector<std::thread> workers;
    for (int i = 0; i < n; ++i) {
        workers.emplace_back([&, i] {
            cudaError_t e = cudaSetDevice(devices[i]);
            if (e == cudaSuccess) e = cudaMallocAsync((void**)&d_film[i], bytes, ctx->streams[i]);
            if (e == cudaSuccess) e = cudaMemsetAsync(d_film[i], 0, bytes, ctx->streams[i]);
            if (e != cudaSuccess) { rc[i] = cuda_fail(e, "film of ftn_render_multi", __FILE__, __LINE__); err[i] = last_error_string(); return; }
            FtnSampler shard = *smp;     // device i renders s = begin + (i + k n) stride
            shard.sample_begin = smp->sample_begin + i * smp->sample_stride;
            shard.sample_stride = smp->sample_stride * n;
            std::memset(&st[i], 0, sizeof(FtnStats)); st[i].flags = flags;
            rc[i] = render_device(scenes[i], cam, film, &shard, integ, d_film[i], &st[i], ctx->streams[i]);
            if (rc[i] != FTN_OK) err[i] = last_error_string();
        });
    }
    for (std::thread& t : workers) t.join();
    int result = FTN_OK;
    for (int i = 0; i < n; ++i) if (rc[i] != FTN_OK && result == FTN_OK) result = set_error(rc[i], err[i]);
    const bool readable = result == FTN_OK || result == FTN_ERR_NAN_RADIANCE || result == FTN_ERR_UNSUPPORTED;   // as ftn_render
    bool have_all = true;
    for (int i = 0; i < n; ++i) if (!d_film[i]) have_all = false;
    if (readable && have_all) {
        const std::string keep = last_error_string();
        const NcclApi& api = nccl_api();
        ncclResult_t r = api.GroupStart();
        for (int i = 0; i < n && r == ncclSuccess; ++i)
            r = api.Reduce(d_film[i], d_film[i], n_px * 4, ncclFloat32, ncclSum, 0, ctx->comms[i], ctx->streams[i]);
        const ncclResult_t r2 = api.GroupEnd();
        if (r == ncclSuccess) r = r2;
        if (r != ncclSuccess) result = nccl_fail(r, "ncclReduce (film)");
        else {
            cudaError_t e = cudaSetDevice(devices[0]);
            if (e == cudaSuccess) e = cudaMemcpyAsync(out_pixels, d_film[0], bytes, cudaMemcpyDeviceToHost, ctx->streams[0]);
            for (int i = 0; i < n && e == cudaSuccess; ++i) { e = cudaSetDevice(devices[i]); if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->streams[i]); }
            if (e != cudaSuccess) result = cuda_fail(e, "film read-back of ftn_render_multi", __FILE__, __LINE__);
            else restore_error_string(keep);
        }
    }
    for (int i = 0; i < n; ++i) if (d_film[i]) { cudaSetDevice(devices[i]); cudaFreeAsync(d_film[i], ctx->streams[i]); }
    cudaSetDevice(prev_device);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->flags = flags;
        for (int i = 0; i < n; ++i) {
            const FtnStats& a = st[i];
            stats->camera_samples += a.camera_samples; stats->rays_closest += a.rays_closest; stats->rays_any += a.rays_any;
            stats->node_visits += a.node_visits; stats->tri_tests += a.tri_tests; stats->kernel_launches += a.kernel_launches;
            if (a.device_seconds > stats->device_seconds) stats->device_seconds = a.device_seconds;
            for (int c = 0; c < 3; ++c) {
                stats->trace_seconds[c] += a.trace_seconds[c]; stats->trace_launches[c] += a.trace_launches[c];
                stats->trace_rays[c] += a.trace_rays[c]; stats->trace_nodes[c] += a.trace_nodes[c]; stats->trace_tris[c] += a.trace_tris[c];
            }
            stats->shade_seconds += a.shade_seconds; stats->shade_launches += a.shade_launches;
        }
        stats->bvh_build_seconds = st[0].bvh_build_seconds; stats->bvh_nodes = st[0].bvh_nodes;
        stats->bvh_node_bytes = st[0].bvh_node_bytes; stats->bvh_tri_bytes = st[0].bvh_tri_bytes;
    }
    return result;
}

}  // namespace ftn
