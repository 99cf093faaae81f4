// Internal (non-ABI) declarations shared by the .cu translation units of libfountain_gpu.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/fountain_gpu.h"
#include "ftn_bvh.cuh"
#include "ftn_bvh8.cuh"

// runtime bools -> template arguments B0, B1
#define FTN_BOOL2(f0, f1, CALL)                                                      \
    do {                                                                             \
        if (f0) { constexpr bool B0 = true;  if (f1) { constexpr bool B1 = true; CALL; } else { constexpr bool B1 = false; CALL; } } \
        else    { constexpr bool B0 = false; if (f1) { constexpr bool B1 = true; CALL; } else { constexpr bool B1 = false; CALL; } } \
    } while (0)
#define FTN_BOOL1(f0, CALL)                                                          \
    do {                                                                             \
        if (f0) { constexpr bool B0 = true; CALL; } else { constexpr bool B0 = false; CALL; } \
    } while (0)
#define FTN_BOOL3(f0, f1, f2, CALL)                                                  \
    do {                                                                             \
        if (f2) { constexpr bool B2 = true;  FTN_BOOL2(f0, f1, CALL); }              \
        else    { constexpr bool B2 = false; FTN_BOOL2(f0, f1, CALL); }              \
    } while (0)

namespace ftn {

// ---- error plumbing ------------------------------------------------------------------------------
#define FTN_MAX_DEVICES 64
int set_error(int code, const std::string& msg);
std::string last_error_string();
void restore_error_string(const std::string& s);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(uint64_t n = 1);

#define FTN_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) return ftn::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)
#define FTN_TRY(call)                 \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != FTN_OK) return rc__; \
    } while (0)
// after every kernel launch
#define FTN_LAUNCHED()                                   \
    do {                                                 \
        ftn::count_launch();                             \
        FTN_CUDA(cudaGetLastError());                    \
    } while (0)

// ---- device-side scene tables ----------------------------------------------------------------------
struct MeshData {
    uint32_t first_tri, n_tris;
    int32_t material;
    uint32_t flags;          // FTN_MESH_*
    int32_t light_base;      // emissive mesh: index of the area light of its first triangle (one DiffuseAreaLight per triangle,
                             // consecutive, loaders/pbrt.rs:275-316); -1 = not emissive
    int32_t pad[3];
};

#define FTN_CLASS_OREN_NAYAR 5   /* matte.rs:45-49: its own class keeps the Lambert shade kernel free of the rough-diffuse code */
#define FTN_N_CLASSES 6
struct MaterialData {
    int32_t type;            // material CLASS = shade queue: FtnMaterialType, or FTN_CLASS_OREN_NAYAR for a matte with sigma != 0
    float kd[3], ks[3], eta[3], k[3];   // mirror: Kr in kd;  glass: Kr in kd, Kt in ks, index of refraction in eta[0]
    float alpha_x, alpha_y;  // after the roughness remap (microfacet.rs:40-45)
    int32_t kd_texture;      // FtnTextureType of Kd
    float tex1[3], tex2[3], uv_scale[2], uv_delta[2];
    // FTN_TEXTURE_IMAGE: MIPMap::pyramid (mipmap.rs:19-23) as RGBA texels (A unused), levels concatenated from 0;
    // level l is max(1, w >> l) x max(1, h >> l)
    const F4* image;
    int32_t img_w, img_h, img_levels, img_wrap;
    // the texture table (FtnMaterial::param_texture): per FtnMaterialParam 0 = the constant above, k = SceneView::textures[k - 1];
    // the raw roughnesses / sigma / remap flag are kept for parameters that are evaluated per hit
    uint32_t ptex[FTN_PARAM_COUNT];
    float u_rough, v_rough, sigma;
    int32_t remap;
    int32_t uses_image;      // any of its textures is an image: the hit needs texture differentials
};
// one entry of the texture table; the field names of the image part match MaterialData's inline slot (shared mip code)
struct TextureData {
    int32_t type;            // FtnTextureType
    float v[3], tex1[3], tex2[3], uv_scale[2], uv_delta[2];
    const F4* image;
    int32_t img_w, img_h, img_levels, img_wrap;
};

// FtnMaterial's pyramid (RGB f32, include/fountain_gpu.h) -> RGBA texels; false when the description breaks the
// level rule of mipmap.rs:107-121.  Host-side, shared by scene.cu and the host harness of the tests.
template <class Vec, class Desc>
inline bool pack_image_pyramid(const Desc& fm, Vec* out) {
    if (!fm.image || fm.image_width < 1 || fm.image_height < 1 || fm.image_wrap < FTN_WRAP_REPEAT || fm.image_wrap > FTN_WRAP_CLAMP) return false;
    int expect = 1;
    for (int m = fm.image_width > fm.image_height ? fm.image_width : fm.image_height; m > 1; m >>= 1) ++expect;
    if (fm.image_levels != expect || expect > FTN_MAX_MIP_LEVELS) return false;
    size_t n = 0;
    for (int l = 0; l < expect; ++l) {
        const int lw = (fm.image_width >> l) > 1 ? (fm.image_width >> l) : 1, lh = (fm.image_height >> l) > 1 ? (fm.image_height >> l) : 1;
        n += (size_t)lw * lh;
    }
    out->resize(n);
    for (size_t k = 0; k < n; ++k) { F4 t; t.x = fm.image[3 * k]; t.y = fm.image[3 * k + 1]; t.z = fm.image[3 * k + 2]; t.w = 0.0f; (*out)[k] = t; }
    return true;
}

// light/infinite.rs: level-0 texels + the Distribution2D tables (sampling.rs:137-180)
struct EnvLightData {
    const F4* texels;        // w*h RGBA (A unused)
    int32_t w, h;
    int32_t nu, nv;          // distribution grid: nu = h, nv = w (infinite.rs:64 swaps them)
    const float* cond_func;  // nv * nu
    const float* cond_cdf;   // nv * (nu+1)
    const float* cond_integral;  // nv  (== marginal func)
    const float* marg_cdf;   // nv + 1
    // guide tables of the two cdf searches (ftn_shade.cuh env_guide_entry): entry g = number of cdf entries <= g / n of
    // that row, so the search for u only looks between entries floor(u n) - 1 and floor(u n) + 2.  Same index as the
    // full binary search (sampling.rs:66-81), ~3 dependent loads instead of ~11 per search.  nullptr = full search.
    const uint32_t* cond_guide;  // nv * (nu + 1)
    const uint32_t* marg_guide;  // nv + 1
    float marg_integral;
    int32_t levels;          // 1 + floor(log2(max(w,h)))  (mipmap.rs:103)
    M4 l2w, w2l;
    float world_radius;      // Scene::new -> preprocess (infinite.rs:93-97)
    float world_center[3];
};

struct LightData {
    int32_t type;            // 0 infinite, 1 diffuse area on a sphere, 2 point, 3 distant, 4 diffuse area on a triangle
    int32_t sphere;          // sphere area light: sphere index; triangle area light: primitive id (scene-wide triangle index)
    float emit[3];           // area: L; point: I; distant: L
    float vec[3];            // point: world position; distant: normalised direction towards the light
    int32_t mesh;            // triangle area light: mesh of the triangle
    EnvLightData env;
};

// Host side: a parameter whose table entry is a CONSTANT texture is that constant (texture/mod.rs:34-42) -- fold it into
// the material at scene creation, so the per-hit table lookup (and the per-hit roughness remap, whose device logf need not
// round like the host's) is kept for the textures that vary.  Ids out of range are left for material_params_from_abi to refuse.
inline FtnMaterial fold_constant_param_textures(const FtnMaterial& in, const FtnTexture* textures, uint32_t n_textures) {
    FtnMaterial fm = in;
    for (int p = 0; p < FTN_PARAM_COUNT; ++p) {
        const uint32_t id = fm.param_texture[p];
        if (!id || id > n_textures || !textures || textures[id - 1].type != FTN_TEXTURE_CONSTANT) continue;
        const float* v = textures[id - 1].value;
        const int t = fm.type;
        bool folded = true;
        if (p == FTN_PARAM_KD && (t == FTN_MATERIAL_MATTE || t == FTN_MATERIAL_PLASTIC)) { for (int c = 0; c < 3; ++c) fm.kd[c] = v[c]; fm.kd_texture = FTN_TEXTURE_CONSTANT; }
        else if (p == FTN_PARAM_KS && t == FTN_MATERIAL_PLASTIC) { for (int c = 0; c < 3; ++c) fm.ks[c] = v[c]; }
        else if (p == FTN_PARAM_ETA && t == FTN_MATERIAL_METAL) { for (int c = 0; c < 3; ++c) fm.eta[c] = v[c]; }
        else if (p == FTN_PARAM_K && t == FTN_MATERIAL_METAL) { for (int c = 0; c < 3; ++c) fm.k[c] = v[c]; }
        else if (p == FTN_PARAM_KR && (t == FTN_MATERIAL_MIRROR || t == FTN_MATERIAL_GLASS)) { for (int c = 0; c < 3; ++c) fm.kr[c] = v[c]; if (t == FTN_MATERIAL_MIRROR) fm.kd_texture = FTN_TEXTURE_CONSTANT; }
        else if (p == FTN_PARAM_KT && t == FTN_MATERIAL_GLASS) { for (int c = 0; c < 3; ++c) fm.kt[c] = v[c]; }
        else if (p == FTN_PARAM_UROUGHNESS && (t == FTN_MATERIAL_METAL || t == FTN_MATERIAL_GLASS || t == FTN_MATERIAL_PLASTIC)) fm.u_roughness = v[0];
        else if (p == FTN_PARAM_VROUGHNESS && (t == FTN_MATERIAL_METAL || t == FTN_MATERIAL_GLASS)) fm.v_roughness = v[0];
        else if (p == FTN_PARAM_SIGMA && t == FTN_MATERIAL_MATTE) fm.sigma = v[0];
        else if (p == FTN_PARAM_INDEX && t == FTN_MATERIAL_GLASS) fm.eta[0] = v[0];
        else folded = false;
        if (folded) fm.param_texture[p] = 0;
    }
    return fm;
}

// Host side, shared by scene.cu and the host harness of the tests: the per-parameter texture ids, raw roughnesses and the
// class of a matte whose sigma is textured.  `tex_is_image(k)` tells whether table entry k (1-based) is an image texture.
template <class IsImage>
inline int material_params_from_abi(const FtnMaterial& fm, uint32_t n_textures, IsImage tex_is_image, MaterialData* md) {
    md->uses_image = md->kd_texture == FTN_TEXTURE_IMAGE ? 1 : 0;
    for (int p = 0; p < FTN_PARAM_COUNT; ++p) {
        const uint32_t id = fm.param_texture[p];
        if (id > n_textures) return FTN_ERR_INVALID_ARGUMENT;
        md->ptex[p] = id;
        if (id && tex_is_image(id)) md->uses_image = 1;
    }
    md->u_rough = fm.u_roughness; md->v_rough = fm.type == FTN_MATERIAL_PLASTIC ? fm.u_roughness : fm.v_roughness;
    if (fm.type == FTN_MATERIAL_PLASTIC) md->ptex[FTN_PARAM_VROUGHNESS] = md->ptex[FTN_PARAM_UROUGHNESS];
    md->sigma = fm.sigma; md->remap = fm.remap_roughness ? 1 : 0;
    if (fm.type == FTN_MATERIAL_MATTE && md->ptex[FTN_PARAM_SIGMA]) md->type = FTN_CLASS_OREN_NAYAR;   // sigma decided per hit (0 => a = 1, b = 0 = Lambert)
    return FTN_OK;
}
template <class Vec>
inline int texture_from_abi(const FtnTexture& ft, TextureData* td, Vec* texels) {
    if (ft.type < FTN_TEXTURE_CONSTANT || ft.type > FTN_TEXTURE_IMAGE) return FTN_ERR_INVALID_ARGUMENT;
    std::memset(td, 0, sizeof(*td));
    td->type = ft.type;
    for (int c = 0; c < 3; ++c) { td->v[c] = ft.value[c]; td->tex1[c] = ft.tex1[c]; td->tex2[c] = ft.tex2[c]; }
    for (int c = 0; c < 2; ++c) { td->uv_scale[c] = ft.uv_scale[c]; td->uv_delta[c] = ft.uv_delta[c]; }
    if (ft.type == FTN_TEXTURE_IMAGE) {
        if (!pack_image_pyramid(ft, texels)) return FTN_ERR_INVALID_ARGUMENT;
        td->img_w = ft.image_width; td->img_h = ft.image_height; td->img_levels = ft.image_levels; td->img_wrap = ft.image_wrap;
    }
    return FTN_OK;
}

#define FTN_LIGHT_TYPE_TRIANGLE 4
// Fills MeshData (incl. light_base) and appends the per-triangle area lights of emissive meshes to `lights`, which must
// hold exactly the explicit lights (scene/mod.rs:32-49: explicit lights, then the primitives' area lights -- here in
// primitive order: triangles, then spheres).  Host side, shared by scene.cu and the host harness of the tests.
template <class MeshVec, class LightVec>
inline void build_mesh_table(const FtnSceneDesc* d, MeshVec* meshes, LightVec* lights) {
    meshes->resize(d->n_meshes);
    for (uint32_t m = 0; m < d->n_meshes; ++m) {
        const FtnMeshDesc& fm = d->meshes[m];
        MeshData md; std::memset(&md, 0, sizeof(md));
        md.first_tri = fm.first_tri; md.n_tris = fm.n_tris; md.material = fm.material_id; md.flags = fm.flags; md.light_base = -1;
        if (fm.emissive && fm.n_tris) {
            md.light_base = (int32_t)lights->size();
            for (uint32_t t = 0; t < fm.n_tris; ++t) {
                LightData ld; std::memset(&ld, 0, sizeof(ld));
                ld.type = FTN_LIGHT_TYPE_TRIANGLE; ld.sphere = (int32_t)(fm.first_tri + t); ld.mesh = (int32_t)m;
                ld.emit[0] = fm.emit[0]; ld.emit[1] = fm.emit[1]; ld.emit[2] = fm.emit[2];
                lights->push_back(ld);
            }
        }
        (*meshes)[m] = md;
    }
}

struct SceneView {
    BvhView bvh;
    const float* pos; const float* nrm; const float* uv;   // nrm / uv may be null
    const uint32_t* idx;
    const MeshData* meshes;
    const MaterialData* materials;
    const TextureData* textures;     // the scene's texture table (may be null)
    const SphereData* spheres; uint32_t n_spheres;
    const LightData* lights; uint32_t n_lights;
    uint32_t n_tris;
    int refill_threshold;    // persistent traversal: leave the traverse loop when fewer lanes are active
    bool vote;               // persistent traversal over BVH2x64: per-step node/leaf vote (large scenes) or while-while (small)
    int vote_bias;           // persistent traversal: node step when 16 * #node lanes >= vote_bias * #leaf lanes
};

}  // namespace ftn

struct FtnScene {
    int device = 0;
    uint32_t n_verts = 0, n_tris = 0, n_meshes = 0, n_spheres = 0, n_materials = 0, n_lights = 0;
    float* d_pos = nullptr; float* d_nrm = nullptr; float* d_uv = nullptr; uint32_t* d_idx = nullptr;
    ftn::MeshData* d_meshes = nullptr;
    ftn::MaterialData* d_materials = nullptr;
    ftn::TextureData* d_textures = nullptr;
    ftn::SphereData* d_spheres = nullptr;
    ftn::LightData* d_lights = nullptr;
    std::vector<ftn::SphereData> h_spheres;
    std::vector<ftn::LightData> h_lights;
    std::vector<void*> owned;        // every other device allocation (env tables, ...)
    // aggregate
    bool built = false;
    ftn::F4* d_nodes = nullptr; ftn::F4* d_tris = nullptr;
    uint32_t n_nodes = 0;
    bool wide = false;               // d_nodes holds BVH8q records (ftn_bvh8.cuh) instead of BVH2x64
    uint32_t bvh_levels = 0;         // wide layout: levels of the tree (bounds the traversal stack)
    uint32_t* d_codes = nullptr;     // Morton code per triangle, INPUT order
    uint32_t* d_order = nullptr;     // sorted primitive order
    float bounds[6] = {0, 0, 0, 0, 0, 0};
    double build_seconds = 0.0;
    double sort_seconds = 0.0;       // Morton codes + radix sort, inside build_seconds
    // dynamic work-fetch counters of the batch queries: a ring of FTN_MAX_QUERIES_IN_FLIGHT slots, one per call, so
    // that queries enqueued on different streams never share a counter
    unsigned long long* d_work = nullptr;
    mutable std::atomic<uint32_t> work_slot{0};
    bool material_present[FTN_N_CLASSES] = {false, false, false, false, false, false};   // which shade kernels a render launches
    bool has_image_texture = false;         // any Kd image texture: selects the shade kernels that carry the mip lookup
    bool has_null_material = false;         // any primitive with a null BSDF (path.rs:76-80)
    ftn::SceneView view() const;
};

namespace ftn {
int scene_create(const FtnSceneDesc* d, FtnScene** out);
int scene_destroy(FtnScene* s);
int bvh_build(FtnScene* s);
// device-wide exclusive scan of n uint32 (in place allowed: out may equal in); scratch =
// scan_scratch_elems(n) uint32.  Enqueue-only (no allocation, no synchronisation).
size_t scan_scratch_elems(size_t n);
int exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, size_t n, uint32_t* d_scratch, cudaStream_t st);
// stable LSD radix sort of (key,value) pairs on `bits` low bits; result in d_keys/d_vals; scratch =
// radix_sort_scratch_bytes(n) bytes.  Enqueue-only.
size_t radix_sort_scratch_bytes(size_t n);
int radix_sort_pairs(uint32_t* d_keys, uint32_t* d_vals, size_t n, int bits, void* d_scratch, cudaStream_t st);
// Process-wide grow-only device arenas (one set per GPU): the wavefront path state, the film of
// host-buffer renders, the temporaries of a BVH build.  Callers hold the arena's mutex while they use it.
struct DeviceArena {
    void* p[4] = {nullptr, nullptr, nullptr, nullptr}; size_t bytes[4] = {0, 0, 0, 0};
    std::recursive_mutex m;   // recursive: ftn_render holds it from the film reservation to the read-back, around render_device's own lock
    enum { PATHS = 0, FILM = 1, BUILD = 2, BATCH = 3 };
    int reserve(int which, size_t need, const char* what, void** out);
    void* h_pinned = nullptr; size_t h_pinned_bytes = 0;   // small pinned host block (queue-counter read-back)
    int pinned_counts(uint32_t** out, size_t bytes);
};
DeviceArena& device_arena(int device);
int release_cached_memory();
}  // namespace ftn
