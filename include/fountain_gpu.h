/*
 * fountain_gpu.h -- C ABI of the B200-native wavefront path tracer that replaces
 * fountain's data-parallel hot path (camera ray -> BVH closest/any hit -> shade ->
 * light MIS -> film).  Plain C: pointers and sizes only, no C++/torch types.
 *
 * Every entry point names the reference interface it replaces (file:line relative
 * to akofke/fountain).  The reference is pure Rust; a Rust host crate binds these
 * with `extern "C"` (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - All functions return 0 on success, a negative FtnStatus on failure; the
 *     message is available from ftn_last_error() (thread-local).  Nothing throws
 *     or aborts across the ABI.
 *   - The caller owns every input array and every output buffer.  The library
 *     copies inputs during ftn_scene_create and keeps nothing after return.
 *   - `*_device` variants take pointers into the scene's CUDA device's memory
 *     (e.g. a torch tensor's data_ptr) and run on `stream` (a cudaStream_t passed
 *     as void*; NULL = legacy default stream).  The ray queries
 *     (ftn_intersect*_device) and ftn_film_to_rgb_device only ENQUEUE: they return
 *     before the work has run and never block the host; up to FTN_MAX_QUERIES_IN_FLIGHT
 *     queries on one scene may overlap on different streams.  ftn_render_device is
 *     different: it runs on `stream` but RETURNS ONLY WHEN THE FILM IS COMPLETE (it
 *     reads queue lengths back while the bounces run and the statistics at the end), so
 *     the caller may use d_pixels on `stream` or after a stream sync straight away.
 *     Renders on one device serialise on the library's per-device workspace.
 *   - Primitive ids crossing the ABI are insertion indices: triangles first, in
 *     index-buffer order (mesh order, tri_id), then spheres.  The reference's own
 *     post-build permutation (bvh.rs:52) is not stable and is not exposed.
 *   - Matrices are 16 floats, COLUMN-major (cgmath Matrix4 layout: m[4*c + r]),
 *     as Transform::from_flat (geometry/transform.rs:34-42).
 */
#ifndef FOUNTAIN_GPU_H
#define FOUNTAIN_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define FTN_API __declspec(dllexport)
#else
#define FTN_API __attribute__((visibility("default")))
#endif

#define FTN_ABI_VERSION 3u
#define FTN_MAX_QUERIES_IN_FLIGHT 64
#define FTN_NO_HIT 0xFFFFFFFFu

typedef enum FtnStatus {
    FTN_OK = 0,
    FTN_ERR_INVALID_ARGUMENT = -1,
    FTN_ERR_CUDA = -2,             /* a CUDA call failed; message carries cudaGetErrorString */
    FTN_ERR_NO_DEVICE = -3,        /* no CUDA device: there is NO CPU fallback */
    FTN_ERR_NAN_RADIANCE = -4,     /* mirrors check_radiance panic, integrator/mod.rs:285 */
    FTN_ERR_UNSUPPORTED = -5,      /* e.g. env-map sample with map_pdf == 0 (infinite.rs:101 unimplemented!()) */
    FTN_ERR_OUT_OF_MEMORY = -6
} FtnStatus;

/* ---- rays and hits --------------------------------------------------------------- */

/* geometry/mod.rs:88-93 `Ray { origin, dir, t_max, time }` -- 32 bytes, same field order. */
typedef struct FtnRay {
    float o[3];
    float d[3];
    float t_max;
    float time;
} FtnRay;

/* Compact hit record (replaces the fat SurfaceInteraction of interaction.rs:61-84 at the
 * aggregate seam; everything else is re-derivable from prim + barycentrics, triangle.rs:246-250).
 * prim == FTN_NO_HIT on a miss.  For triangles b1,b2 are the barycentrics of vertices 1,2
 * (b0 = e0*inv_det is returned implicitly as the reference computes it, NOT as 1-b1-b2: use
 * ftn_intersect_full if b0 is needed bit-exactly).  For spheres b1,b2 hold (phi, unused). */
typedef struct FtnHit {
    uint32_t prim;
    float t;
    float b1;
    float b2;
} FtnHit;

/* ---- scene description ------------------------------------------------------------ */

enum { FTN_MESH_FLIP_NORMALS = 1u };   /* reverse_orientation ^ transform_swaps_handedness, shapes/mod.rs:27-29 */

/* One TriangleMesh (shapes/triangle.rs:10-26).  Vertices are ALREADY in world space, as
 * TriangleMesh::new leaves them (triangle.rs:42-58). */
typedef struct FtnMeshDesc {
    uint32_t first_tri;      /* first triangle in the scene-wide index buffer */
    uint32_t n_tris;
    int32_t  material_id;    /* index into materials[], -1 = no material (null BSDF, path.rs:76-80) */
    uint32_t flags;          /* FTN_MESH_* */
    /* `AreaLightSource "diffuse"` in front of the shape: EVERY triangle of the mesh carries its own
     * DiffuseAreaLight<Triangle> (loaders/pbrt.rs:275-316, light/diffuse.rs:24-93); the lights are
     * appended to the scene's light list in primitive order (scene/mod.rs:40-44). */
    int32_t  emissive;       /* 1 = every triangle is a diffuse area light */
    float    emit[3];        /* L of those lights */
} FtnMeshDesc;

typedef enum FtnMaterialType {
    FTN_MATERIAL_MATTE = 0,   /* material/matte.rs:36-52: Lambert for sigma == 0, Oren-Nayar otherwise */
    FTN_MATERIAL_METAL = 1,   /* material/metal.rs:38-65 */
    FTN_MATERIAL_PLASTIC = 2, /* material/plastic.rs:24-48 */
    FTN_MATERIAL_MIRROR = 3,  /* material/mirror.rs:21-30: SpecularReflection with FresnelNoOp */
    /* material/glass.rs:52-96, ROUGH glass: MicrofacetReflection<TrowbridgeReitz, FresnelDielectric(1, eta)> (Kr) +
     * MicrofacetTransmission<TrowbridgeReitz>(Kt, 1, eta, Radiance) (reflection/mod.rs:365-443).  With remap_roughness
     * (the loader's default, constructors.rs:198-205) roughness 0 is remapped to a small non-zero alpha and is rough glass
     * too; a glass whose alphas are exactly 0 is FresnelSpecular = `todo!()` under the reference's path integrator
     * (glass.rs:66) and a two-branch recursion under its direct-lighting integrator: ftn_scene_create answers
     * FTN_ERR_UNSUPPORTED for it. */
    FTN_MATERIAL_GLASS = 4
} FtnMaterialType;

/* Texture of a spectrum parameter (src/texture): constant, the 2D checkerboard without anti-aliasing
 * (AAMethod::None, the only one implemented, checkerboard.rs:54-64) or the uv debug texture (uv.rs), both
 * through UVMapping (mapping.rs:36-53: st = scale * uv + delta), or an image texture (image.rs:26-34): the
 * host hands over the MIPMap pyramid it built (mipmap.rs:78-143) and the device does lookup_trilinear
 * (mipmap.rs:245-279) with the ray's differentials (interaction.rs:124-176): the camera ray's (camera/mod.rs:145-205,
 * scaled by 1/sqrt(spp), integrator/mod.rs:249), which the path integrator hands on unchanged to every ray it
 * spawns (path.rs:73,79), and under the direct-lighting integrator the differentials specular_reflect derives for
 * the mirrored ray from the hit's dndu / dndv (integrator/mod.rs:59-83). */
typedef enum FtnTextureType {
    FTN_TEXTURE_CONSTANT = 0,      /* texture/mod.rs:34-42: the value in the material's kd field */
    FTN_TEXTURE_CHECKERBOARD = 1,  /* tex1 where (floor(s) + floor(t)) % 2 == 0, else tex2 */
    FTN_TEXTURE_UV = 2,            /* (s - floor(s), t - floor(t), 0) */
    FTN_TEXTURE_IMAGE = 3          /* ImageTexture<Spectrum, UVMapping>: MIPMap::lookup_trilinear */
} FtnTextureType;

/* mipmap.rs:15-17 `ImageWrap` (texel coordinates outside a level, mipmap.rs:297-311). */
typedef enum FtnImageWrap {
    FTN_WRAP_REPEAT = 0,           /* rem_euclid */
    FTN_WRAP_BLACK = 1,            /* 0 outside */
    FTN_WRAP_CLAMP = 2
} FtnImageWrap;
#define FTN_MAX_MIP_LEVELS 16

/* The textured parameters of the materials (loaders/constructors.rs:192-238: every one of them is read with
 * get_texture_or_default / get_texture_or_const).  Spectrum parameters take any FtnTextureType; float parameters
 * (roughnesses, sigma, the glass index) take the float textures the loader knows -- constant and checkerboard
 * (pbrt.rs:372) -- and read component 0 of the texture's value. */
typedef enum FtnMaterialParam {
    FTN_PARAM_KD = 0,          /* matte, plastic */
    FTN_PARAM_KS = 1,          /* plastic */
    FTN_PARAM_ETA = 2,         /* metal (spectrum) */
    FTN_PARAM_K = 3,           /* metal (spectrum) */
    FTN_PARAM_KR = 4,          /* mirror, glass */
    FTN_PARAM_KT = 5,          /* glass */
    FTN_PARAM_UROUGHNESS = 6,  /* metal, glass; plastic / isotropic metal: `roughness` */
    FTN_PARAM_VROUGHNESS = 7,  /* metal, glass */
    FTN_PARAM_SIGMA = 8,       /* matte, degrees */
    FTN_PARAM_INDEX = 9,       /* glass `eta` (a float texture in the loader) */
    FTN_PARAM_COUNT = 10
} FtnMaterialParam;

/* One entry of the scene's texture table (FtnSceneDesc::textures). */
typedef struct FtnTexture {
    int32_t type;            /* FtnTextureType */
    float value[3];          /* CONSTANT */
    float tex1[3], tex2[3];  /* CHECKERBOARD */
    float uv_scale[2];       /* UVMapping (constructors.rs:247-261) */
    float uv_delta[2];
    const float* image;      /* IMAGE: the MIPMap pyramid, laid out as FtnMaterial::image */
    int32_t image_width, image_height, image_levels, image_wrap;
} FtnTexture;

/* Materials.  Every parameter is the constant below unless param_texture[] names an entry of the texture table; Kd of
 * matte / plastic and Kr of mirror may also use the inline slot (kd_texture and the fields after it), kept from ABI v2. */
typedef struct FtnMaterial {
    int32_t type;            /* FtnMaterialType */
    float kd[3];             /* matte Kd / plastic Kd */
    float ks[3];             /* plastic Ks */
    float eta[3];            /* metal eta; glass: eta[0] = index of refraction (constructors.rs:203, default 1.5) */
    float k[3];              /* metal k */
    float u_roughness;       /* metal uroughness / plastic+metal isotropic roughness */
    float v_roughness;
    int32_t remap_roughness; /* constructors.rs:227 default true */
    float kr[3];             /* mirror Kr (constructors.rs:207-210 default 0.9); glass Kr (default 1) */
    int32_t kd_texture;      /* FtnTextureType of Kd (matte, plastic) or of Kr (mirror) */
    float tex1[3], tex2[3];  /* checkerboard: the two constant sub-textures (constructors.rs:276-287) */
    float uv_scale[2];       /* UVMapping uscale, vscale (constructors.rs:251-252, default 1) */
    float uv_delta[2];       /* UVMapping udelta, vdelta (default 0) */
    float sigma;             /* matte: Oren-Nayar roughness in DEGREES, clamped to [0, 90] (matte.rs:42); 0 = Lambert */
    /* IMAGE: MIPMap::pyramid (mipmap.rs:19-23), levels concatenated from level 0; each level RGB f32, row-major
     * with s fastest (texel (s, t) of a level of width w at 3 * (t * w + s)); level l is
     * max(1, width >> l) x max(1, height >> l) and there are 1 + floor(log2(max(width, height))) levels
     * (mipmap.rs:107-121).  Copied at scene creation. */
    const float* image;
    int32_t image_width, image_height;
    int32_t image_levels;    /* checked against the rule above */
    int32_t image_wrap;      /* FtnImageWrap */
    float kt[3];             /* glass Kt (constructors.rs:200, default 1) */
    uint32_t param_texture[FTN_PARAM_COUNT];   /* per FtnMaterialParam: 0 = the constant above; k = FtnSceneDesc::textures[k - 1] */
} FtnMaterial;

/* shapes/sphere.rs:16-27 (+ the DiffuseAreaLight it may carry, light/diffuse.rs:24-41). */
typedef struct FtnSphere {
    float object_to_world[16];
    float world_to_object[16];
    float radius;
    float z_min, z_max;      /* as given to Sphere::new (sphere.rs:30-49), before clamping */
    float phi_max_deg;
    int32_t reverse_orientation;
    int32_t material_id;     /* -1 = none */
    int32_t emissive;        /* 1 = has a diffuse area light */
    float emit[3];           /* L of the area light */
} FtnSphere;

typedef enum FtnLightType {
    FTN_LIGHT_INFINITE = 0,  /* light/infinite.rs:13-165 */
    FTN_LIGHT_POINT = 1,     /* light/point.rs:9-66   (delta position) */
    FTN_LIGHT_DISTANT = 2    /* light/distant.rs:10-75 (delta direction) */
} FtnLightType;

/* Explicit `LightSource`s in file order (scene/mod.rs:32-49).  Area lights are implied by
 * emissive spheres and are appended after these, in primitive order. */
typedef struct FtnLight {
    int32_t type;            /* FtnLightType */
    const float* texels;     /* INFINITE: RGB f32, w*h*3, row-major (s fastest); level-0 of the MIPMap */
    int32_t width, height;   /* INFINITE: 1x1 for new_uniform (infinite.rs:42-61) */
    float light_to_world[16];
    float world_to_light[16];
    float point[3];          /* POINT: world-space position, light_to_world * origin (point.rs:20) */
    float direction[3];      /* DISTANT: NORMALISED direction towards the light (distant.rs:22; the host normalises) */
    float intensity[3];      /* POINT: I, radiance = I / distance^2 (point.rs:56); DISTANT: L (distant.rs:66) */
} FtnLight;

typedef struct FtnSceneDesc {
    uint32_t abi_version;    /* FTN_ABI_VERSION */
    const float* positions;  /* 3*n_vertices, world space */
    const float* normals;    /* 3*n_vertices or NULL (un-normalised is fine, triangle.rs:335) */
    const float* uvs;        /* 2*n_vertices or NULL (default uv (0,0),(1,0),(1,1), triangle.rs:131-143) */
    uint32_t n_vertices;
    const uint32_t* indices; /* 3*n_triangles */
    uint32_t n_triangles;
    const FtnMeshDesc* meshes;
    uint32_t n_meshes;
    const FtnSphere* spheres;
    uint32_t n_spheres;
    const FtnMaterial* materials;
    uint32_t n_materials;
    const FtnLight* lights;
    uint32_t n_lights;
    const FtnTexture* textures;  /* the texture table FtnMaterial::param_texture indexes (may be NULL when n_textures == 0) */
    uint32_t n_textures;
} FtnSceneDesc;

/* ---- sensor ----------------------------------------------------------------------- */

/* camera/mod.rs:74-114 PerspectiveCamera after construction. */
typedef struct FtnCamera {
    float camera_to_world[16];
    float raster_to_camera[16];
    float lens_radius;
    float focal_distance;
    float shutter_open, shutter_close;
} FtnCamera;

/* film.rs:43-81 Film::new + filter/mod.rs:10-32 BoxFilter. */
typedef struct FtnFilm {
    int32_t x_resolution, y_resolution;
    float crop_window[4];    /* x0,x1,y0,y1 as the pbrt `cropwindow` (pbrt.rs:491-495) */
    float filter_radius[2];  /* box filter radius (0.5,0.5 default) */
} FtnFilm;

typedef enum FtnSamplerMode {
    /* Counter-based stream: u(pixel, sample, dimension) -- the GPU's native mode. */
    FTN_SAMPLER_COUNTER = 0,
    /* The reference's sequential per-tile xoshiro256+ stream (sampler/random.rs:6-76,
     * integrator/mod.rs:182-204).  Oracle only; the GPU library rejects it. */
    FTN_SAMPLER_REFERENCE_TILE_STREAM = 1
} FtnSamplerMode;

typedef struct FtnSampler {
    int32_t samples_per_pixel;
    uint64_t seed;
    int32_t mode;            /* FtnSamplerMode */
    /* sample-index sharding for multi-GPU: this call renders samples
     * s = sample_begin + i*sample_stride < samples_per_pixel.  (0,1) renders all. */
    int32_t sample_begin;
    int32_t sample_stride;
} FtnSampler;

typedef enum FtnIntegratorType {
    FTN_INTEGRATOR_PATH = 0,            /* integrator/path.rs:10-96 */
    FTN_INTEGRATOR_DIRECT_LIGHTING = 1  /* integrator/direct_lighting.rs (UniformSampleOne) */
} FtnIntegratorType;

typedef struct FtnIntegrator {
    int32_t type;
    int32_t max_depth;       /* PathIntegrator::new(max_depth, rr_threshold), render.rs:76 uses (5, 1.0) */
    float rr_threshold;
} FtnIntegrator;

/* film.rs:12-16 `Pixel { xyz, filter_weight_sum }` -- 16 bytes. */
typedef struct FtnPixel {
    float xyz[3];
    float filter_weight_sum;
} FtnPixel;

typedef struct FtnStats {
    uint64_t camera_samples;
    uint64_t rays_closest;   /* Scene::intersect calls (primary + continuation + MIS) */
    uint64_t rays_any;       /* Scene::intersect_test calls (shadow) */
    uint64_t node_visits;    /* 0 unless FTN_STATS_COUNT_TRAVERSAL was set in `flags` */
    uint64_t tri_tests;
    uint64_t kernel_launches;
    double   device_seconds; /* CUDA-event time of the whole call on its stream */
    double   bvh_build_seconds;
    uint32_t bvh_nodes;
    uint32_t bvh_node_bytes;
    uint32_t bvh_tri_bytes;
    uint32_t flags;          /* INPUT to ftn_render*: FTN_STATS_* bits; set before the call (0 = plain render) */
    /* per traversal-kernel class: 0 = extend (closest hit), 1 = shadow (any hit), 2 = MIS */
    double   trace_seconds[3];   /* FTN_STATS_TIME_KERNELS: sum of CUDA-event durations of that class's launches */
    uint64_t trace_launches[3];
    uint64_t trace_rays[3];
    uint64_t trace_nodes[3];     /* FTN_STATS_COUNT_TRAVERSAL */
    uint64_t trace_tris[3];
    double   shade_seconds;      /* FTN_STATS_TIME_KERNELS: all shading launches (miss, null, per material class) */
    uint64_t shade_launches;
    double   morton_sort_seconds; /* inside bvh_build_seconds: Morton codes (k_morton) + the stable radix sort of (code, primitive) */
} FtnStats;
/* FtnStats.flags (the only INPUT field of the struct; everything else is written by the call) */
#define FTN_STATS_COUNT_TRAVERSAL 1u   /* also count node visits / triangle tests (slower kernels) */
#define FTN_STATS_TIME_KERNELS    2u   /* CUDA-event pairs around every traversal / shading launch */

typedef struct FtnScene FtnScene;

/* ---- library ---------------------------------------------------------------------- */

FTN_API uint32_t    ftn_abi_version(void);
FTN_API const char* ftn_last_error(void);
FTN_API int         ftn_device_count(int* out_count);
FTN_API int         ftn_set_device(int device);
/* Total kernels launched by this library in this process (the bench's gpu_launches). */
FTN_API uint64_t    ftn_kernel_launch_count(void);

/* ---- scene / aggregate (replaces Scene::new scene/mod.rs:32, BVH::build bvh.rs:27) ------ */

/* Uploads the scene into flat SoA device buffers.  Does not build the BVH. */
FTN_API int ftn_scene_create(const FtnSceneDesc* desc, FtnScene** out_scene);
FTN_API int ftn_scene_destroy(FtnScene* scene);

/* Morton-code LBVH build on the device: morton3 (morton.rs:3-36) of the primitive-bound
 * centroids normalised by Bounds3f::offset (bounds.rs:200-206) in the centroid bounds,
 * stable radix sort, Karras hierarchy, refit, wide-node collapse.  Idempotent. */
FTN_API int ftn_bvh_build(FtnScene* scene);

/* Bit-exact check hooks: the 30-bit Morton code of every triangle in INPUT order and the
 * sorted primitive order (ties by index).  Either pointer may be NULL. */
FTN_API int ftn_bvh_debug_morton(const FtnScene* scene, uint32_t* codes, uint32_t* order);

/* Scene::world_bound (scene/mod.rs:66): min xyz, max xyz. */
FTN_API int ftn_scene_world_bound(const FtnScene* scene, float out_min_max[6]);

FTN_API int ftn_scene_stats(const FtnScene* scene, FtnStats* out);

/* Scene::intersect (scene/mod.rs:51) over a batch: closest hit, HOST buffers. */
FTN_API int ftn_intersect(const FtnScene* scene, size_t n, const FtnRay* rays, FtnHit* hits);
/* Scene::intersect_test (scene/mod.rs:55) over a batch: out[i] = 1 if any hit. */
FTN_API int ftn_intersect_test(const FtnScene* scene, size_t n, const FtnRay* rays, uint8_t* out);

FTN_API int ftn_intersect_device(const FtnScene* scene, size_t n, const FtnRay* d_rays,
                                 FtnHit* d_hits, void* stream);
FTN_API int ftn_intersect_test_device(const FtnScene* scene, size_t n, const FtnRay* d_rays,
                                      uint8_t* d_out, void* stream);
/* Same as ftn_intersect_device but also accumulates node-visit / triangle-test counts
 * (two uint64 on the device: [0]=nodes, [1]=tris).  Used for the bytes-per-ray roofline. */
FTN_API int ftn_intersect_count_device(const FtnScene* scene, size_t n, const FtnRay* d_rays,
                                       FtnHit* d_hits, uint64_t* d_counters, void* stream);

/* ---- integrator (replaces SamplerIntegrator::render_parallel, integrator/mod.rs:218) ---- */

/* Renders and writes film.pixels (film.rs:24) for the cropped pixel bounds, row-major:
 * XYZ sums and filter-weight sums, exactly what merge_film_tile (film.rs:121) leaves. */
FTN_API int ftn_render(const FtnScene* scene, const FtnCamera* camera, const FtnFilm* film,
                       const FtnSampler* sampler, const FtnIntegrator* integrator,
                       FtnPixel* out_pixels, FtnStats* out_stats);

/* Device variant for multi-GPU sharding: ACCUMULATES this call's samples into d_pixels
 * (4 floats per pixel: xyz + weight; caller zero-fills).  The caller then sum-reduces
 * d_pixels across ranks (NCCL) -- the analogue of merge_film_tile's Mutex merge. */
FTN_API int ftn_render_device(const FtnScene* scene, const FtnCamera* camera, const FtnFilm* film,
                              const FtnSampler* sampler, const FtnIntegrator* integrator,
                              FtnPixel* d_pixels, FtnStats* out_stats, void* stream);

/* The render of one process driving SEVERAL GPUs (what render.rs:82-89 would call on a multi-GPU box):
 * scenes[i] is the same scene created and built on device i (ftn_set_device(i); ftn_scene_create; ftn_bvh_build
 * -- the scene is replicated, every GPU holds all of it).  The samples of `sampler` are sharded by index:
 * device i renders s = sample_begin + (i + k * n_scenes) * sample_stride, on its own host thread and stream;
 * the partial films are summed onto scenes[0]'s device with ONE ncclReduce over NVLink -- the analogue of
 * merge_film_tile's mutex merge (film.rs:121-132) -- and copied to out_pixels (host).  NCCL (libnccl.so.2) is
 * loaded on first use; without it the call fails with FTN_ERR_UNSUPPORTED.  out_stats sums the devices' ray
 * counts; device_seconds is the slowest device's.  n_scenes == 1 is ftn_render. */
FTN_API int ftn_render_multi(FtnScene* const* scenes, int32_t n_scenes, const FtnCamera* camera, const FtnFilm* film,
                             const FtnSampler* sampler, const FtnIntegrator* integrator,
                             FtnPixel* out_pixels, FtnStats* out_stats);

/* Film::into_spectrum_buffer (film.rs:195-210): XYZ -> RGB, divide by weight, clamp >= 0.
 * n pixels; out_rgb is 3 floats per pixel.  Device buffers. */
FTN_API int ftn_film_to_rgb_device(size_t n, const FtnPixel* d_pixels, float* d_rgb, void* stream);

/* The library keeps one grow-only device arena per GPU for the wavefront path state (sized by the
 * largest render so far) so that renders -- also of newly created scenes -- allocate nothing.
 * A render works in passes of up to 64 Mi paths (220 B per path: 14.8 GB of a B200's 180 GB); when
 * the arena has to grow, the pass is halved until it fits in 3/4 of the device's free memory, so
 * a GPU that is shared or holds a very large scene still renders (FTN_PATHS_PER_PASS overrides).
 * This returns that memory to the driver; the next render allocates again. */
FTN_API int ftn_release_cached_memory(void);

/* Page-locked host memory for buffers that cross the boundary every call (the film of ftn_render, ray / hit
 * batches of ftn_intersect): copies to and from it run at full PCIe rate and need no staging. */
FTN_API int ftn_host_alloc(size_t bytes, void** out);
FTN_API int ftn_host_free(void* p);

/* Number of pixels ftn_render writes for this film (cropped_pixel_bounds area, film.rs:49-58). */
FTN_API int ftn_film_pixel_count(const FtnFilm* film, int32_t* out_w, int32_t* out_h);

#ifdef __cplusplus
}
#endif
#endif /* FOUNTAIN_GPU_H */
