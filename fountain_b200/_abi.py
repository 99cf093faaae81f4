"""ctypes mirror of include/fountain_gpu.h (struct layouts and function prototypes).

One definition serves both shared libraries that speak this ABI: the product
(`libfountain_gpu.so`, prefix ``ftn_``) and the test oracle (prefix ``orc_``), which is
bound from ``oracle/orc.py`` -- never from this package.
"""
import ctypes as C

FTN_ABI_VERSION = 3
FTN_MAX_QUERIES_IN_FLIGHT = 64
FTN_STATS_COUNT_TRAVERSAL, FTN_STATS_TIME_KERNELS = 1, 2
FTN_NO_HIT = 0xFFFFFFFF

FTN_OK = 0
FTN_ERR_INVALID_ARGUMENT = -1
FTN_ERR_CUDA = -2
FTN_ERR_NO_DEVICE = -3
FTN_ERR_NAN_RADIANCE = -4
FTN_ERR_UNSUPPORTED = -5
FTN_ERR_OUT_OF_MEMORY = -6

FTN_MESH_FLIP_NORMALS = 1
FTN_MATERIAL_MATTE, FTN_MATERIAL_METAL, FTN_MATERIAL_PLASTIC, FTN_MATERIAL_MIRROR, FTN_MATERIAL_GLASS = 0, 1, 2, 3, 4
FTN_TEXTURE_CONSTANT, FTN_TEXTURE_CHECKERBOARD, FTN_TEXTURE_UV, FTN_TEXTURE_IMAGE = 0, 1, 2, 3
FTN_WRAP_REPEAT, FTN_WRAP_BLACK, FTN_WRAP_CLAMP = 0, 1, 2
FTN_MAX_MIP_LEVELS = 16
FTN_LIGHT_INFINITE = 0
FTN_LIGHT_POINT = 1
FTN_LIGHT_DISTANT = 2
FTN_SAMPLER_COUNTER, FTN_SAMPLER_REFERENCE_TILE_STREAM = 0, 1
FTN_INTEGRATOR_PATH, FTN_INTEGRATOR_DIRECT_LIGHTING = 0, 1

f32 = C.c_float
u32 = C.c_uint32
i32 = C.c_int32
u64 = C.c_uint64


class FtnRay(C.Structure):
    _fields_ = [("o", f32 * 3), ("d", f32 * 3), ("t_max", f32), ("time", f32)]


class FtnHit(C.Structure):
    _fields_ = [("prim", u32), ("t", f32), ("b1", f32), ("b2", f32)]


class FtnMeshDesc(C.Structure):
    _fields_ = [("first_tri", u32), ("n_tris", u32), ("material_id", i32), ("flags", u32), ("emissive", i32), ("emit", f32 * 3)]


class FtnMaterial(C.Structure):
    _fields_ = [("type", i32), ("kd", f32 * 3), ("ks", f32 * 3), ("eta", f32 * 3), ("k", f32 * 3),
                ("u_roughness", f32), ("v_roughness", f32), ("remap_roughness", i32), ("kr", f32 * 3),
                ("kd_texture", i32), ("tex1", f32 * 3), ("tex2", f32 * 3), ("uv_scale", f32 * 2), ("uv_delta", f32 * 2), ("sigma", f32),
                ("image", C.POINTER(f32)), ("image_width", i32), ("image_height", i32), ("image_levels", i32), ("image_wrap", i32),
                ("kt", f32 * 3), ("param_texture", u32 * 10)]


class FtnTexture(C.Structure):
    _fields_ = [("type", i32), ("value", f32 * 3), ("tex1", f32 * 3), ("tex2", f32 * 3), ("uv_scale", f32 * 2), ("uv_delta", f32 * 2),
                ("image", C.POINTER(f32)), ("image_width", i32), ("image_height", i32), ("image_levels", i32), ("image_wrap", i32)]


(FTN_PARAM_KD, FTN_PARAM_KS, FTN_PARAM_ETA, FTN_PARAM_K, FTN_PARAM_KR, FTN_PARAM_KT, FTN_PARAM_UROUGHNESS, FTN_PARAM_VROUGHNESS,
 FTN_PARAM_SIGMA, FTN_PARAM_INDEX, FTN_PARAM_COUNT) = range(11)


class FtnSphere(C.Structure):
    _fields_ = [("object_to_world", f32 * 16), ("world_to_object", f32 * 16), ("radius", f32),
                ("z_min", f32), ("z_max", f32), ("phi_max_deg", f32), ("reverse_orientation", i32),
                ("material_id", i32), ("emissive", i32), ("emit", f32 * 3)]


class FtnLight(C.Structure):
    _fields_ = [("type", i32), ("texels", C.POINTER(f32)), ("width", i32), ("height", i32),
                ("light_to_world", f32 * 16), ("world_to_light", f32 * 16),
                ("point", f32 * 3), ("direction", f32 * 3), ("intensity", f32 * 3)]


class FtnSceneDesc(C.Structure):
    _fields_ = [("abi_version", u32), ("positions", C.POINTER(f32)), ("normals", C.POINTER(f32)),
                ("uvs", C.POINTER(f32)), ("n_vertices", u32), ("indices", C.POINTER(u32)),
                ("n_triangles", u32), ("meshes", C.POINTER(FtnMeshDesc)), ("n_meshes", u32),
                ("spheres", C.POINTER(FtnSphere)), ("n_spheres", u32),
                ("materials", C.POINTER(FtnMaterial)), ("n_materials", u32),
                ("lights", C.POINTER(FtnLight)), ("n_lights", u32),
                ("textures", C.POINTER(FtnTexture)), ("n_textures", u32)]


class FtnCamera(C.Structure):
    _fields_ = [("camera_to_world", f32 * 16), ("raster_to_camera", f32 * 16), ("lens_radius", f32),
                ("focal_distance", f32), ("shutter_open", f32), ("shutter_close", f32)]


class FtnFilm(C.Structure):
    _fields_ = [("x_resolution", i32), ("y_resolution", i32), ("crop_window", f32 * 4),
                ("filter_radius", f32 * 2)]


class FtnSampler(C.Structure):
    _fields_ = [("samples_per_pixel", i32), ("seed", u64), ("mode", i32), ("sample_begin", i32),
                ("sample_stride", i32)]


class FtnIntegrator(C.Structure):
    _fields_ = [("type", i32), ("max_depth", i32), ("rr_threshold", f32)]


class FtnPixel(C.Structure):
    _fields_ = [("xyz", f32 * 3), ("filter_weight_sum", f32)]


class FtnStats(C.Structure):
    _fields_ = [("camera_samples", u64), ("rays_closest", u64), ("rays_any", u64), ("node_visits", u64),
                ("tri_tests", u64), ("kernel_launches", u64), ("device_seconds", C.c_double),
                ("bvh_build_seconds", C.c_double), ("bvh_nodes", u32), ("bvh_node_bytes", u32),
                ("bvh_tri_bytes", u32), ("flags", u32), ("trace_seconds", C.c_double * 3),
                ("trace_launches", u64 * 3), ("trace_rays", u64 * 3), ("trace_nodes", u64 * 3), ("trace_tris", u64 * 3),
                ("shade_seconds", C.c_double), ("shade_launches", u64), ("morton_sort_seconds", C.c_double)]

    def as_dict(self):
        return {name: (list(getattr(self, name)) if hasattr(getattr(self, name), "__len__") else getattr(self, name))
                for name, _ in self._fields_}


VOIDP = C.c_void_p
P = C.POINTER

# name (without prefix) -> (restype, argtypes).  Exactly the symbols fountain_gpu.h declares.
PROTOTYPES = {
    "abi_version": (u32, []),
    "last_error": (C.c_char_p, []),
    "device_count": (C.c_int, [P(C.c_int)]),
    "set_device": (C.c_int, [C.c_int]),
    "kernel_launch_count": (u64, []),
    "scene_create": (C.c_int, [P(FtnSceneDesc), P(VOIDP)]),
    "scene_destroy": (C.c_int, [VOIDP]),
    "bvh_build": (C.c_int, [VOIDP]),
    "bvh_debug_morton": (C.c_int, [VOIDP, P(u32), P(u32)]),
    "scene_world_bound": (C.c_int, [VOIDP, P(f32)]),
    "scene_stats": (C.c_int, [VOIDP, P(FtnStats)]),
    "intersect": (C.c_int, [VOIDP, C.c_size_t, P(FtnRay), P(FtnHit)]),
    "intersect_test": (C.c_int, [VOIDP, C.c_size_t, P(FtnRay), P(C.c_uint8)]),
    "intersect_device": (C.c_int, [VOIDP, C.c_size_t, VOIDP, VOIDP, VOIDP]),
    "intersect_test_device": (C.c_int, [VOIDP, C.c_size_t, VOIDP, VOIDP, VOIDP]),
    "intersect_count_device": (C.c_int, [VOIDP, C.c_size_t, VOIDP, VOIDP, VOIDP, VOIDP]),
    "render": (C.c_int, [VOIDP, P(FtnCamera), P(FtnFilm), P(FtnSampler), P(FtnIntegrator), P(FtnPixel), P(FtnStats)]),
    "render_device": (C.c_int, [VOIDP, P(FtnCamera), P(FtnFilm), P(FtnSampler), P(FtnIntegrator), VOIDP, P(FtnStats), VOIDP]),
    "render_multi": (C.c_int, [P(VOIDP), i32, P(FtnCamera), P(FtnFilm), P(FtnSampler), P(FtnIntegrator), P(FtnPixel), P(FtnStats)]),
    "film_to_rgb_device": (C.c_int, [C.c_size_t, VOIDP, VOIDP, VOIDP]),
    "film_pixel_count": (C.c_int, [P(FtnFilm), P(i32), P(i32)]),
    "release_cached_memory": (C.c_int, []),
    "host_alloc": (C.c_int, [C.c_size_t, P(VOIDP)]),
    "host_free": (C.c_int, [VOIDP]),
}

# Subset the CPU oracle implements (host buffers only), plus its own extras bound in oracle/orc.py.
ORACLE_SUBSET = ("abi_version", "last_error", "scene_create", "scene_destroy", "bvh_build",
                 "bvh_debug_morton", "scene_world_bound", "scene_stats", "intersect", "intersect_test",
                 "render", "film_pixel_count")


def bind(lib, prefix, names):
    """Attach restype/argtypes for `names` and return {name: function}."""
    out = {}
    for name in names:
        fn = getattr(lib, prefix + name)   # AttributeError if the symbol is missing: fail loudly
        fn.restype, fn.argtypes = PROTOTYPES[name]
        out[name] = fn
    return out
