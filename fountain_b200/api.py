"""Host-side mirror of the reference's Scene / Primitive / Camera / Sampler / Film / Integrator
API for the hot path, written above the C ABI (include/fountain_gpu.h).

The reference is Rust and its toolchain is absent here, so this mirror is Python (what the
tests and bench drive) with a C++ twin in include/fountain_host.hpp; the Rust binding a
fountain maintainer would add is in INTEGRATION.md.  Names, argument meaning and error
behaviour follow the reference:

  TriangleMesh::new             src/shapes/triangle.rs:29      -> TriangleMesh
  Sphere::new                   src/shapes/sphere.rs:30        -> Sphere
  MatteMaterial / MetalMaterial src/material/{matte,metal}.rs  -> MatteMaterial, MetalMaterial
  InfiniteAreaLight::new_*      src/light/infinite.rs:23,42    -> InfiniteAreaLight
  Scene::new / BVH::build       src/scene/mod.rs:32, bvh.rs:27 -> Scene
  Scene::intersect(_test)       src/scene/mod.rs:51,55         -> Scene.intersect / intersect_test (batched)
  PerspectiveCamera::new        src/camera/mod.rs:85           -> PerspectiveCamera
  Film::new / into_spectrum_buffer  src/film.rs:43,195         -> Film
  RandomSampler::new_with_seed  src/sampler/random.rs:12       -> RandomSampler
  PathIntegrator::new           src/integrator/path.rs:16      -> PathIntegrator
  SamplerIntegrator::render_parallel  src/integrator/mod.rs:218 -> SamplerIntegrator.render_parallel

Every compute call goes through a `Backend` (a bound shared library).  The default backend is
the CUDA library; it raises if the extension or a GPU is missing -- there is no CPU path here.
"""
import ctypes as C
import weakref

import numpy as np

from . import _abi as A
from .transform import Transform


class FountainError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("fountain error %d: %s" % (code, msg))
        self.code = code


class Backend:
    """A shared library speaking the fountain_gpu.h ABI under `prefix`."""

    def __init__(self, lib, prefix, names, name):
        self.lib = lib
        self.prefix = prefix
        self.name = name
        self.fn = A.bind(lib, prefix, names)

    def call(self, fname, *args):
        rc = self.fn[fname](*args)
        if rc != A.FTN_OK:
            msg = self.fn["last_error"]()
            raise FountainError(rc, (msg or b"").decode("utf-8", "replace"))
        return rc

    def has(self, fname):
        return fname in self.fn


_default_backend = None


def default_backend():
    """The CUDA library.  Raises (never falls back) when it cannot be used."""
    global _default_backend
    if _default_backend is None:
        from .lib import load_gpu_backend
        _default_backend = load_gpu_backend()
    return _default_backend


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _spectrum(v):
    a = np.asarray(v, dtype=np.float32).reshape(-1)
    return np.repeat(a, 3) if a.size == 1 else a


# ---------------------------------------------------------------------------------------------
# shapes / materials / lights
# ---------------------------------------------------------------------------------------------

class TriangleMesh:
    """shapes/triangle.rs:29-74.  Vertices (and normals) are moved to world space here, on the
    host, exactly as `TriangleMesh::new` does; the device only ever sees world space."""

    def __init__(self, object_to_world, vertex_indices, vertices, normals=None, tex_coords=None,
                 reverse_orientation=False):
        idx = np.ascontiguousarray(vertex_indices, dtype=np.uint32).reshape(-1)
        assert idx.size % 3 == 0                      # triangle.rs:38
        self.vertex_indices = idx.reshape(-1, 3)
        self.n_triangles = self.vertex_indices.shape[0]
        v = _f32(vertices).reshape(-1, 3)
        self.vertices = object_to_world.apply_points_f32(v)
        self.normals = None
        if normals is not None:
            n = _f32(normals).reshape(-1, 3)
            assert n.shape[0] == v.shape[0]           # triangle.rs:47
            self.normals = object_to_world.apply_normals_f32(n)
        self.tex_coords = None
        if tex_coords is not None:
            self.tex_coords = _f32(tex_coords).reshape(-1, 2)
            assert self.tex_coords.shape[0] == v.shape[0]   # triangle.rs:61
        self.reverse_orientation = bool(reverse_orientation)
        self.object_to_world = object_to_world

    def flip_normals(self):                           # shapes/mod.rs:27-29
        return self.reverse_orientation ^ self.object_to_world.swaps_handedness()

    @staticmethod
    def from_ply(path, object_to_world=None, reverse_orientation=False):
        """make_triangle_mesh_from_ply, loaders/constructors.rs:94-190."""
        from .ply import load_ply_cached
        d = load_ply_cached(path)
        return TriangleMesh(object_to_world or Transform.identity(), d["indices"], d["vertices"],
                            d["normals"], d["uvs"], reverse_orientation)


class Sphere:
    """shapes/sphere.rs:30-58."""

    def __init__(self, object_to_world, reverse_orientation=False, radius=1.0, z_min=None, z_max=None,
                 phi_max=360.0):
        self.object_to_world = object_to_world
        self.reverse_orientation = bool(reverse_orientation)
        self.radius = float(radius)
        self.z_min = -self.radius if z_min is None else float(z_min)
        self.z_max = self.radius if z_max is None else float(z_max)
        self.phi_max = float(phi_max)


class UVMapping:
    """texture/mapping.rs:13-53; defaults uscale = vscale = 1, udelta = vdelta = 0 (constructors.rs:251-254)."""

    def __init__(self, scale_u=1.0, scale_v=1.0, offset_u=0.0, offset_v=0.0):
        self.scale = (float(scale_u), float(scale_v))
        self.offset = (float(offset_u), float(offset_v))


class Checkerboard2DTexture:
    """texture/checkerboard.rs:10-64 over two constant spectra (AAMethod::None)."""
    type = A.FTN_TEXTURE_CHECKERBOARD

    def __init__(self, tex1=0.0, tex2=1.0, mapping=None):
        self.tex1, self.tex2 = _spectrum(tex1), _spectrum(tex2)
        self.mapping = mapping or UVMapping()

    @staticmethod
    def default():   # checkerboard.rs:38-42
        return Checkerboard2DTexture(0.0, 1.0, UVMapping(10.0, 10.0, 0.0, 0.0))


class UVTexture:
    """texture/uv.rs:6-24."""
    type = A.FTN_TEXTURE_UV

    def __init__(self, mapping=None):
        self.tex1 = self.tex2 = _spectrum(0.0)
        self.mapping = mapping or UVMapping()


class MIPMap:
    """mipmap.rs:19-143 `MIPMap<Spectrum>`: the image pyramid an ImageTexture filters.  Level l is
    max(1, w >> l) x max(1, h >> l); there are 1 + floor(log2(max(w, h))) levels (:107-121).  The reference
    halves each level with the `resize` crate's Triangle filter (0.4.3, a third-party dependency that is not
    in the repository); `downsample` restates that filter for a factor of two -- weights (1, 3, 3, 1) / 8 per
    axis, renormalised at the borders -- and is host-side setup, outside the device path.  A caller that owns
    a pyramid (the Rust host) passes its levels through `levels=` unchanged."""

    WRAP = {"repeat": A.FTN_WRAP_REPEAT, "black": A.FTN_WRAP_BLACK, "clamp": A.FTN_WRAP_CLAMP}

    def __init__(self, image=None, wrap="repeat", levels=None):
        self.wrap = self.WRAP[wrap] if isinstance(wrap, str) else int(wrap)
        if levels is None:
            img = np.ascontiguousarray(image, dtype=np.float32)
            if img.ndim != 3 or img.shape[2] != 3:
                raise ValueError("MIPMap image must be (height, width, 3)")
            h, w = img.shape[:2]
            levels = [img]
            for _ in range(1, 1 + int(np.floor(np.log2(max(w, h))))):
                levels.append(self.downsample(levels[-1]))
        self.levels = [np.ascontiguousarray(l, dtype=np.float32) for l in levels]
        self.height, self.width = self.levels[0].shape[:2]
        for l, lv in enumerate(self.levels):
            if lv.shape != (max(1, self.height >> l), max(1, self.width >> l), 3):
                raise ValueError("MIPMap level %d has shape %s" % (l, lv.shape))
        self.packed = np.concatenate([lv.reshape(-1) for lv in self.levels]).astype(np.float32)

    @staticmethod
    def _halve(a, axis):
        a = np.moveaxis(a, axis, 0).astype(np.float64)
        n = a.shape[0]
        m = max(1, n // 2)
        out = np.zeros((m,) + a.shape[1:])
        wsum = np.zeros(m)
        for k, wt in zip((-1, 0, 1, 2), (1.0, 3.0, 3.0, 1.0)):
            src = 2 * np.arange(m) + k
            ok = (src >= 0) & (src < n)
            out[ok] += wt * a[src[ok]]
            wsum[ok] += wt
        out /= wsum.reshape((m,) + (1,) * (a.ndim - 1))
        return np.moveaxis(out, 0, axis)

    @classmethod
    def downsample(cls, img):
        out = img
        if img.shape[1] > 1:
            out = cls._halve(out, 1)
        if img.shape[0] > 1:
            out = cls._halve(out, 0)
        return out.astype(np.float32)


class ImageTexture:
    """texture/image.rs:8-34 `ImageTexture<Spectrum, UVMapping>` (make_imagemap_spect, constructors.rs:295-319;
    file loading, scale and gamma are the host's, before the pyramid is built)."""
    type = A.FTN_TEXTURE_IMAGE

    def __init__(self, mipmap, mapping=None):
        self.mipmap = mipmap
        self.mapping = mapping or UVMapping()
        self.tex1 = self.tex2 = _spectrum(0.0)


_TEXTURES = (Checkerboard2DTexture, UVTexture, ImageTexture)


class ConstantTexture:
    """texture/mod.rs:34-42 as an entry of the texture table."""
    type = A.FTN_TEXTURE_CONSTANT

    def __init__(self, value):
        self.value = _spectrum(value)
        self.tex1 = self.tex2 = _spectrum(0.0)
        self.mapping = UVMapping()


class InTable:
    """Marks a Kd / Kr texture to travel through the scene's texture table (FtnMaterial::param_texture) instead of the
    material's inline slot; every other parameter's texture always goes through the table."""

    def __init__(self, texture):
        self.texture = texture


_TABLE_TEXTURES = (Checkerboard2DTexture, UVTexture, ImageTexture, ConstantTexture)


class TextureTable:
    """FtnSceneDesc::textures under construction: add() returns the 1-based id FtnMaterial::param_texture carries."""

    def __init__(self):
        self.entries, self._ids = [], {}

    def add(self, tex):
        if id(tex) in self._ids:
            return self._ids[id(tex)]
        t = A.FtnTexture()
        t.type = tex.type
        t.value[:] = getattr(tex, "value", _spectrum(0.0)).tolist()
        t.tex1[:] = tex.tex1.tolist(); t.tex2[:] = tex.tex2.tolist()
        t.uv_scale[:] = list(tex.mapping.scale); t.uv_delta[:] = list(tex.mapping.offset)
        if isinstance(tex, ImageTexture):
            mp = tex.mipmap
            t.image = mp.packed.ctypes.data_as(C.POINTER(C.c_float))
            t.image_width, t.image_height, t.image_levels, t.image_wrap = mp.width, mp.height, len(mp.levels), mp.wrap
        self.entries.append((t, tex))        # the texture object (and its pyramid) stays alive with the table
        self._ids[id(tex)] = len(self.entries)
        return len(self.entries)

    def c_array(self):
        arr = (A.FtnTexture * max(1, len(self.entries)))()
        for i, (t, _) in enumerate(self.entries):
            arr[i] = t
        return arr


def _param(m, table, param, value, set_constant):
    """One material parameter (loaders/constructors.rs get_texture_or_default): a texture goes into the table, anything
    else into the constant field."""
    if isinstance(value, InTable):
        value = value.texture
    if isinstance(value, _TABLE_TEXTURES):
        if table is None:
            raise ValueError("a textured material parameter needs the scene's texture table")
        m.param_texture[param] = table.add(value)
    else:
        set_constant(value)


def _tex_or(value, convert):
    return value if isinstance(value, _TABLE_TEXTURES + (InTable,)) else convert(value)


def _fill_kd(m, kd, table=None, param=A.FTN_PARAM_KD):
    """Kd is a constant spectrum or one of the textures above (inline slot), or InTable(texture)."""
    if isinstance(kd, InTable):
        _param(m, table, param, kd, None)
        m.uv_scale[:] = [1.0, 1.0]
        return
    if isinstance(kd, _TEXTURES):
        m.kd_texture = kd.type
        m.tex1[:] = kd.tex1.tolist(); m.tex2[:] = kd.tex2.tolist()
        m.uv_scale[:] = list(kd.mapping.scale); m.uv_delta[:] = list(kd.mapping.offset)
        if isinstance(kd, ImageTexture):   # the numpy array stays alive with the material object
            mp = kd.mipmap
            m.image = mp.packed.ctypes.data_as(C.POINTER(C.c_float))
            m.image_width, m.image_height, m.image_levels, m.image_wrap = mp.width, mp.height, len(mp.levels), mp.wrap
    else:
        m.kd_texture = A.FTN_TEXTURE_CONSTANT
        m.kd[:] = _spectrum(kd).tolist()
        m.uv_scale[:] = [1.0, 1.0]


class MatteMaterial:
    """material/matte.rs; constant Kd, sigma = 0 (Lambert).  Default Kd 0.5 (constructors.rs:193)."""
    type = A.FTN_MATERIAL_MATTE

    def __init__(self, kd=0.5, sigma=0.0):
        self.kd = kd if isinstance(kd, _TEXTURES + (InTable,)) else _spectrum(kd)
        self.sigma = _tex_or(sigma, float)   # degrees; != 0 selects Oren-Nayar (matte.rs:42-49); a float texture is allowed

    def fill(self, m, table=None):
        m.type = self.type
        _fill_kd(m, self.kd, table)
        _param(m, table, A.FTN_PARAM_SIGMA, self.sigma, lambda v: setattr(m, "sigma", v))


class MetalMaterial:
    """material/metal.rs; constant eta/k; roughness default 0.01, remap default true
    (constructors.rs:213-230)."""
    type = A.FTN_MATERIAL_METAL

    def __init__(self, eta, k, roughness=0.01, u_roughness=None, v_roughness=None, remap_roughness=True):
        self.eta = _tex_or(eta, _spectrum)
        self.k = _tex_or(k, _spectrum)
        if u_roughness is not None and v_roughness is not None:     # RoughnessTex::Anisotropic, constructors.rs:221-226
            self.u, self.v = _tex_or(u_roughness, float), _tex_or(v_roughness, float)
        else:
            self.u = self.v = _tex_or(roughness, float)
        self.remap = bool(remap_roughness)

    def fill(self, m, table=None):
        m.type = self.type
        _param(m, table, A.FTN_PARAM_ETA, self.eta, lambda v: m.eta.__setitem__(slice(None), v.tolist()))
        _param(m, table, A.FTN_PARAM_K, self.k, lambda v: m.k.__setitem__(slice(None), v.tolist()))
        _param(m, table, A.FTN_PARAM_UROUGHNESS, self.u, lambda v: setattr(m, "u_roughness", v))
        _param(m, table, A.FTN_PARAM_VROUGHNESS, self.v, lambda v: setattr(m, "v_roughness", v))
        m.remap_roughness = int(self.remap)


class PlasticMaterial:
    """material/plastic.rs; defaults Kd = Ks = 0.25, roughness 0.1 (constructors.rs:232-238)."""
    type = A.FTN_MATERIAL_PLASTIC

    def __init__(self, kd=0.25, ks=0.25, roughness=0.1, remap_roughness=True):
        self.kd = kd if isinstance(kd, _TEXTURES + (InTable,)) else _spectrum(kd)
        self.ks = _tex_or(ks, _spectrum)
        self.roughness, self.remap = _tex_or(roughness, float), bool(remap_roughness)

    def fill(self, m, table=None):
        m.type = self.type
        _fill_kd(m, self.kd, table)
        _param(m, table, A.FTN_PARAM_KS, self.ks, lambda v: m.ks.__setitem__(slice(None), v.tolist()))

        def rough(v):
            m.u_roughness = m.v_roughness = v
        _param(m, table, A.FTN_PARAM_UROUGHNESS, self.roughness, rough)
        m.remap_roughness = int(self.remap)


class MirrorMaterial:
    """material/mirror.rs; Kr default 0.9 (constructors.rs:207-210)."""
    type = A.FTN_MATERIAL_MIRROR

    def __init__(self, kr=0.9):
        self.kr = kr if isinstance(kr, _TEXTURES + (InTable,)) else _spectrum(kr)     # Kr may be textured (mirror.rs:23)

    def fill(self, m, table=None):
        m.type = self.type
        if isinstance(self.kr, _TEXTURES + (InTable,)):
            _fill_kd(m, self.kr, table, A.FTN_PARAM_KR)   # the inline slot of the ABI serves Kd (matte, plastic) or Kr (mirror)
        else:
            m.kr[:] = self.kr.tolist()


class GlassMaterial:
    """material/glass.rs; defaults of constructors.rs:198-205 (Kr 1, Kt 1, index 1.5, roughness 0, remaproughness true --
    which turns roughness 0 into a small alpha: rough glass).  Alphas of exactly 0 (FresnelSpecular, `todo!()` in the
    reference) make scene creation fail with FTN_ERR_UNSUPPORTED."""
    type = A.FTN_MATERIAL_GLASS

    def __init__(self, kr=1.0, kt=1.0, eta=1.5, u_roughness=0.0, v_roughness=0.0, remap_roughness=True):
        self.kr, self.kt, self.eta = _tex_or(kr, _spectrum), _tex_or(kt, _spectrum), _tex_or(eta, float)
        self.u_roughness, self.v_roughness, self.remap_roughness = _tex_or(u_roughness, float), _tex_or(v_roughness, float), bool(remap_roughness)

    def fill(self, m, table=None):
        m.type = self.type
        _param(m, table, A.FTN_PARAM_KR, self.kr, lambda v: m.kr.__setitem__(slice(None), v.tolist()))
        _param(m, table, A.FTN_PARAM_KT, self.kt, lambda v: m.kt.__setitem__(slice(None), v.tolist()))
        _param(m, table, A.FTN_PARAM_INDEX, self.eta, lambda v: m.eta.__setitem__(slice(None), [v] * 3))
        _param(m, table, A.FTN_PARAM_UROUGHNESS, self.u_roughness, lambda v: setattr(m, "u_roughness", v))
        _param(m, table, A.FTN_PARAM_VROUGHNESS, self.v_roughness, lambda v: setattr(m, "v_roughness", v))
        m.remap_roughness = int(self.remap_roughness)


class DiffuseAreaLight:
    """light/diffuse.rs:24-41; attached to a shape through GeometricPrimitive.light."""

    def __init__(self, emit=1.0):
        self.emit = _spectrum(emit)


class InfiniteAreaLight:
    """light/infinite.rs:23-61."""

    def __init__(self, texels, light_to_world):
        t = _f32(texels)
        assert t.ndim == 3 and t.shape[2] == 3, "texels must be (height, width, 3)"
        self.texels = t
        self.height, self.width = t.shape[0], t.shape[1]
        self.light_to_world = light_to_world

    @staticmethod
    def new_uniform(radiance, light_to_world=None):
        return InfiniteAreaLight(_spectrum(radiance).reshape(1, 1, 3), light_to_world or Transform.identity())

    @staticmethod
    def new_envmap(texels, light_to_world=None):
        return InfiniteAreaLight(texels, light_to_world or Transform.identity())


class PointLight:
    """light/point.rs:16-28.  `make_point_light` (constructors.rs:330-337) builds light_to_world as a
    translation to `from` only, so `PointLight.from_params(I, scale, from_)` ignores the CTM too."""

    def __init__(self, light_to_world, intensity):
        self.intensity = _spectrum(intensity)
        self.world_point = light_to_world.apply_points_f32(np.zeros((1, 3), np.float32))[0]

    @staticmethod
    def from_params(I=1.0, scale=1.0, from_=(0.0, 0.0, 0.0)):
        return PointLight(Transform.translate(from_), _spectrum(I) * _spectrum(scale))


class DistantLight:
    """light/distant.rs:18-31: `from_to(from, to, L)` -> direction towards the light = normalize(from - to)
    (cgmath: v * (1 / |v|), evaluated here in f32 so that every backend receives the same vector)."""

    def __init__(self, radiance, dir_to_light):
        self.radiance = _spectrum(radiance)
        d = _f32(dir_to_light).reshape(3)
        mag = np.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2], dtype=np.float32)
        self.dir_to_light = d * (np.float32(1.0) / mag)

    @staticmethod
    def from_to(from_, to, radiance):
        return DistantLight(radiance, _f32(from_).reshape(3) - _f32(to).reshape(3))

    @staticmethod
    def from_params(L=1.0, scale=1.0, from_=(0.0, 0.0, 0.0), to=(0.0, 0.0, 1.0)):   # constructors.rs:321-328
        return DistantLight.from_to(from_, to, _spectrum(L) * _spectrum(scale))


class GeometricPrimitive:
    """primitive.rs:25-29: a shape (TriangleMesh expands to one primitive per triangle, as
    PbrtSceneBuilder::shape does, loaders/pbrt.rs:275-317) with its material and area light."""

    def __init__(self, shape, material=None, light=None):
        self.shape, self.material, self.light = shape, material, light


# ---------------------------------------------------------------------------------------------
# Scene
# ---------------------------------------------------------------------------------------------

RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("t_max", np.float32), ("time", np.float32)])
HIT_DTYPE = np.dtype([("prim", np.uint32), ("t", np.float32), ("b1", np.float32), ("b2", np.float32)])
assert RAY_DTYPE.itemsize == 32 and HIT_DTYPE.itemsize == 16


def make_rays(origins, dirs, t_max=np.inf, time=0.0):
    """Ray::new (geometry/mod.rs:98-102) over arrays."""
    o = _f32(origins).reshape(-1, 3)
    d = _f32(dirs).reshape(-1, 3)
    n = max(o.shape[0], d.shape[0])
    r = np.zeros(n, dtype=RAY_DTYPE)
    r["o"] = o
    r["d"] = d
    r["t_max"] = t_max
    r["time"] = time
    return r


class Scene:
    """scene/mod.rs:14-68.  `Scene(primitives, lights)` flattens everything into the SoA
    `FtnSceneDesc`, uploads it and builds the aggregate (BVH::build, bvh.rs:27)."""

    def __init__(self, primitives, lights=(), backend=None, build=True, device=None):
        self.backend = backend or default_backend()
        if device is not None:          # the scene lives on the CUDA device that is current when it is created
            self.backend.call("set_device", int(device))
        self._handle = A.VOIDP()
        meshes = [p for p in primitives if isinstance(p.shape, TriangleMesh)]
        spheres = [p for p in primitives if isinstance(p.shape, Sphere)]
        materials = []

        def mat_id(m):
            if m is None:
                return -1
            for i, e in enumerate(materials):
                if e is m:
                    return i
            materials.append(m)
            return len(materials) - 1

        # triangles first (mesh order, tri_id), then spheres: the primitive ids of the ABI
        vbase = 0
        pos, nrm, uvs, idx, mdescs = [], [], [], [], []
        any_normals = any(p.shape.normals is not None for p in meshes)
        any_uvs = any(p.shape.tex_coords is not None for p in meshes)
        if meshes and any_normals != all(p.shape.normals is not None for p in meshes):
            raise ValueError("either every mesh carries normals or none does")
        if meshes and any_uvs != all(p.shape.tex_coords is not None for p in meshes):
            raise ValueError("either every mesh carries uvs or none does")
        first = 0
        for p in meshes:
            m = p.shape
            pos.append(m.vertices)
            if any_normals:
                nrm.append(m.normals)
            if any_uvs:
                uvs.append(m.tex_coords)
            idx.append(m.vertex_indices + np.uint32(vbase))
            vbase += m.vertices.shape[0]
            # `AreaLightSource "diffuse"` in front of the shape: every triangle carries its own DiffuseAreaLight (loaders/pbrt.rs:275-316)
            mdescs.append((first, m.n_triangles, mat_id(p.material), A.FTN_MESH_FLIP_NORMALS if m.flip_normals() else 0,
                           None if p.light is None else p.light.emit))
            first += m.n_triangles
        self.n_triangles = first
        self.n_spheres = len(spheres)
        self._positions = _f32(np.concatenate(pos)) if pos else np.zeros((0, 3), np.float32)
        self._normals = _f32(np.concatenate(nrm)) if nrm else None
        self._uvs = _f32(np.concatenate(uvs)) if uvs else None
        self._indices = np.ascontiguousarray(np.concatenate(idx), dtype=np.uint32) if idx else np.zeros((0, 3), np.uint32)

        c_meshes = (A.FtnMeshDesc * max(1, len(mdescs)))()
        for i, (f, n, mid, fl, emit) in enumerate(mdescs):
            c_meshes[i].first_tri, c_meshes[i].n_tris, c_meshes[i].material_id, c_meshes[i].flags = f, n, mid, fl
            c_meshes[i].emissive = int(emit is not None)
            if emit is not None:
                c_meshes[i].emit[:] = emit.tolist()
        c_spheres = (A.FtnSphere * max(1, len(spheres)))()
        for i, p in enumerate(spheres):
            s, cs = p.shape, c_spheres[i]
            cs.object_to_world[:] = s.object_to_world.flat().tolist()
            cs.world_to_object[:] = s.object_to_world.flat_inv().tolist()
            cs.radius, cs.z_min, cs.z_max, cs.phi_max_deg = s.radius, s.z_min, s.z_max, s.phi_max
            cs.reverse_orientation = int(s.reverse_orientation)
            cs.material_id = mat_id(p.material)
            cs.emissive = int(p.light is not None)
            if p.light is not None:
                cs.emit[:] = p.light.emit.tolist()
        c_mats = (A.FtnMaterial * max(1, len(materials)))()
        self._textures = TextureTable()
        for i, m in enumerate(materials):
            m.fill(c_mats[i], self._textures)
        c_texs = self._textures.c_array()
        self._light_texels = []
        c_lights = (A.FtnLight * max(1, len(lights)))()
        for i, l in enumerate(lights):
            cl = c_lights[i]
            ident = Transform.identity()
            if isinstance(l, InfiniteAreaLight):
                tex = _f32(l.texels)
                self._light_texels.append(tex)
                cl.type = A.FTN_LIGHT_INFINITE
                cl.texels = _ptr(tex, A.f32)
                cl.width, cl.height = l.width, l.height
                cl.light_to_world[:] = l.light_to_world.flat().tolist()
                cl.world_to_light[:] = l.light_to_world.flat_inv().tolist()
            elif isinstance(l, PointLight):
                cl.type = A.FTN_LIGHT_POINT
                cl.light_to_world[:] = ident.flat().tolist(); cl.world_to_light[:] = ident.flat().tolist()
                cl.point[:] = l.world_point.tolist()
                cl.intensity[:] = l.intensity.tolist()
            elif isinstance(l, DistantLight):
                cl.type = A.FTN_LIGHT_DISTANT
                cl.light_to_world[:] = ident.flat().tolist(); cl.world_to_light[:] = ident.flat().tolist()
                cl.direction[:] = l.dir_to_light.tolist()
                cl.intensity[:] = l.radiance.tolist()
            else:
                raise NotImplementedError("lights listed explicitly must be infinite, point or distant")

        d = A.FtnSceneDesc()
        d.abi_version = A.FTN_ABI_VERSION
        d.positions = _ptr(self._positions, A.f32)
        d.normals = _ptr(self._normals, A.f32) if self._normals is not None else None
        d.uvs = _ptr(self._uvs, A.f32) if self._uvs is not None else None
        d.n_vertices = self._positions.shape[0]
        d.indices = _ptr(self._indices, A.u32)
        d.n_triangles = self.n_triangles
        d.meshes, d.n_meshes = c_meshes, len(mdescs)
        d.spheres, d.n_spheres = c_spheres, len(spheres)
        d.materials, d.n_materials = c_mats, len(materials)
        d.lights, d.n_lights = c_lights, len(lights)
        d.textures, d.n_textures = c_texs, len(self._textures.entries)
        self.backend.call("scene_create", C.byref(d), C.byref(self._handle))
        self.upload_bytes = int(self._positions.nbytes + self._indices.nbytes
                                + (self._normals.nbytes if self._normals is not None else 0)
                                + (self._uvs.nbytes if self._uvs is not None else 0)
                                + sum(t.nbytes for t in self._light_texels))
        if build:
            self.build()

    def build(self):
        """BVH::build (bvh.rs:27) + Scene::new's light preprocessing (scene/mod.rs:32-49)."""
        self.backend.call("bvh_build", self._handle)

    def close(self):
        if self._handle:
            self.backend.call("scene_destroy", self._handle)
            self._handle = A.VOIDP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._handle

    def world_bound(self):
        """Scene::world_bound (scene/mod.rs:66): (min xyz, max xyz)."""
        out = (A.f32 * 6)()
        self.backend.call("scene_world_bound", self._handle, out)
        a = np.array(out[:], dtype=np.float32)
        return a[:3], a[3:]

    def stats(self):
        st = A.FtnStats()
        self.backend.call("scene_stats", self._handle, C.byref(st))
        return st.as_dict()

    def morton_codes_and_order(self):
        codes = np.zeros(self.n_triangles, np.uint32)
        order = np.zeros(self.n_triangles, np.uint32)
        self.backend.call("bvh_debug_morton", self._handle, _ptr(codes, A.u32), _ptr(order, A.u32))
        return codes, order

    def intersect(self, rays):
        """Scene::intersect (scene/mod.rs:51) over a batch -> HIT_DTYPE array; prim == FTN_NO_HIT on miss."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        self.backend.call("intersect", self._handle, rays.shape[0],
                          rays.ctypes.data_as(C.POINTER(A.FtnRay)), hits.ctypes.data_as(C.POINTER(A.FtnHit)))
        return hits

    def intersect_test(self, rays):
        """Scene::intersect_test (scene/mod.rs:55) over a batch -> bool array."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        out = np.zeros(rays.shape[0], dtype=np.uint8)
        self.backend.call("intersect_test", self._handle, rays.shape[0],
                          rays.ctypes.data_as(C.POINTER(A.FtnRay)), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)


# ---------------------------------------------------------------------------------------------
# Sensor
# ---------------------------------------------------------------------------------------------

class PerspectiveCamera:
    """camera/mod.rs:85-114 (CameraProjection::new :50-72 inlined)."""

    def __init__(self, camera_to_world, full_resolution, screen_window=None, shutter_interval=(0.0, 1.0),
                 lens_radius=0.0, focal_dist=1e6, fov=90.0):
        xres, yres = int(full_resolution[0]), int(full_resolution[1])
        if screen_window is None:       # PbrtHeader::make_camera, loaders/pbrt.rs:440-451
            aspect = np.float32(xres) / np.float32(yres)
            if aspect > 1.0:
                screen_window = ((-aspect, -1.0), (aspect, 1.0))
            else:
                screen_window = ((-1.0, -1.0 / aspect), (1.0, 1.0 / aspect))
        (x0, y0), (x1, y1) = screen_window
        camera_to_screen = Transform.perspective(fov, 1.0e-2, 1000.0)
        screen_to_raster = (Transform.scale(xres, yres, 1.0)
                            * Transform.scale(1.0 / (x1 - x0), 1.0 / (y0 - y1), 1.0)
                            * Transform.translate((-x0, -y1, 0.0)))
        raster_to_screen = screen_to_raster.inverse()
        self.raster_to_camera = camera_to_screen.inverse() * raster_to_screen
        self.camera_to_world = camera_to_world
        self.shutter_interval = (float(shutter_interval[0]), float(shutter_interval[1]))
        self.lens_radius = float(lens_radius)
        self.focal_dist = float(focal_dist)
        self.full_resolution = (xres, yres)

    def to_abi(self):
        c = A.FtnCamera()
        c.camera_to_world[:] = self.camera_to_world.flat().tolist()
        c.raster_to_camera[:] = self.raster_to_camera.flat().tolist()
        c.lens_radius, c.focal_distance = self.lens_radius, self.focal_dist
        c.shutter_open, c.shutter_close = self.shutter_interval
        return c


class BoxFilter:
    """filter/mod.rs:10-32."""

    def __init__(self, radius=(0.5, 0.5)):
        self.radius = (float(radius[0]), float(radius[1]))


class _PinnedPool:
    """Page-locked host buffers (ftn_host_alloc) recycled by size: a Film's pixel buffer is written by the
    device every render, so it lives in pinned memory when the backend offers it -- and cudaHostAlloc is far
    too slow to pay per Film."""

    def __init__(self):
        self.free = {}     # (backend id, bytes) -> [pointer]

    def take(self, backend, nbytes):
        key = (id(backend), nbytes)
        if self.free.get(key):
            return self.free[key].pop()
        p = A.VOIDP()
        backend.call("host_alloc", nbytes, C.byref(p))
        return p.value

    def give(self, backend, nbytes, ptr):
        self.free.setdefault((id(backend), nbytes), []).append(ptr)


_PINNED = _PinnedPool()


class Film:
    """film.rs:18-81.  After a render `pixels` is an (h, w, 4) float32 array of XYZ sums and
    filter-weight sums -- `Film.pixels` (film.rs:24)."""

    def __init__(self, resolution, crop_window=((0.0, 0.0), (1.0, 1.0)), filter=None, diagonal=35.0,
                 backend=None):
        self.full_resolution = (int(resolution[0]), int(resolution[1]))
        (cx0, cy0), (cx1, cy1) = crop_window
        self.crop_window = (float(cx0), float(cx1), float(cy0), float(cy1))   # pbrt `cropwindow` order
        self.filter = filter or BoxFilter()
        self.diagonal = diagonal
        self.backend = backend or default_backend()
        w, h = A.i32(), A.i32()
        self.backend.call("film_pixel_count", C.byref(self.to_abi()), C.byref(w), C.byref(h))
        self.width, self.height = w.value, h.value
        self._pinned = None
        nbytes = self.height * self.width * 16
        if nbytes and self.backend.has("host_alloc"):
            self._pinned = _PINNED.take(self.backend, nbytes)
            buf = (C.c_float * (self.height * self.width * 4)).from_address(self._pinned)
            # the buffer goes back to the pool when the LAST array over it dies (every numpy view keeps `buf` alive
            # through its base chain), not when the Film does: a caller may keep film.pixels beyond the Film
            weakref.finalize(buf, _PINNED.give, self.backend, nbytes, self._pinned)
            self._pixels = np.frombuffer(buf, dtype=np.float32).reshape(self.height, self.width, 4)
            self._stale = True     # a recycled buffer: zeroed on first read unless a render overwrote it first
        else:
            self._pixels = np.zeros((self.height, self.width, 4), dtype=np.float32)
            self._stale = False

    @property
    def pixels(self):
        if self._stale:
            self._pixels[...] = 0.0
            self._stale = False
        return self._pixels

    @pixels.setter
    def pixels(self, value):
        self._pixels = value
        self._stale = False

    def to_abi(self):
        f = A.FtnFilm()
        f.x_resolution, f.y_resolution = self.full_resolution
        f.crop_window[:] = list(self.crop_window)
        f.filter_radius[:] = list(self.filter.radius)
        return f

    def into_spectrum_buffer(self):
        """film.rs:195-210: (rgb (h*w, 3), (w, h)); XYZ -> RGB, times 1/weight, clamped at 0."""
        xyz = self.pixels[..., :3].reshape(-1, 3)
        wsum = self.pixels[..., 3].reshape(-1)
        m = np.array([[3.240479, -1.537150, -0.498535], [-0.969256, 1.875991, 0.041556],
                      [0.055648, -0.204043, 1.057311]], dtype=np.float32)
        rgb = np.empty_like(xyz)
        for i in range(3):
            rgb[:, i] = (m[i, 0] * xyz[:, 0] + m[i, 1] * xyz[:, 1]) + m[i, 2] * xyz[:, 2]
        nz = wsum != 0
        inv = np.float32(1.0) / wsum[nz]
        rgb[nz] = np.maximum(np.float32(0.0), rgb[nz] * inv[:, None])
        return rgb, (self.width, self.height)


class RandomSampler:
    """sampler/random.rs:6-21.  `mode` selects the stream: the GPU's counter stream (default) or
    the reference's sequential per-tile xoshiro256+ stream (CPU oracle only)."""

    def __init__(self, samples_per_pixel, seed=0, mode=A.FTN_SAMPLER_COUNTER):
        self.samples_per_pixel, self.seed, self.mode = int(samples_per_pixel), int(seed), int(mode)

    @staticmethod
    def new_with_seed(samples_per_pixel, seed, mode=A.FTN_SAMPLER_COUNTER):
        return RandomSampler(samples_per_pixel, seed, mode)

    def to_abi(self, sample_begin=0, sample_stride=1):
        s = A.FtnSampler()
        s.samples_per_pixel, s.seed, s.mode = self.samples_per_pixel, self.seed, self.mode
        s.sample_begin, s.sample_stride = sample_begin, sample_stride
        return s


class PathIntegrator:
    """integrator/path.rs:10-20."""

    def __init__(self, max_depth, rr_threshold):
        self.max_depth, self.rr_threshold = int(max_depth), float(rr_threshold)

    def to_abi(self):
        i = A.FtnIntegrator()
        i.type, i.max_depth, i.rr_threshold = A.FTN_INTEGRATOR_PATH, self.max_depth, self.rr_threshold
        return i


class DirectLightingIntegrator:
    """integrator/direct_lighting.rs:20-25, LightStrategy::UniformSampleOne only
    (UniformSampleAll is unimplemented!() in the reference, :108-116)."""

    def __init__(self, max_depth, strategy="one"):
        if strategy != "one":
            raise NotImplementedError("uniform_sample_all_lights is unimplemented!() in the reference")
        self.max_depth = int(max_depth)

    def to_abi(self):
        i = A.FtnIntegrator()
        i.type, i.max_depth, i.rr_threshold = A.FTN_INTEGRATOR_DIRECT_LIGHTING, self.max_depth, 0.0
        return i


class SamplerIntegrator:
    """integrator/mod.rs:22-25, 218-227."""

    def __init__(self, camera, radiance):
        self.camera, self.radiance = camera, radiance
        self.last_stats = None

    def render_parallel(self, scene, film, sampler, sample_begin=0, sample_stride=1):
        """Renders into film.pixels.  Raises FountainError(FTN_ERR_NAN_RADIANCE) where the
        reference panics in check_radiance (integrator/mod.rs:285)."""
        st = A.FtnStats()
        out = film._pixels         # ftn_render overwrites every pixel of the cropped bounds: no need to clear a recycled buffer first
        cam, f, s, it = self.camera.to_abi(), film.to_abi(), sampler.to_abi(sample_begin, sample_stride), self.radiance.to_abi()
        scene.backend.call("render", scene.handle, C.byref(cam), C.byref(f), C.byref(s), C.byref(it),
                           out.ctypes.data_as(C.POINTER(A.FtnPixel)), C.byref(st))
        film._stale = False        # only after success: a failed render leaves the buffer to be cleared on first read
        self.last_stats = st.as_dict()
        return self.last_stats

    render = render_parallel   # integrator/mod.rs:206 (sequential variant: same result here)

    def render_multi(self, scenes, film, sampler, sample_begin=0, sample_stride=1):
        """One process, several GPUs (ftn_render_multi): `scenes` holds the same scene built on each device
        (`Scene(..., device=i)`); the samples are sharded by index, the partial films are summed with one NCCL
        reduce onto scenes[0]'s device and read back into film.pixels."""
        if len(scenes) > 1:
            # The library binds NCCL at run time by soname (libnccl.so.2).  A Python process that will also import torch
            # must load torch's bundled (newer) copy FIRST, or torch's own import later fails on the system copy's
            # missing symbols; a host without torch (the Rust binary) simply gets the system library.
            try:
                import torch  # noqa: F401
            except ImportError:
                pass
        st = A.FtnStats()
        out = film._pixels
        cam, f, s, it = self.camera.to_abi(), film.to_abi(), sampler.to_abi(sample_begin, sample_stride), self.radiance.to_abi()
        handles = (A.VOIDP * len(scenes))(*[sc.handle.value for sc in scenes])
        scenes[0].backend.call("render_multi", handles, len(scenes), C.byref(cam), C.byref(f), C.byref(s), C.byref(it),
                               out.ctypes.data_as(C.POINTER(A.FtnPixel)), C.byref(st))
        film._stale = False
        self.last_stats = st.as_dict()
        return self.last_stats
