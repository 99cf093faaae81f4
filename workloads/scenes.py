"""Builders for the workloads BASELINE.json names.  Pure host-side
scene description (numpy) -- the same arrays feed the CUDA library and, in tests, the oracle.
"""
import functools
import math
import os

import numpy as np

from fountain_b200 import api
from fountain_b200.transform import Transform

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROUNDED_CUBE_PLY = os.path.join(_ROOT, "data", "rounded_cube.ply")


# ---- C1: testscenes/furnace_empty.pbrt --------------------------------------------------------
def furnace_scene(backend=None):
    """Radius-100 sphere, matte Kd .5, diffuse area light L 1, ReverseOrientation; camera at
    (0,-2,0) looking at the origin, up +z, fov 60, 16x16, box filter (furnace_empty.pbrt:2-19)."""
    sphere = api.Sphere(Transform.identity(), reverse_orientation=True, radius=100.0)
    prim = api.GeometricPrimitive(sphere, api.MatteMaterial(0.5), api.DiffuseAreaLight(1.0))
    scene = api.Scene([prim], [], backend=backend)
    cam_to_world = Transform.look_at((0, -2, 0), (0, 0, 0), (0, 0, 1)).inverse()   # pbrt.rs:430
    camera = api.PerspectiveCamera(cam_to_world, (16, 16), fov=60.0)
    film = api.Film((16, 16), backend=backend)
    return scene, camera, film


# ---- C2: data/rounded_cube.ply, Lambert, uniform env ---------------------------------------------
def rounded_cube_scene(backend=None, resolution=(512, 512), material=None, light=None, ply=None):
    """rounded_cube.ply at the identity, `matte` default Kd .5, `LightSource "infinite" L 1`;
    camera LookAt 0 -40 0 -> 0 0 0 up 0 0 1, fov 40, no lens."""
    mesh = api.TriangleMesh.from_ply(ply or ROUNDED_CUBE_PLY)
    prim = api.GeometricPrimitive(mesh, material or api.MatteMaterial(0.5))
    lights = [light or api.InfiniteAreaLight.new_uniform(1.0)]
    scene = api.Scene([prim], lights, backend=backend)
    cam_to_world = Transform.look_at((0, -40, 0), (0, 0, 0), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=40.0)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


# ---- C3 / C5: synthetic tessellated sphere with seeded radial displacement ----------------------
def _hash01(ix, iy, seed):
    """Stateless integer hash -> [0,1) float64 (lowbias32 finaliser)."""
    x = (ix.astype(np.uint64) * np.uint64(0x9E3779B1) + iy.astype(np.uint64) * np.uint64(0x85EBCA77)
         + np.uint64(seed) * np.uint64(0xC2B2AE3D)) & np.uint64(0xFFFFFFFF)
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x.astype(np.uint64) * np.uint64(0x7FEB352D) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x.astype(np.uint64) * np.uint64(0x846CA68B) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x.astype(np.float64) / 4294967296.0


@functools.lru_cache(maxsize=4)
def displaced_sphere_mesh(n_lon, n_lat, radius=10.0, amplitude=0.05, seed=1, with_normals=True):
    """Lat-long grid of n_lon x n_lat quads x 2 triangles (1000 x 500 -> 1M tris).
    Radius is scaled by 1 + amplitude * smooth value noise so the BVH is not trivial.  Poles are
    pinched (degenerate quads collapse to triangles that the watertight test rejects cleanly)."""
    lon = np.arange(n_lon + 1)
    lat = np.arange(n_lat + 1)
    LON, LAT = np.meshgrid(lon % n_lon, lat, indexing="xy")           # wrap the seam
    phi = 2.0 * math.pi * (LON / n_lon)
    theta = math.pi * (LAT / n_lat)
    # value noise on a coarse 64 x 32 lattice, bilinear, periodic in longitude
    gx, gy = 64, 32
    fx = LON / n_lon * gx
    fy = LAT / n_lat * gy
    x0 = np.floor(fx).astype(np.int64); y0 = np.floor(fy).astype(np.int64)
    tx = fx - x0; ty = fy - y0
    tx = tx * tx * (3 - 2 * tx); ty = ty * ty * (3 - 2 * ty)

    def g(ix, iy):
        return _hash01(np.mod(ix, gx), np.clip(iy, 0, gy), seed)

    noise = ((g(x0, y0) * (1 - tx) + g(x0 + 1, y0) * tx) * (1 - ty)
             + (g(x0, y0 + 1) * (1 - tx) + g(x0 + 1, y0 + 1) * tx) * ty)
    fine = _hash01(LON.astype(np.int64), LAT.astype(np.int64), seed + 17)
    pole_fade = np.sin(theta) ** 2                                     # single point at each pole
    r = radius * (1.0 + amplitude * pole_fade * (noise - 0.5) * 2.0 + 0.002 * pole_fade * (fine - 0.5))
    st = np.sin(theta)
    P = np.stack([r * st * np.cos(phi), r * st * np.sin(phi), r * np.cos(theta)], axis=-1)
    verts = P.reshape(-1, 3).astype(np.float32)
    W = n_lon + 1
    i, j = np.meshgrid(np.arange(n_lon), np.arange(n_lat), indexing="xy")
    a = (j * W + i).reshape(-1); b = a + 1; c = a + W; d = c + 1
    tris = np.empty((a.size * 2, 3), dtype=np.uint32)
    tris[0::2] = np.stack([a, c, b], axis=1)       # outward-facing (counter-clockwise from outside)
    tris[1::2] = np.stack([b, c, d], axis=1)
    normals = None
    if with_normals:
        n = verts / np.maximum(np.linalg.norm(verts, axis=1, keepdims=True), 1e-20)
        normals = n.astype(np.float32)
    return verts, tris, normals


def synthetic_mesh_scene(n_lon=1000, n_lat=500, backend=None, material=None, light=None, resolution=(2048, 2048),
                         seed=1, build=True):
    """C3 (1000 x 500 -> 1M triangles) / C5 (5000 x 5000 -> 50M): Lambert + uniform env."""
    v, t, n = displaced_sphere_mesh(n_lon, n_lat, seed=seed)
    mesh = api.TriangleMesh(Transform.identity(), t, v, n)
    prim = api.GeometricPrimitive(mesh, material or api.MatteMaterial(0.5))
    scene = api.Scene([prim], [light or api.InfiniteAreaLight.new_uniform(1.0)], backend=backend, build=build)
    cam_to_world = Transform.look_at((0, -40, 0), (0, 0, 0), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=40.0)
    return scene, camera


def primary_ray_batch(camera, resolution, oracle_lib=None):
    """Batch A of C3: one pinhole ray through every pixel centre (numpy restatement of
    camera/mod.rs:117-143 for lens_radius == 0; tests check it against the oracle's camera)."""
    w, h = resolution
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32) + np.float32(0.5),
                         np.arange(h, dtype=np.float32) + np.float32(0.5), indexing="xy")
    r2c = camera.raster_to_camera.m
    px = np.stack([xs.reshape(-1), ys.reshape(-1), np.zeros(w * h), np.ones(w * h)], axis=0).astype(np.float64)
    pc = r2c @ px
    pc = pc[:3] / pc[3]
    d = pc / np.linalg.norm(pc, axis=0, keepdims=True)
    c2w = camera.camera_to_world.m
    dw = (c2w[:3, :3] @ d).T
    o = np.broadcast_to(c2w[:3, 3], dw.shape)
    return api.make_rays(o.astype(np.float32), dw.astype(np.float32))


def diffuse_bounce_batch(rays, hits, positions, indices, seed=2):
    """Batch B of C3: for every batch-A hit, one cosine-hemisphere direction about the geometric
    normal (facing the incoming ray) from a seeded generator; origin nudged off the surface."""
    m = hits["prim"] != 0xFFFFFFFF
    hr, hh = rays[m], hits[m]
    tri = indices[hh["prim"]]
    p0, p1, p2 = positions[tri[:, 0]].astype(np.float64), positions[tri[:, 1]].astype(np.float64), positions[tri[:, 2]].astype(np.float64)
    n = np.cross(p1 - p0, p2 - p0)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
    d_in = hr["d"].astype(np.float64)
    n = np.where((np.sum(n * d_in, axis=1) > 0)[:, None], -n, n)
    p = hr["o"].astype(np.float64) + d_in * hh["t"][:, None].astype(np.float64)
    rng = np.random.default_rng(seed)
    u = rng.random((p.shape[0], 2))
    r = np.sqrt(u[:, 0]); ph = 2 * math.pi * u[:, 1]
    lx, ly, lz = r * np.cos(ph), r * np.sin(ph), np.sqrt(np.maximum(0.0, 1 - u[:, 0]))
    a = np.where((np.abs(n[:, 0]) > 0.9)[:, None], np.array([0.0, 1.0, 0.0]), np.array([1.0, 0.0, 0.0]))
    t = np.cross(n, a); t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(n, t)
    d = t * lx[:, None] + b * ly[:, None] + n * lz[:, None]
    o = p + n * 1e-3
    return api.make_rays(o.astype(np.float32), d.astype(np.float32))


# ---- C4: "Rust-logo-style": gear-ring mesh, TR metal, thin lens, image env -----------------------
@functools.lru_cache(maxsize=4)
def gear_ring_mesh(n_teeth=48, seg_per_tooth=16, n_height=24, n_radial=24, r_in=4.0, r_out=7.0, tooth=0.8,
                   height=1.5):
    """Extruded gear ring (outer wall with teeth, inner wall, top and bottom annuli), uniformly
    tessellated; default ~ 200 k triangles with the default arguments scaled by `detail`."""
    n_ang = n_teeth * seg_per_tooth
    ang = np.linspace(0.0, 2.0 * math.pi, n_ang, endpoint=False)
    prof = r_out + tooth * np.clip(np.sin(ang * n_teeth) * 3.0, -1.0, 1.0) * 0.5

    verts, tris = [], []

    def add_grid(P, flip):
        # P: (rows, n_ang, 3) closed in the angular direction
        rows = P.shape[0]
        base = sum(v.shape[0] for v in verts)
        verts.append(P.reshape(-1, 3))
        i, j = np.meshgrid(np.arange(n_ang), np.arange(rows - 1), indexing="xy")
        a = (j * n_ang + i).reshape(-1); b = (j * n_ang + (i + 1) % n_ang).reshape(-1)
        c = a + n_ang; d = b + n_ang
        t = np.empty((a.size * 2, 3), dtype=np.int64)
        if flip:
            t[0::2] = np.stack([a, c, b], axis=1); t[1::2] = np.stack([b, c, d], axis=1)
        else:
            t[0::2] = np.stack([a, b, c], axis=1); t[1::2] = np.stack([b, d, c], axis=1)
        tris.append(t + base)

    z = np.linspace(0.0, height, n_height + 1)
    ca, sa = np.cos(ang), np.sin(ang)
    outer = np.stack([np.broadcast_to(prof * ca, (z.size, n_ang)), np.broadcast_to(prof * sa, (z.size, n_ang)),
                      np.broadcast_to(z[:, None], (z.size, n_ang))], axis=-1)
    add_grid(outer, flip=False)
    inner = np.stack([np.broadcast_to(r_in * ca, (z.size, n_ang)), np.broadcast_to(r_in * sa, (z.size, n_ang)),
                      np.broadcast_to(z[:, None], (z.size, n_ang))], axis=-1)
    add_grid(inner, flip=True)
    s = np.linspace(0.0, 1.0, n_radial + 1)[:, None]
    rad = r_in + (prof[None, :] - r_in) * s
    top = np.stack([rad * ca, rad * sa, np.full_like(rad, height)], axis=-1)
    add_grid(top, flip=False)
    bot = np.stack([rad * ca, rad * sa, np.zeros_like(rad)], axis=-1)
    add_grid(bot, flip=True)
    return np.concatenate(verts).astype(np.float32), np.concatenate(tris).astype(np.uint32)


@functools.lru_cache(maxsize=4)
def sky_sun_envmap(width=2048, height=1024, seed=3, peak=1.0e4):
    """Procedural lat-long sky + sun map (the reference's HDR is not in the repository)."""
    rng = np.random.default_rng(seed)
    t = (np.arange(height) + 0.5) / height * math.pi
    p = (np.arange(width) + 0.5) / width * 2 * math.pi
    T, P = np.meshgrid(t, p, indexing="ij")
    d = np.stack([np.sin(T) * np.cos(P), np.sin(T) * np.sin(P), np.cos(T)], axis=-1)
    up = np.clip(d[..., 2], 0, 1)
    sky = np.stack([0.25 + 0.35 * (1 - up), 0.35 + 0.35 * (1 - up), 0.6 + 0.3 * (1 - up)], axis=-1)
    ground = np.array([0.18, 0.16, 0.14])
    img = np.where((d[..., 2] > 0)[..., None], sky, ground)
    sun_dir = np.array([math.cos(1.0) * math.cos(0.7), math.cos(1.0) * math.sin(0.7), math.sin(1.0)])
    cosang = np.clip(d @ sun_dir, -1, 1)
    img = img + peak * np.exp(-((np.arccos(cosang) / 0.02) ** 2))[..., None] * np.array([1.0, 0.9, 0.75])
    img = img * (1.0 + 0.05 * rng.random(img.shape[:2])[..., None])
    return img.astype(np.float32)


def logo_style_scene(backend=None, resolution=(1920, 1080), detail=1.0, env_size=(2048, 1024)):
    """C4: gear ring on a ground quad, Cu metal roughness 0.01 (remapped), thin lens focused on
    the ring, sky+sun environment."""
    k = max(1, int(round(detail * 4)))
    v, t = gear_ring_mesh(seg_per_tooth=4 * k, n_height=6 * k, n_radial=6 * k)
    ring = api.TriangleMesh(Transform.translate((0, 0, 0.01)), t, v)
    g = 40.0
    gv = np.array([[-g, -g, 0], [g, -g, 0], [g, g, 0], [-g, g, 0]], dtype=np.float32)
    gt = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    ground = api.TriangleMesh(Transform.identity(), gt, gv)
    copper = api.MetalMaterial(eta=(0.2, 0.92, 1.1), k=(3.9, 2.45, 2.14), roughness=0.01, remap_roughness=True)
    prims = [api.GeometricPrimitive(ring, copper), api.GeometricPrimitive(ground, api.MatteMaterial(0.5))]
    env = api.InfiniteAreaLight.new_envmap(sky_sun_envmap(env_size[0], env_size[1]))
    scene = api.Scene(prims, [env], backend=backend)
    eye = np.array([0.0, -18.0, 9.0]); look = np.array([0.0, 0.0, 0.75])
    cam_to_world = Transform.look_at(eye, look, (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=40.0, lens_radius=0.1,
                                   focal_dist=float(np.linalg.norm(look - eye)))
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


# ---- delta lights: a lit floor with an occluder ----------------------------------
def delta_lights_scene(backend=None, resolution=(48, 48), lights=("point", "distant"), occluder=True, material=None, fov=50.0):
    """A 10x10 floor quad at z = 0 (normal +z), an optional occluding quad at z = 1 over x in [1, 3],
    a point light at (0, 0, 4) and/or a distant light arriving from direction (1, 0, 2); camera at
    (0, -9, 7) looking at the origin.  No environment: unlit pixels are exactly black."""
    def quad(x0, x1, y0, y1, z):
        v = np.array([[x0, y0, z], [x1, y0, z], [x1, y1, z], [x0, y1, z]], np.float32)
        return api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v)
    mat = material or api.MatteMaterial((0.6, 0.5, 0.4))
    prims = [api.GeometricPrimitive(quad(-5, 5, -5, 5, 0.0), mat)]
    if occluder:
        prims.append(api.GeometricPrimitive(quad(1, 3, -1, 1, 1.0), mat))
    ls = []
    for name in lights:
        if name == "point":
            ls.append(api.PointLight.from_params(I=(30.0, 28.0, 26.0), from_=(0.0, 0.0, 4.0)))
        elif name == "distant":
            ls.append(api.DistantLight.from_params(L=(1.5, 1.6, 1.7), from_=(1.0, 0.0, 2.0), to=(0.0, 0.0, 0.0)))
        else:
            raise ValueError(name)
    scene = api.Scene(prims, ls, backend=backend)
    cam_to_world = Transform.look_at((0, -9, 7), (0, 0, 0), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=fov)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


# ---- mirror (material/mirror.rs) -------------------------------------------------------
def mirror_scene(backend=None, resolution=(48, 48), env=1.0, with_floor=True):
    """A mirror quad (Kr .8) standing at y = 2 facing the camera, a matte floor at z = -1 and the rounded
    cube's material-less twin replaced by a small matte box made of two quads; uniform environment plus a
    point light so that the reflection shows lit diffuse surfaces."""
    def quad(p0, p1, p2, p3):
        v = np.array([p0, p1, p2, p3], np.float32)
        return api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v)
    prims = [api.GeometricPrimitive(quad((-3, 2, -1), (3, 2, -1), (3, 2, 3), (-3, 2, 3)), api.MirrorMaterial((0.8, 0.7, 0.6)))]
    if with_floor:
        prims.append(api.GeometricPrimitive(quad((-6, -8, -1), (6, -8, -1), (6, 4, -1), (-6, 4, -1)), api.MatteMaterial((0.6, 0.5, 0.4))))
        prims.append(api.GeometricPrimitive(quad((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1)), api.MatteMaterial((0.2, 0.6, 0.3))))
    lights = [api.InfiniteAreaLight.new_uniform(env)]
    if with_floor:
        lights.append(api.PointLight.from_params(I=40.0, from_=(0.0, -2.0, 4.0)))
    scene = api.Scene(prims, lights, backend=backend)
    cam_to_world = Transform.look_at((0, -7, 1.5), (0, 2, 0.5), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=45.0)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


# ---- textured Kd (texture/checkerboard.rs, texture/uv.rs through UVMapping) ----------------
def textured_floor_scene(backend=None, resolution=(48, 48), texture="checkerboard", material="matte", look_at=None, fov=50.0):
    """A 12x12 floor whose uv are its world (x, y), so texture cells are unit squares aligned with the
    integer grid; lit by a distant light from straight above (+ a weak uniform environment unless a
    narrow-fov probe camera `look_at` = (x, y) is asked for, which keeps the result closed-form)."""
    v = np.array([[-6, -6, 0], [6, -6, 0], [6, 6, 0], [-6, 6, 0]], np.float32)
    mesh = api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v, tex_coords=v[:, :2].copy())
    if texture == "checkerboard":
        tex = api.Checkerboard2DTexture((0.8, 0.2, 0.2), (0.1, 0.1, 0.9), api.UVMapping(1.0, 1.0, 0.0, 0.0))
    elif texture == "checkerboard_scaled":
        tex = api.Checkerboard2DTexture((0.8, 0.2, 0.2), (0.1, 0.1, 0.9), api.UVMapping(2.0, 0.5, 0.25, -0.5))
    elif texture == "uv":
        tex = api.UVTexture(api.UVMapping(0.5, 0.25, 0.1, 0.2))
    else:
        tex = texture   # a texture object, e.g. api.ImageTexture
    mat = api.MatteMaterial(tex) if material == "matte" else api.PlasticMaterial(tex, 0.2, 0.2)
    lights = [api.DistantLight.from_params(L=3.0, from_=(0.0, 0.0, 1.0), to=(0.0, 0.0, 0.0))]
    if look_at is None:
        lights.append(api.InfiniteAreaLight.new_uniform(0.2))
    scene = api.Scene([api.GeometricPrimitive(mesh, mat)], lights, backend=backend)
    if look_at is None:
        cam_to_world = Transform.look_at((0, -9, 6), (0, 0, 0), (0, 0, 1)).inverse()
    else:
        cam_to_world = Transform.look_at((look_at[0], look_at[1], 30.0), (look_at[0], look_at[1], 0.0), (0, 1, 0)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=fov)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


# ---- image-textured Kd (texture/image.rs + mipmap.rs) ------------------------------------------
@functools.lru_cache(maxsize=4)
def procedural_image(width=64, height=48, seed=5):
    """A deterministic colourful image with fine and coarse detail (fine stripes over smooth blobs), so that
    different pyramid levels differ visibly."""
    rng = np.random.default_rng(seed)
    s = (np.arange(width)[None, :] + 0.5) / width
    t = (np.arange(height)[:, None] + 0.5) / height
    img = np.zeros((height, width, 3), np.float32)
    img[..., 0] = 0.5 + 0.4 * np.sin(2 * np.pi * 3 * s) * np.cos(2 * np.pi * 2 * t)
    img[..., 1] = 0.15 + 0.7 * ((np.arange(width)[None, :] + np.arange(height)[:, None]) % 2)      # texel-sized checker
    img[..., 2] = 0.1 + 0.8 * t * np.ones_like(s)
    img += (0.05 * rng.random(img.shape)).astype(np.float32)
    return np.clip(img, 0.0, 1.0).astype(np.float32)


def image_texture_scene(backend=None, resolution=(48, 48), wrap="repeat", material="matte", uscale=0.25, lens_radius=0.0):
    """The 12x12 floor of textured_floor_scene seen at a grazing angle (the footprint, hence the mip level, varies
    from the foreground to the horizon) plus a sphere carrying the same image through its own (phi, theta) uv;
    a distant light and a uniform environment, so bounce hits look textures up too."""
    tex = api.ImageTexture(api.MIPMap(procedural_image(), wrap), api.UVMapping(uscale, uscale * 1.5, 0.1, 0.2))
    stex = api.ImageTexture(api.MIPMap(procedural_image(), wrap), api.UVMapping(2.0, 1.0, 0.0, 0.0))
    make = (lambda t: api.MatteMaterial(t)) if material == "matte" else ((lambda t: api.MatteMaterial(t, sigma=30.0)) if material == "oren_nayar"
                                                                         else (lambda t: api.PlasticMaterial(t, 0.2, 0.2)))
    v = np.array([[-6, -6, 0], [6, -6, 0], [6, 6, 0], [-6, 6, 0]], np.float32)
    mesh = api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v, tex_coords=v[:, :2].copy())
    sphere = api.Sphere(Transform.translate((1.5, 1.0, 1.2)), radius=1.2)
    prims = [api.GeometricPrimitive(mesh, make(tex)), api.GeometricPrimitive(sphere, make(stex))]
    lights = [api.DistantLight.from_params(L=2.5, from_=(0.3, -0.4, 1.0), to=(0.0, 0.0, 0.0)), api.InfiniteAreaLight.new_uniform(0.3)]
    scene = api.Scene(prims, lights, backend=backend)
    cam_to_world = Transform.look_at((0, -8.5, 1.6), (0, 1, 0.4), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=50.0, lens_radius=lens_radius, focal_dist=9.0)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


def textured_mirror_probe(backend=None, xz=(0.5, 0.5), texture=None, resolution=(5, 5)):
    """A mirror quad in the plane y = 2 whose uv are its world (x, z), Kr through `texture`, under a uniform
    environment of radiance 1; a narrow camera looks at (x, 2, z) head-on, so every pixel reads Kr(x, z)."""
    v = np.array([[-6, 2, -6], [6, 2, -6], [6, 2, 6], [-6, 2, 6]], np.float32)
    mesh = api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v, tex_coords=v[:, [0, 2]].copy())
    tex = texture if texture is not None else api.Checkerboard2DTexture((0.8, 0.7, 0.6), (0.3, 0.2, 0.1), api.UVMapping(1.0, 1.0, 0.0, 0.0))
    scene = api.Scene([api.GeometricPrimitive(mesh, api.MirrorMaterial(tex))], [api.InfiniteAreaLight.new_uniform(1.0)], backend=backend)
    camera = api.PerspectiveCamera(Transform.look_at((xz[0], -28.0, xz[1]), (xz[0], 2.0, xz[1]), (0, 0, 1)).inverse(), resolution, fov=0.5)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


# ---- emissive triangle mesh: a quad light over a small Cornell-style set (light/diffuse.rs on triangles) ----
def quad_light_scene(backend=None, resolution=(48, 48), light_order=(0, 1), emit=(17.0, 12.0, 4.0), two_lights=False, box=True):
    """Floor, back wall and a small box of matte quads lit ONLY by a quad light (two emissive triangles, no material
    hit from below: matte black) facing down at z = 3; `light_order` permutes the two triangles of the light inside
    its mesh (the reference lists area lights in its BVH's order, the ABI in primitive order: the parity tests pick
    the permutation that makes both enumerations agree).  `two_lights` adds a second, smaller emissive quad."""
    def quad(p0, p1, p2, p3, order=(0, 1)):
        v = np.array([p0, p1, p2, p3], np.float32)
        tris = [[0, 1, 2], [0, 2, 3]]
        return api.TriangleMesh(Transform.identity(), np.array([tris[order[0]], tris[order[1]]], np.uint32).reshape(-1), v)
    white, red = api.MatteMaterial((0.7, 0.7, 0.7)), api.MatteMaterial((0.6, 0.1, 0.1))
    prims = [api.GeometricPrimitive(quad((-3, -3, 0), (3, -3, 0), (3, 3, 0), (-3, 3, 0)), white),          # floor, normal +z
             api.GeometricPrimitive(quad((-3, 3, 0), (3, 3, 0), (3, 3, 3.5), (-3, 3, 3.5)), red)]            # back wall
    if box:
        prims += [api.GeometricPrimitive(quad((-1, -1, 1), (0.5, -1, 1), (0.5, 0.5, 1), (-1, 0.5, 1)), white),   # box top
                  api.GeometricPrimitive(quad((-1, -1, 0), (0.5, -1, 0), (0.5, -1, 1), (-1, -1, 1)), white)]     # box front
    # the light: wound so that its geometric normal (p1 - p0) x (p2 - p0) points DOWN
    prims.append(api.GeometricPrimitive(quad((-1, -1, 3), (-1, 1, 3), (1, 1, 3), (1, -1, 3), light_order),
                                        api.MatteMaterial((0.0, 0.0, 0.0)), api.DiffuseAreaLight(emit)))
    if two_lights:
        prims.append(api.GeometricPrimitive(quad((2.0, -2, 1.0), (2.0, -1, 1.0), (2.0, -1, 2.0), (2.0, -2, 2.0), light_order),
                                            api.MatteMaterial((0.0, 0.0, 0.0)), api.DiffuseAreaLight((2.0, 6.0, 9.0))))
    scene = api.Scene(prims, [], backend=backend)
    cam_to_world = Transform.look_at((0, -8, 2.2), (0, 0, 1.2), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=45.0)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


def quad_light_probe(backend=None, side=1.0, height=2.0, emit=5.0, kd=0.5, resolution=(3, 3)):
    """Closed-form check of a triangle-mesh area light: a square Lambertian emitter of side `side` facing down at
    `height` over the origin of a large matte floor; a narrow camera off to the side looks at the origin.  Radiance
    leaving the floor there = Kd / pi * E with E = 4 L (a/s) atan(a/s), a = side / 2, s = sqrt(a^2 + height^2)."""
    def quad(p0, p1, p2, p3):
        v = np.array([p0, p1, p2, p3], np.float32)
        return api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v)
    a = 0.5 * side
    prims = [api.GeometricPrimitive(quad((-50, -50, 0), (50, -50, 0), (50, 50, 0), (-50, 50, 0)), api.MatteMaterial(kd)),
             api.GeometricPrimitive(quad((-a, -a, height), (-a, a, height), (a, a, height), (a, -a, height)),
                                    api.MatteMaterial(0.0), api.DiffuseAreaLight(emit))]
    scene = api.Scene(prims, [], backend=backend)
    cam_to_world = Transform.look_at((4, 0, 1), (0, 0, 0), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=0.5)
    film = api.Film(resolution, backend=backend)
    s_ = float(np.sqrt(a * a + height * height))
    expected = kd / np.pi * 4.0 * emit * (a / s_) * np.arctan(a / s_)
    return scene, camera, film, expected


# ---- rough glass (material/glass.rs: MicrofacetReflection + MicrofacetTransmission) ----------------------
def rough_glass_scene(backend=None, resolution=(48, 48), roughness=0.25, remap=True, eta=1.5):
    """A rough-glass pane (two quads, front and back face, 0.3 apart) standing in front of a checkerboard floor and a red
    wall, lit by a uniform environment and a point light behind the pane: paths refract in, scatter inside, refract out."""
    def quad(p0, p1, p2, p3):
        v = np.array([p0, p1, p2, p3], np.float32)
        return api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v, tex_coords=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32))
    glass = api.GlassMaterial(kr=(1.0, 0.95, 0.9), kt=(0.9, 1.0, 0.95), eta=eta, u_roughness=roughness, v_roughness=roughness, remap_roughness=remap)
    floor = api.MatteMaterial(api.Checkerboard2DTexture((0.7, 0.7, 0.7), (0.15, 0.15, 0.4), api.UVMapping(1.0, 1.0, 0.0, 0.0)))
    fv = np.array([(-6, -8, -1), (6, -8, -1), (6, 6, -1), (-6, 6, -1)], np.float32)
    fmesh = api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), fv, tex_coords=fv[:, :2].copy())
    prims = [api.GeometricPrimitive(quad((-2, 0.0, -1), (2, 0.0, -1), (2, 0.0, 2.5), (-2, 0.0, 2.5)), glass),       # front face (normal -y)
             api.GeometricPrimitive(quad((2, 0.3, -1), (-2, 0.3, -1), (-2, 0.3, 2.5), (2, 0.3, 2.5)), glass),       # back face (normal +y)
             api.GeometricPrimitive(fmesh, floor),
             api.GeometricPrimitive(quad((-6, 6, -1), (6, 6, -1), (6, 6, 5), (-6, 6, 5)), api.MatteMaterial((0.6, 0.1, 0.1)))]
    lights = [api.InfiniteAreaLight.new_uniform(0.4), api.PointLight.from_params(I=60.0, from_=(1.0, 3.0, 3.0))]
    scene = api.Scene(prims, lights, backend=backend)
    cam_to_world = Transform.look_at((0.5, -7, 1.5), (0, 0.5, 0.6), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=45.0)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


# ---- the texture table: every material parameter may be textured (loaders/constructors.rs:192-238; missing in round 1) ----
def textured_params_scene(backend=None, resolution=(48, 48), variant="table"):
    """Five quads in a row on a dark floor, each with a material whose NON-Kd parameters are textured through the scene's
    texture table: plastic (Ks + roughness checkerboards, Kd image through the table), metal (eta / k checkerboards,
    anisotropic roughness textures), matte (sigma checkerboard 0 / 35 degrees: Lambert and Oren-Nayar cells), rough glass
    (Kt checkerboard, roughness checkerboard, index texture) and a mirror (Kr uv texture through the table).
    variant = "constant" builds the same scene with ConstantTextures of each checkerboard's first value in the table
    and plain constants -- what "table" must reduce to where only tex1 cells are seen; "inline" puts Kd / Kr textures in the
    materials' inline slot instead of the table (must equal "table" bit for bit)."""
    cb = lambda a, b, su=2.0, sv=2.0: api.Checkerboard2DTexture(a, b, api.UVMapping(su, sv, 0.0, 0.0))
    img = api.ImageTexture(api.MIPMap(procedural_image(32, 16, 7), "repeat"), api.UVMapping(1.0, 1.0, 0.0, 0.0))
    uvt = api.UVTexture(api.UVMapping(1.0, 1.0, 0.0, 0.0))
    wrap = (lambda t: t) if variant == "inline" else api.InTable
    if variant == "constant":
        c = api.ConstantTexture
        mats = [api.PlasticMaterial(0.3, c((0.5, 0.4, 0.3)), c(0.05)),
                api.MetalMaterial(c((0.2, 0.92, 1.1)), c((3.9, 2.45, 2.14)), u_roughness=c(0.02), v_roughness=0.1),
                api.MatteMaterial(0.6, c(35.0)),
                api.GlassMaterial(1.0, c((0.9, 0.9, 1.0)), c(1.5), c(0.2), 0.2),
                api.MirrorMaterial(0.8)]
    else:
        mats = [api.PlasticMaterial(wrap(img), cb((0.5, 0.4, 0.3), (0.05, 0.05, 0.05)), cb(0.05, 0.4, 3.0, 1.0)),
                api.MetalMaterial(cb((0.2, 0.92, 1.1), (1.5, 0.9, 0.3)), cb((3.9, 2.45, 2.14), (2.0, 2.0, 2.0)),
                                  u_roughness=cb(0.02, 0.3), v_roughness=cb(0.1, 0.05, 1.0, 4.0)),
                api.MatteMaterial(0.6, cb(35.0, 0.0)),
                api.GlassMaterial(1.0, cb((0.9, 0.9, 1.0), (0.2, 0.9, 0.3)), cb(1.5, 1.2, 1.0, 1.0), cb(0.2, 0.05), 0.2),
                api.MirrorMaterial(wrap(uvt))]
    uv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)

    def quad(p0, p1, p2, p3):
        return api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), np.array([p0, p1, p2, p3], np.float32), tex_coords=uv)
    prims = [api.GeometricPrimitive(quad((-8, -6, -0.01), (8, -6, -0.01), (8, 8, -0.01), (-8, 8, -0.01)), api.MatteMaterial(0.05))]
    for i, m in enumerate(mats):
        x0 = -5.0 + 2.05 * i
        if isinstance(m, api.GlassMaterial):      # a standing pane so that light passes through it
            prims.append(api.GeometricPrimitive(quad((x0, 0.5, 0.0), (x0 + 2, 0.5, 0.0), (x0 + 2, 0.5, 2.0), (x0, 0.5, 2.0)), m))
        else:
            prims.append(api.GeometricPrimitive(quad((x0, -1, 0.0), (x0 + 2, -1, 0.0), (x0 + 2, 1, 0.6), (x0, 1, 0.6)), m))
    lights = [api.InfiniteAreaLight.new_uniform(0.5), api.PointLight.from_params(I=80.0, from_=(0.0, -3.0, 4.0))]
    scene = api.Scene(prims, lights, backend=backend)
    cam_to_world = Transform.look_at((0, -9, 4.5), (0, 0, 0.3), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=50.0)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


# ---- ray differentials behind a mirror (specular_reflect, integrator/mod.rs:59-83) -------------------------------
def mirror_ceiling_probe(backend=None, texture=None, xy=(0.3, -0.2), h_cam=10.0, h_ceiling=20.0, resolution=(5, 5), fov=0.5):
    """A narrow camera looks straight down at a small flat mirror (Kr 1) in the plane z = 0; the mirror shows a textured
    matte ceiling at z = h_ceiling > h_cam (uv = world (x, y)) lit by a distant light from below at 60 degrees (its shadow rays pass
    beside the mirror).  Unfolded, the camera sees the ceiling from h_cam + h_ceiling away: with the mirrored differentials
    the texture footprint is that of the unfolded path; without them it would be a point."""
    def quad(z, half, flip):
        v = np.array([[-half, -half, z], [half, -half, z], [half, half, z], [-half, half, z]], np.float32)
        idx = np.array([0, 2, 1, 0, 3, 2] if flip else [0, 1, 2, 0, 2, 3], np.uint32)
        return api.TriangleMesh(Transform.identity(), idx, v, tex_coords=v[:, :2].copy())
    prims = [api.GeometricPrimitive(quad(0.0, 1.0, False), api.MirrorMaterial((1.0, 1.0, 1.0))),
             api.GeometricPrimitive(quad(h_ceiling, 30.0, True), api.MatteMaterial(texture))]
    lights = [api.DistantLight.from_params(L=3.0, from_=(np.sqrt(3.0), 0.0, -1.0), to=(0.0, 0.0, 0.0))]
    scene = api.Scene(prims, lights, backend=backend)
    cam_to_world = Transform.look_at((xy[0], xy[1], h_cam), (xy[0], xy[1], 0.0), (0, 1, 0)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=fov)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film


def mirrored_image_texture_scene(backend=None, resolution=(48, 48), wrap="repeat"):
    """The image-textured floor of image_texture_scene seen directly and in two mirrors: a quad whose vertex normals are
    splayed (a convex mirror: dndu / dndv of triangle.rs:357-375 are not zero) and a sphere (sphere.rs:160-178), so the
    texture is filtered with footprints that the reflections widened."""
    tex = api.ImageTexture(api.MIPMap(procedural_image(), wrap), api.UVMapping(0.25, 0.375, 0.1, 0.2))
    v = np.array([[-6, -6, 0], [6, -6, 0], [6, 6, 0], [-6, 6, 0]], np.float32)
    up = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    floor = api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), v, normals=up, tex_coords=v[:, :2].copy())
    mv = np.array([[-3.5, 3, 0.05], [0.5, 3, 0.05], [0.5, 3, 3.5], [-3.5, 3, 3.5]], np.float32)
    mn = np.array([[-0.35, -1, -0.3], [0.35, -1, -0.3], [0.35, -1, 0.3], [-0.35, -1, 0.3]], np.float32)
    mn /= np.linalg.norm(mn, axis=1, keepdims=True)
    mirror = api.TriangleMesh(Transform.identity(), np.array([0, 1, 2, 0, 2, 3], np.uint32), mv, normals=mn, tex_coords=mv[:, [0, 2]].copy())
    ball = api.Sphere(Transform.translate((2.6, 0.5, 1.1)), radius=1.1)
    prims = [api.GeometricPrimitive(floor, api.MatteMaterial(tex)),
             api.GeometricPrimitive(mirror, api.MirrorMaterial((0.9, 0.9, 0.9))),
             api.GeometricPrimitive(ball, api.MirrorMaterial((0.8, 0.85, 0.9)))]
    lights = [api.DistantLight.from_params(L=2.5, from_=(0.3, -0.4, 1.0), to=(0.0, 0.0, 0.0)), api.InfiniteAreaLight.new_uniform(0.3)]
    scene = api.Scene(prims, lights, backend=backend)
    cam_to_world = Transform.look_at((0, -8.5, 2.2), (0, 1, 0.9), (0, 0, 1)).inverse()
    camera = api.PerspectiveCamera(cam_to_world, resolution, fov=50.0)
    film = api.Film(resolution, backend=backend)
    return scene, camera, film
