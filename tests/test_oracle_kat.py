"""Pins the CPU oracle against every known-answer / property test the reference's own
test-suite holds for the hot path.  file:line = akofke/fountain.

CPU only.  The reference is Rust and cannot be run here, so these KATs are how the oracle
is anchored; what no reference test pins (RNG stream, cgmath ulps) is listed in DESIGN.md.
"""
import ctypes as C
import math

import numpy as np
import pytest

from fountain_b200 import _abi as A
from fountain_b200 import api
from fountain_b200.transform import Transform
from tests.conftest import unit_sphere_dirs

f32 = np.float32


def _arr3(v):
    return (A.f32 * 3)(*[float(x) for x in v])


def _ray(o, d, t_max=float("inf")):
    r = A.FtnRay()
    r.o[:] = [float(x) for x in o]
    r.d[:] = [float(x) for x in d]
    r.t_max, r.time = t_max, 0.0
    return r


# ---- src/morton.rs:43-60 ------------------------------------------------------------------
def test_morton_kat(oracle):
    lib = oracle.library()
    assert lib.orc_kat_morton3(0.9999, 0.0, 0.9999) == 0b00_101101101101101101101101101101
    assert lib.orc_kat_expand_bits(0x3FF) == 0b00_001001001001001001001001001001
    assert lib.orc_kat_to_fixed_point(0.99999) == 0x3FF
    assert lib.orc_kat_to_fixed_point(0.0) == 0


def test_morton_matches_bit_interleave(oracle):
    lib = oracle.library()
    rng = np.random.default_rng(0)
    for x, y, z in rng.random((200, 3)).astype(f32):
        fx, fy, fz = (int(np.trunc(f32(v) * f32(1024.0))) for v in (x, y, z))
        ref = 0
        for b in range(10):
            ref |= ((fx >> b) & 1) << (3 * b + 2) | ((fy >> b) & 1) << (3 * b + 1) | ((fz >> b) & 1) << (3 * b)
        assert lib.orc_kat_morton3(float(x), float(y), float(z)) == ref


# ---- src/shapes/triangle.rs:441-450 -----------------------------------------------------------
def test_sign_differs_kat(oracle):
    sd = oracle.library().orc_kat_sign_differs
    assert sd(1.0, 2.0, -1.0) == 1
    assert sd(1.0, 2.0, 1.0) == 0
    assert sd(-1.0, -2.0, 1.0) == 1
    assert sd(-1.0, -2.0, -1.0) == 0
    assert sd(-1.0, 2.0, -1.0) == 1
    assert sd(-1.0, 2.0, 1.0) == 1
    assert sd(0.0, 0.0, 0.0) == 0
    assert sd(0.0, 0.0, -0.0) == 1


# ---- src/geometry/bounds.rs:292-323 -----------------------------------------------------------
@pytest.mark.parametrize("bmin,bmax,o,d,expected", [
    ((1, 1, 1), (2, 2, 2), (0, 0, 0), (1, 1, 1), (1.0, 2.0)),
    ((-.5, -.5, -.5), (.5, .5, .5), (0, 0, -2), (0, 0, 1), (1.5, 2.5)),
    ((1, 1, 1), (2, 2, 2), (0, 0, 0), (-1, 1, 1), None),
    ((1, 1, 1), (2, 2, 2), (1, 1, 1), (1, 0, 0), (0.0, 1.0)),
])
def test_bounds3f_intersect_kat(oracle, bmin, bmax, o, d, expected):
    out = (A.f32 * 2)()
    r = _ray(o, d)
    hit = oracle.library().orc_kat_bounds_intersect(_arr3(bmin), _arr3(bmax), C.byref(r), out)
    if expected is None:
        assert hit == 0
    else:
        assert hit == 1
        assert abs(out[0] - expected[0]) < 1e-3 and abs(out[1] - expected[1]) < 1e-3


# ---- src/err_float.rs:5-30 ---------------------------------------------------------------------
def test_gamma_and_next_float(oracle):
    lib = oracle.library()
    eps = f32(np.finfo(np.float32).eps) * f32(0.5)
    for n in (2, 3, 5, 6, 7):
        nf = f32(n)
        assert f32(lib.orc_kat_gamma(n)) == f32((nf * eps) / (f32(1.0) - nf * eps))
    for v in (0.0, -0.0, 1.0, -1.0, 1e-30, -3.5e7, float(np.finfo(np.float32).tiny)):
        assert f32(lib.orc_kat_next_float_up(v)) == np.nextafter(f32(v), f32(np.inf))
        if v != 0.0:
            assert f32(lib.orc_kat_next_float_down(v)) == np.nextafter(f32(v), f32(-np.inf))
    # Reference quirk kept on purpose: next_float_down maps 0.0 to -0.0 and then tests
    # `v >= 0.0` (true for -0.0), so it decrements 0x8000_0000 to 0x7FFF_FFFF = NaN
    # (err_float.rs:22-30; pbrt tests `v > 0`).  It feeds EFloat bounds of the sphere test.
    assert math.isnan(lib.orc_kat_next_float_down(0.0)) and math.isnan(lib.orc_kat_next_float_down(-0.0))
    assert lib.orc_kat_next_float_up(float("inf")) == float("inf")
    assert lib.orc_kat_next_float_down(float("-inf")) == float("-inf")


# ---- src/fresnel.rs:110-116: exact f32 equality --------------------------------------------------
def test_fresnel_dielectric_kat(oracle):
    got = f32(oracle.library().orc_kat_fresnel_dielectric(0.087642014, 1.0, 1.5))
    assert got == f32(0.611180067)


# ---- src/sampling.rs:188-208 -------------------------------------------------------------------
def test_distribution_1d_kat(oracle):
    func = (A.f32 * 4)(0.0, 0.0, 1.0, 0.0)
    for u in (0.0, 0.1, 0.5, 0.9):
        x, pdf, idx = A.f32(), A.f32(), C.c_int()
        oracle.library().orc_kat_distribution1d_sample(func, 4, u, C.byref(x), C.byref(pdf), C.byref(idx))
        assert idx.value == 2
        assert pdf.value == 4.0
        assert 0.5 <= x.value < 0.75


def test_concentric_sample_disk(oracle):
    rng = np.random.default_rng(1)
    out = (A.f32 * 2)()
    for u0, u1 in rng.random((100, 2)):
        oracle.library().orc_kat_concentric_sample_disk(u0, u1, out)
        assert math.hypot(out[0], out[1]) <= 1.0 + 1e-6
    oracle.library().orc_kat_concentric_sample_disk(0.5, 0.5, out)
    assert (out[0], out[1]) == (0.0, 0.0)


# ---- src/camera/mod.rs:220-366 ------------------------------------------------------------------
def test_camera_look_at_and_fov(oracle):
    cam = api.PerspectiveCamera(Transform.camera_look_at((0, 0, -1), (0, 0, 0), (0, 1, 0)), (100, 100),
                                screen_window=((-1, -1), (1, 1)), fov=90.0)
    abi = cam.to_abi()
    r = A.FtnRay()
    lib = oracle.library()
    lib.orc_kat_camera_ray(C.byref(abi), 50.0, 50.0, 0.5, 0.5, 0.0, C.byref(r))
    assert abs(r.d[0]) < 1e-5 and abs(r.d[1]) < 1e-5 and r.d[2] > 0.99999      # forward is +z
    assert abs(r.o[2] + 1.0) < 1e-4
    # measured FOV == 90 deg +- 0.01 on both axes (camera/mod.rs:340-365)
    for axis in (0, 1):
        a, b = [50.0, 50.0], [50.0, 50.0]
        a[axis], b[axis] = 0.0, 100.0
        ra, rb = A.FtnRay(), A.FtnRay()
        lib.orc_kat_camera_ray(C.byref(abi), a[0], a[1], 0.5, 0.5, 0.0, C.byref(ra))
        lib.orc_kat_camera_ray(C.byref(abi), b[0], b[1], 0.5, 0.5, 0.0, C.byref(rb))
        da, db = np.array(ra.d[:]), np.array(rb.d[:])
        ang = math.degrees(math.acos(float(np.dot(da, db) / (np.linalg.norm(da) * np.linalg.norm(db)))))
        assert abs(ang - 90.0) < 0.01


# ---- src/shapes/sphere.rs:241-273 ---------------------------------------------------------------
def test_whole_sphere_intersect(oracle):
    lib = oracle.library()
    s = A.FtnSphere()
    ident = Transform.identity()
    s.object_to_world[:] = ident.flat().tolist()
    s.world_to_object[:] = ident.flat_inv().tolist()
    s.radius, s.z_min, s.z_max, s.phi_max_deg = 1.0, -1.0, 1.0, 360.0
    out = (A.f32 * 13)()
    rng = np.random.default_rng(4)
    orig = np.array([3.0, 3.0, 3.0])
    n = 0
    while n < 100:
        p = rng.uniform(-1, 1, 3)
        if p @ p >= 1.0:
            continue
        n += 1
        r = _ray(orig, p - orig)
        assert lib.orc_kat_sphere_intersect(C.byref(s), C.byref(r), out) == 1
        assert max(abs(out[4]), abs(out[5]), abs(out[6])) < 1e-4      # p_err
    r = _ray((1, 0, -2), (0, 0, 2))
    assert lib.orc_kat_sphere_intersect(C.byref(s), C.byref(r), out) == 1
    r = _ray((1, 0, -2), (0.0001, 0, 2))
    assert lib.orc_kat_sphere_intersect(C.byref(s), C.byref(r), out) == 0


# ---- src/bvh.rs:401-444: BVH == brute-force list ----------------------------------------------
def _random_spheres_scene(backend, n=100, seed=3):
    rng = np.random.default_rng(seed)
    prims = []
    for _ in range(n):
        o2w = Transform.translate(rng.uniform(-10, 10, 3).astype(np.float32))
        prims.append(api.GeometricPrimitive(api.Sphere(o2w, radius=float(rng.uniform(0.5, 3.0)))))
    return api.Scene(prims, [], backend=backend)


def test_bvh_intersect_many_nodes(oracle, orc_backend):
    scene = _random_spheres_scene(orc_backend)
    dirs = unit_sphere_dirs(500, 33)
    rays = api.make_rays(np.zeros((500, 3)), dirs)
    hits = scene.intersect(rays)
    anyhit = scene.intersect_test(rays)
    brute = np.zeros(500, dtype=api.HIT_DTYPE)
    oracle.library().orc_intersect_brute(scene.handle, 500, rays.ctypes.data_as(C.POINTER(A.FtnRay)),
                                         brute.ctypes.data_as(C.POINTER(A.FtnHit)))
    assert np.array_equal(anyhit, hits["prim"] != A.FTN_NO_HIT)
    assert np.array_equal(hits["prim"], brute["prim"])
    assert np.array_equal(hits["t"], brute["t"])
    assert (hits["prim"] != A.FTN_NO_HIT).sum() > 50


# ---- tests/tri_watertight.rs --------------------------------------------------------------------
@pytest.fixture(scope="module")
def cube_scene_oracle(orc_backend, rounded_cube_path):
    mesh = api.TriangleMesh.from_ply(rounded_cube_path)
    return api.Scene([api.GeometricPrimitive(mesh)], [], backend=orc_backend)


def test_rounded_cube_facts(cube_scene_oracle):
    # SURVEY section 2 "Assets": 8664 verts, 4332 faces, bounds +-9.986
    assert cube_scene_oracle.n_triangles == 4332
    lo, hi = cube_scene_oracle.world_bound()
    assert np.allclose(lo, -9.986, atol=2e-3) and np.allclose(hi, 9.986, atol=2e-3)


def test_watertight_rounded_cube_oracle(cube_scene_oracle):
    dirs = unit_sphere_dirs(100_000, 7)
    rays = api.make_rays(np.zeros((dirs.shape[0], 3)), dirs)
    assert cube_scene_oracle.intersect_test(rays).all()
    hits = cube_scene_oracle.intersect(rays)
    assert (hits["prim"] != A.FTN_NO_HIT).all()
    assert np.isfinite(hits["t"]).all() and (hits["t"] > 0).all()


def test_oracle_bvh_equals_brute_force_on_mesh(oracle, cube_scene_oracle):
    rng = np.random.default_rng(11)
    n = 2000
    o = rng.uniform(-30, 30, (n, 3)).astype(np.float32)
    tgt = rng.uniform(-9, 9, (n, 3)).astype(np.float32)
    rays = api.make_rays(o, tgt - o)
    hits = cube_scene_oracle.intersect(rays)
    brute = np.zeros(n, dtype=api.HIT_DTYPE)
    oracle.library().orc_intersect_brute(cube_scene_oracle.handle, n, rays.ctypes.data_as(C.POINTER(A.FtnRay)),
                                         brute.ctypes.data_as(C.POINTER(A.FtnHit)))
    assert np.array_equal(hits["prim"] == A.FTN_NO_HIT, brute["prim"] == A.FTN_NO_HIT)
    m = hits["prim"] != A.FTN_NO_HIT
    # exact-t ties on shared edges may pick either triangle (triangle.rs:237-238 accepts t == t_max)
    differ = m & (hits["prim"] != brute["prim"])
    assert np.all(np.abs(hits["t"][differ] - brute["t"][differ]) <= 4 * np.spacing(np.abs(brute["t"][differ])))
    assert np.array_equal(hits["t"][m & ~differ], brute["t"][m & ~differ])


# ---- Morton ordering of the LBVH key (the GPU's bit-exact target) -------------------------------
def test_oracle_morton_order_is_stable_sort(cube_scene_oracle):
    codes, order = cube_scene_oracle.morton_codes_and_order()
    assert codes.max() < (1 << 30)
    assert np.array_equal(order, np.argsort(codes, kind="stable").astype(np.uint32))


# ---- RNG restatements (parity UNPINNED in the reference; self-consistency only) -------------------
def test_reference_stream_is_xoshiro256plus(oracle):
    out = (A.f32 * 8)()
    oracle.library().orc_kat_reference_stream(0, 8, out)
    m = (1 << 64) - 1

    def splitmix(x):
        x = (x + 0x9E3779B97F4A7C15) & m
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        return x, z ^ (z >> 31)

    s, x = [], 0
    for _ in range(4):
        x, z = splitmix(x)
        s.append(z)
    # first SplitMix64 outputs for seed 0 are published test vectors of the algorithm
    assert s[0] == 0xE220A8397B1DCDAF and s[1] == 0x6E789E6AA1B965F4
    exp = []
    for _ in range(8):
        res = (s[0] + s[3]) & m
        t = (s[1] << 17) & m
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]
        s[2] ^= t
        s[3] = ((s[3] << 45) | (s[3] >> 19)) & m
        exp.append(f32((res >> 32) >> 8) * f32(2.0 ** -24))
    assert [f32(v) for v in out[:]] == exp
    assert all(0.0 <= v < 1.0 for v in out[:])


def test_counter_stream_uniformity(oracle):
    lib = oracle.library()
    v = np.array([lib.orc_kat_counter_uniform(0, i, d) for i in range(4000) for d in range(4)])
    assert 0.0 <= v.min() and v.max() < 1.0
    assert abs(v.mean() - 0.5) < 0.01
    h, _ = np.histogram(v, bins=16, range=(0, 1))
    assert h.min() > 0.85 * len(v) / 16
    # different dimensions / samples decorrelate
    a = v.reshape(-1, 4)
    assert abs(np.corrcoef(a[:, 0], a[:, 1])[0, 1]) < 0.05


# ---- mipmap.rs: MIPMap lookups (the image texture's device half) ---------------------------------------------------
def _mip_lookup(oracle, mp, s, t, width):
    out = (C.c_float * 3)()
    rc = oracle.library().orc_kat_mipmap_lookup(mp.packed.ctypes.data_as(C.POINTER(C.c_float)), mp.width, mp.height, len(mp.levels),
                                               mp.wrap, float(s), float(t), float(width), out)
    assert rc == 0
    return np.array(list(out), np.float32)


def test_mipmap_lookup_reference_kat(oracle):
    """mipmap.rs:364-382 test_mipmap_lookup: a constant 16x15 image filters to the constant (6 ulps) at every
    coordinate and width, including 0."""
    from fountain_b200 import api
    val = np.float32(0.5)
    mp = api.MIPMap(np.full((15, 16, 3), val, np.float32), "repeat")
    assert [l.shape[:2] for l in mp.levels] == [(15, 16), (7, 8), (3, 4), (1, 2), (1, 1)]     # :107-121
    widths = list(np.logspace(-4.0, 0.0, 10)) + [0.0]
    for s in np.linspace(0.0, 1.0, 25):
        for t in np.linspace(0.0, 1.0, 25):
            for w in widths:
                got = _mip_lookup(oracle, mp, s, t, w)
                assert np.all(np.abs(got.view(np.int32) - val.view(np.int32)) <= 6), (s, t, w, got)


def _np_texel(mp, level, s, t):
    lv = mp.levels[level]; h, w = lv.shape[:2]
    if mp.wrap == 0: s, t = s % w, t % h
    elif mp.wrap == 2: s, t = min(max(s, 0), w - 1), min(max(t, 0), h - 1)
    elif s < 0 or s >= w or t < 0 or t >= h: return np.zeros(3)
    return lv[t, s].astype(np.float64)


def _np_triangle(mp, level, s, t):
    lv = mp.levels[level]; h, w = lv.shape[:2]
    x, y = s * w - 0.5, t * h - 0.5
    s0, t0 = int(np.floor(x)), int(np.floor(y)); ds, dt = x - s0, y - t0
    return ((1 - ds) * (1 - dt) * _np_texel(mp, level, s0, t0) + (1 - ds) * dt * _np_texel(mp, level, s0, t0 + 1)
            + ds * (1 - dt) * _np_texel(mp, level, s0 + 1, t0) + ds * dt * _np_texel(mp, level, s0 + 1, t0 + 1))


def _np_lookup(mp, s, t, width):
    n = len(mp.levels)
    level = n - 1 + np.log2(max(width, 1e-8))
    if level < 0: return _np_triangle(mp, 0, s, t)
    if level >= n - 1: return _np_texel(mp, n - 1, 0, 0)
    lf = int(np.floor(level)); d = level - lf
    return (1 - d) * _np_triangle(mp, lf, s, t) + d * _np_triangle(mp, lf + 1, s, t)


@pytest.mark.parametrize("wrap", ["repeat", "black", "clamp"])
def test_mipmap_lookup_closed_forms(oracle, wrap):
    """No reference test reads distinct texels (the image one is #[ignore]d and needs a file): pinned against an
    independent numpy statement of mipmap.rs:245-311."""
    from fountain_b200 import api
    rng = np.random.default_rng(11)
    mp = api.MIPMap(rng.random((6, 8, 3)).astype(np.float32), wrap)
    assert len(mp.levels) == 4
    # texel centres at width 0 return the texel itself
    for (si, ti) in ((0, 0), (7, 5), (3, 2)):
        assert np.allclose(_mip_lookup(oracle, mp, (si + 0.5) / 8, (ti + 0.5) / 6, 0.0), mp.levels[0][ti, si], rtol=1e-6)
    for _ in range(300):
        s, t = rng.uniform(-0.3, 1.3, 2)
        w = float(rng.choice([0.0, 1e-3, 0.13, 0.2, 0.26, 0.4, 0.51, 0.9, 1.0, 3.0]))
        assert np.allclose(_mip_lookup(oracle, mp, s, t, w), _np_lookup(mp, np.float32(s), np.float32(t), w), rtol=2e-4, atol=2e-6), (s, t, w)
