"""CPU tier: the FTN_HD device functions of the CUDA library (compiled for the host by
tests/hostsim) against the oracle.  Same checks as tests/test_gpu_parity.py runs on the B200."""
import ctypes as C

import numpy as np
import pytest

from fountain_b200 import _abi as A
from fountain_b200 import api
from workloads import scenes
from fountain_b200.transform import Transform
from tests import parity


@pytest.fixture(scope="module", params=["bvh8", "bvh2"], autouse=True)
def bvh_layout(request):
    """Every test of this module runs on both node layouts of the aggregate: BVH8q (compressed 8-wide, the default) and
    BVH2x64 -- results must not depend on the layout (FTN_BVH_LAYOUT is read by ftn_bvh_build)."""
    import os
    old = os.environ.get("FTN_BVH_LAYOUT")
    os.environ["FTN_BVH_LAYOUT"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("FTN_BVH_LAYOUT", None)
    else:
        os.environ["FTN_BVH_LAYOUT"] = old


@pytest.fixture(scope="module")
def sim():
    from tests.hostsim import sim as m
    return m


@pytest.fixture(scope="module")
def sim_backend(sim):
    return sim.backend()


@pytest.fixture(scope="module")
def cubes(sim_backend, orc_backend, rounded_cube_path, bvh_layout):
    return parity.cube_scenes(sim_backend, orc_backend, rounded_cube_path)


def test_morton_and_order_bit_exact(cubes):
    assert parity.check_morton(*cubes) == 4332


def test_world_bound_matches(cubes):
    (lo_a, hi_a), (lo_b, hi_b) = cubes[0].world_bound(), cubes[1].world_bound()
    assert np.array_equal(lo_a, lo_b) and np.array_equal(hi_a, hi_b)


def test_ray_batch_parity_cube(cubes):
    st = parity.check_ray_batch(cubes[0], cubes[1], parity.random_ray_batch(20000, 5), "cube")
    assert st["hits"] > 5000


def test_watertight_cube(cubes):
    hits = parity.check_watertight(cubes[0], n=100_000)
    ref = cubes[1].intersect(api.make_rays(np.zeros((100_000, 3)), parity.unit_sphere_dirs(100_000, 7)))
    parity.compare_hits(hits, ref, "watertight")


def test_ray_batch_parity_displaced_sphere(sim_backend, orc_backend):
    v, t, n = scenes.displaced_sphere_mesh(96, 48)
    mesh = api.TriangleMesh(Transform.identity(), t, v, n)
    a = api.Scene([api.GeometricPrimitive(mesh)], [], backend=sim_backend)
    b = api.Scene([api.GeometricPrimitive(mesh)], [], backend=orc_backend)
    parity.check_morton(a, b)
    rays = parity.random_ray_batch(20000, 9, extent=10.0, far=40.0)
    parity.check_ray_batch(a, b, rays, "displaced sphere")


@pytest.mark.parametrize("n_tris", [1, 2, 3, 4, 5, 7, 9])
def test_tiny_scenes(sim_backend, orc_backend, n_tris):
    rng = np.random.default_rng(n_tris)
    v = rng.uniform(-1, 1, (3 * n_tris, 3)).astype(np.float32)
    t = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    mesh = api.TriangleMesh(Transform.identity(), t, v)
    a = api.Scene([api.GeometricPrimitive(mesh)], [], backend=sim_backend)
    b = api.Scene([api.GeometricPrimitive(mesh)], [], backend=orc_backend)
    parity.check_morton(a, b)
    rays = parity.random_ray_batch(3000, 100 + n_tris, extent=1.0, far=4.0)
    parity.compare_hits(a.intersect(rays), b.intersect(rays), "tiny")
    assert np.array_equal(a.intersect_test(rays), b.intersect_test(rays))


def test_coincident_centroids_and_duplicates(sim_backend, orc_backend):
    """Equal Morton codes everywhere (all centroids identical): ties resolve by index."""
    base = np.array([[-1, -1, 0], [1, -1, 0], [0, 2, 0]], dtype=np.float32)
    v = np.concatenate([base * s for s in (1.0, 1.0, 1.0, 1.0, 1.0, 1.0)]).astype(np.float32)
    t = np.arange(18, dtype=np.uint32).reshape(-1, 3)
    mesh = api.TriangleMesh(Transform.identity(), t, v)
    a = api.Scene([api.GeometricPrimitive(mesh)], [], backend=sim_backend)
    b = api.Scene([api.GeometricPrimitive(mesh)], [], backend=orc_backend)
    codes, order = a.morton_codes_and_order()
    assert len(set(codes.tolist())) == 1 and np.array_equal(order, np.arange(6, dtype=np.uint32))
    parity.check_morton(a, b)
    rays = api.make_rays([[0, 0, 5]] * 4, [[0, 0, -1], [0.1, 0.2, -1], [3, 0, -1], [0, 0, 1]])
    ha, hb = a.intersect(rays), b.intersect(rays)
    assert np.array_equal(ha["prim"] == A.FTN_NO_HIT, hb["prim"] == A.FTN_NO_HIT)
    assert np.array_equal(ha["t"], hb["t"])          # all six coincide: any of them, same t


def test_empty_scene(sim_backend):
    s = api.Scene([], [], backend=sim_backend)
    rays = parity.random_ray_batch(64, 1)
    assert (s.intersect(rays)["prim"] == A.FTN_NO_HIT).all()
    assert not s.intersect_test(rays).any()


# ---- per-function parity with the oracle ---------------------------------------------------------
def _ray(o, d, t_max=float("inf")):
    r = A.FtnRay(); r.o[:] = [float(x) for x in o]; r.d[:] = [float(x) for x in d]; r.t_max = t_max; r.time = 0.0
    return r


def test_gamma_constants(sim, oracle):
    for n in range(1, 9):
        assert np.float32(sim.library().sim_kat_gamma(n)) == np.float32(oracle.library().orc_kat_gamma(n))


def test_triangle_function_bit_exact(sim, oracle):
    rng = np.random.default_rng(21)
    A3 = A.f32 * 3
    n_hit = 0
    for _ in range(4000):
        p = rng.uniform(-2, 2, (3, 3)).astype(np.float32)
        o = rng.uniform(-4, 4, 3).astype(np.float32)
        tgt = (p[0] * 0.3 + p[1] * 0.3 + p[2] * 0.4 + rng.normal(0, 0.4, 3)).astype(np.float32)
        r = _ray(o, tgt - o)
        oa, ob = (A.f32 * 4)(), (A.f32 * 4)()
        ha = sim.library().sim_kat_triangle_intersect(A3(*p[0]), A3(*p[1]), A3(*p[2]), C.byref(r), oa)
        hb = oracle.library().orc_kat_triangle_intersect(A3(*p[0]), A3(*p[1]), A3(*p[2]), C.byref(r), ob)
        assert ha == hb
        if ha:
            n_hit += 1
            assert list(oa) == list(ob)
    assert n_hit > 500


def test_triangle_f64_fallback_edges(sim, oracle):
    """Rays through shared vertices / along edges force e == 0 and the f64 retry (triangle.rs:219-223)."""
    A3 = A.f32 * 3
    p0, p1, p2 = (0, 0, 0), (1, 0, 0), (0, 1, 0)
    cases = [((0, 0, 1), (0, 0, -1)), ((0.5, 0, 1), (0, 0, -1)), ((0.5, 0.5, 1), (0, 0, -1)), ((1, 0, 1), (0, 0, -1)),
             ((0.25, 0.25, 2), (0, 0, -1)), ((0, 0.5, -3), (0, 0, 1))]
    for o, d in cases:
        r = _ray(o, d)
        oa, ob = (A.f32 * 4)(), (A.f32 * 4)()
        ha = sim.library().sim_kat_triangle_intersect(A3(*p0), A3(*p1), A3(*p2), C.byref(r), oa)
        hb = oracle.library().orc_kat_triangle_intersect(A3(*p0), A3(*p1), A3(*p2), C.byref(r), ob)
        assert ha == hb and (not ha or list(oa) == list(ob))


def test_offset_ray_origin_bit_exact(sim, oracle):
    rng = np.random.default_rng(3)
    A3 = A.f32 * 3
    for _ in range(2000):
        p = rng.uniform(-50, 50, 3); e = np.abs(rng.normal(0, 1e-5, 3)); n = rng.normal(0, 1, 3); n /= np.linalg.norm(n); d = rng.normal(0, 1, 3)
        if rng.random() < 0.2:
            n[rng.integers(3)] = 0.0
        oa, ob = A3(), A3()
        sim.library().sim_kat_offset_ray_origin(A3(*p), A3(*e), A3(*n), A3(*d), oa)
        oracle.library().orc_kat_offset_ray_origin(A3(*p), A3(*e), A3(*n), A3(*d), ob)
        assert list(oa) == list(ob)


def test_sphere_function_matches(sim, oracle):
    s = A.FtnSphere()
    tf = Transform.translate((1.0, -2.0, 0.5)) * Transform.scale(1.0, 1.0, 1.0)
    s.object_to_world[:] = tf.flat().tolist(); s.world_to_object[:] = tf.flat_inv().tolist()
    s.radius, s.z_min, s.z_max, s.phi_max_deg, s.reverse_orientation = 3.0, -3.0, 3.0, 360.0, 1
    rng = np.random.default_rng(8)
    hits = 0
    for _ in range(3000):
        inside = rng.random() < 0.5
        o = rng.uniform(-1.5, 1.5, 3) + np.array([1.0, -2.0, 0.5]) if inside else rng.uniform(-12, 12, 3)
        d = rng.normal(0, 1, 3)
        r = _ray(o, d)
        oa, ob = (A.f32 * 13)(), (A.f32 * 13)()
        ha = sim.library().sim_kat_sphere_intersect(C.byref(s), C.byref(r), oa)
        hb = oracle.library().orc_kat_sphere_intersect(C.byref(s), C.byref(r), ob)
        assert ha == hb
        if ha:
            hits += 1
            a, b = np.array(oa[:]), np.array(ob[:])
            assert a[0] == b[0]                                   # t: exact (EFloat arithmetic is exact ops)
            assert np.array_equal(a[1:7], b[1:7])                 # p, p_err: exact
            assert np.allclose(a[7:], b[7:], rtol=0, atol=2e-6)   # n, wo: sin(theta) differs by an ulp at most
    assert hits > 1500


def test_camera_rays_match(sim, oracle):
    for lens in (0.0, 0.25):
        cam = api.PerspectiveCamera(Transform.look_at((3, -20, 4), (0, 0, 1), (0, 0, 1)).inverse(), (640, 360), fov=35.0,
                                    lens_radius=lens, focal_dist=20.0, shutter_interval=(0.0, 1.0)).to_abi()
        rng = np.random.default_rng(2)
        for _ in range(500):
            fx, fy = rng.uniform(0, 640), rng.uniform(0, 360)
            lx, ly, tu = rng.random(3)
            ra, rb = A.FtnRay(), A.FtnRay()
            sim.library().sim_kat_camera_ray(C.byref(cam), fx, fy, lx, ly, tu, C.byref(ra))
            oracle.library().orc_kat_camera_ray(C.byref(cam), fx, fy, lx, ly, tu, C.byref(rb))
            a = np.array(list(ra.o) + list(ra.d) + [ra.time]); b = np.array(list(rb.o) + list(rb.d) + [rb.time])
            if lens == 0.0:
                assert np.array_equal(a, b)
            else:
                assert np.allclose(a, b, rtol=0, atol=1e-6)


def test_counter_sampler_matches_oracle(sim, oracle):
    for seed in (0, 1, 12345):
        for si in (0, 1, 77, 2 ** 33 + 5):
            for dim in (0, 1, 5, 12, 44):
                assert sim.library().sim_kat_counter_uniform(seed, si, dim) == oracle.library().orc_kat_counter_uniform(seed, si, dim)


def _material(kind):
    m = A.FtnMaterial()
    if kind == "matte":
        m.type = A.FTN_MATERIAL_MATTE; m.kd[:] = [0.5, 0.4, 0.3]
    elif kind == "metal":
        m.type = A.FTN_MATERIAL_METAL; m.eta[:] = [0.2, 0.92, 1.1]; m.k[:] = [3.9, 2.45, 2.14]
        m.u_roughness = m.v_roughness = 0.01; m.remap_roughness = 1
    elif kind == "metal_aniso":
        m.type = A.FTN_MATERIAL_METAL; m.eta[:] = [0.2, 0.92, 1.1]; m.k[:] = [3.9, 2.45, 2.14]
        m.u_roughness, m.v_roughness, m.remap_roughness = 0.2, 0.05, 0
    elif kind == "glass":          # rough glass: MicrofacetReflection(FresnelDielectric(1, eta)) + MicrofacetTransmission
        m.type = A.FTN_MATERIAL_GLASS; m.kr[:] = [0.9, 0.8, 1.0]; m.kt[:] = [0.7, 1.0, 0.9]; m.eta[:] = [1.5] * 3
        m.u_roughness = m.v_roughness = 0.3; m.remap_roughness = 1
    elif kind == "glass_aniso":
        m.type = A.FTN_MATERIAL_GLASS; m.kr[:] = [1.0] * 3; m.kt[:] = [1.0] * 3; m.eta[:] = [1.33] * 3
        m.u_roughness, m.v_roughness, m.remap_roughness = 0.25, 0.1, 0
    else:
        m.type = A.FTN_MATERIAL_PLASTIC; m.kd[:] = [0.25] * 3; m.ks[:] = [0.25] * 3; m.u_roughness = m.v_roughness = 0.1; m.remap_roughness = 1
    return m


@pytest.mark.parametrize("kind", ["matte", "metal", "metal_aniso", "plastic", "glass", "glass_aniso"])
def test_bsdf_matches_oracle(sim, oracle, kind):
    m = _material(kind)
    rng = np.random.default_rng(5)
    A3, A2 = A.f32 * 3, A.f32 * 2
    worst = 0.0
    for _ in range(1500):
        wo = rng.normal(0, 1, 3); wo /= np.linalg.norm(wo)
        wi = rng.normal(0, 1, 3); wi /= np.linalg.norm(wi)
        if rng.random() < 0.5 and kind != "matte":      # near the specular direction, where the lobe is
            wi = np.array([-wo[0], -wo[1], wo[2]]) + rng.normal(0, 0.02, 3); wi /= np.linalg.norm(wi)
        if kind.startswith("glass") and rng.random() < 0.5:      # through the surface, around the refracted direction
            wi = -wo + rng.normal(0, 0.3, 3); wi /= np.linalg.norm(wi)
        u = rng.random(2)
        oa, ob = (A.f32 * 12)(), (A.f32 * 12)()
        sim.library().sim_kat_bsdf(C.byref(m), A3(*wo), A3(*wi), A2(*u), oa)
        oracle.library().orc_kat_bsdf(C.byref(m), A3(*wo), A3(*wi), A2(*u), ob)
        a, b = np.array(oa[:], dtype=np.float64), np.array(ob[:], dtype=np.float64)
        assert a[4] == b[4]
        scale = np.maximum(np.abs(b), 1e-6)
        rel = np.abs(a - b) / scale
        rel[8:11] = np.abs(a[8:11] - b[8:11])            # directions: absolute
        worst = max(worst, rel.max())
    # same formulas, same libm here: differences come only from FMA-free vs plain evaluation order
    assert worst < 2e-3, worst


def _env_scene(backend, w, h, seed):
    rng = np.random.default_rng(seed)
    tex = (rng.random((h, w, 3)) ** 4 * 5.0).astype(np.float32)
    tex[h // 3, w // 4] = [300.0, 250.0, 200.0]
    light = api.InfiniteAreaLight.new_envmap(tex, Transform.rotate(30.0, (0.2, 0.3, 1.0)))
    return api.Scene([], [light], backend=backend)


@pytest.mark.parametrize("w,h", [(1, 1), (8, 4), (16, 16), (13, 7)])
def test_env_light_matches_oracle(sim, oracle, sim_backend, orc_backend, w, h):
    a, b = _env_scene(sim_backend, w, h, 4), _env_scene(orc_backend, w, h, 4)
    rng = np.random.default_rng(6)
    for _ in range(600):
        u = (A.f32 * 2)(*rng.random(2))
        oa, ob = (A.f32 * 11)(), (A.f32 * 11)()
        ra = sim.library().sim_kat_env(a.handle, u, oa)
        rb = oracle.library().orc_kat_env(b.handle, u, ob)
        assert ra == rb == 0
        x, y = np.array(oa[:], dtype=np.float64), np.array(ob[:], dtype=np.float64)
        assert np.allclose(x[:3], y[:3], atol=2e-6)                      # wi
        assert np.allclose(x[3:], y[3:], rtol=2e-4, atol=1e-6), (x, y)    # pdfs and radiances


def test_slab_test_forms_match_reference(sim, oracle):
    """bounds.rs:214-233.  The kernels' slab test (per-ray dispatch between the exact form and the
    cheaper NaN-free form) accepts exactly the boxes the reference accepts, with the same entry
    distance -- including rays with zero / denormal direction components, origins on slab planes,
    rays starting inside the box and boxes behind t_max."""
    rng = np.random.default_rng(11)
    slib, olib = sim.library(), oracle.library()
    n_fast = n_hit = 0
    for i in range(6000):
        lo = rng.uniform(-2, 2, 3).astype(np.float32)
        hi = (lo + rng.uniform(0, 2, 3).astype(np.float32) * (rng.random(3) > 0.1)).astype(np.float32)   # some flat boxes
        o = rng.uniform(-3, 3, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        kind = i % 6
        if kind == 1:
            d[rng.integers(3)] = 0.0
        elif kind == 2:
            d[rng.integers(3)] = -0.0
            o[rng.integers(3)] = lo[rng.integers(3)]
        elif kind == 3:                                   # origin on a slab plane, axis-parallel ray
            ax = rng.integers(3); o[ax] = (lo if rng.random() < 0.5 else hi)[ax]; d[ax] = 0.0
        elif kind == 4:
            d[rng.integers(3)] = np.float32(1e-42)        # denormal: 1/d overflows to inf
        elif kind == 5:
            o = (lo + (hi - lo) * rng.random(3).astype(np.float32)).astype(np.float32)   # inside
        t_max = np.float32(np.inf if rng.random() < 0.5 else rng.uniform(0, 6))
        r = A.FtnRay()
        r.o[:] = o.tolist(); r.d[:] = d.tolist(); r.t_max = float(t_max); r.time = 0.0
        blo, bhi = (A.f32 * 3)(*lo.tolist()), (A.f32 * 3)(*hi.tolist())
        ref_t = (A.f32 * 2)()
        ref_hit = olib.orc_kat_bounds_intersect(blo, bhi, C.byref(r), ref_t)
        e_k, e_x = (A.f32 * 1)(), (A.f32 * 1)()
        got = slib.sim_kat_slab_test(blo, bhi, C.byref(r), e_k, 0)
        exact = slib.sim_kat_slab_test(blo, bhi, C.byref(r), e_x, 1)
        assert (got & 1) == ref_hit and (exact & 1) == ref_hit, (i, lo, hi, o, d, t_max)
        if ref_hit:
            n_hit += 1
            assert np.float32(e_k[0]) == np.float32(ref_t[0]) and np.float32(e_x[0]) == np.float32(ref_t[0])
        n_fast += (got >> 1) & 1
        if kind in (1, 2, 3, 4):
            assert not (got >> 1) & 1                     # irregular rays must take the exact form
    assert n_fast > 1500 and n_hit > 500


# ---- both topology builders on the host-compiled build code (ftn_lbvh.cuh / ftn_ploc.cuh) ----------------
@pytest.mark.parametrize("builder", ["lbvh", "ploc"])
def test_builders_give_identical_results(sim_backend, orc_backend, rounded_cube_path, monkeypatch, builder):
    monkeypatch.setenv("FTN_BVH_BUILDER", builder)
    a, b = parity.cube_scenes(sim_backend, orc_backend, rounded_cube_path)
    parity.check_morton(a, b)
    parity.check_ray_batch(a, b, parity.random_ray_batch(20000, 15), builder)
    for n_tris in (5, 6, 9, 33, 500):
        rng = np.random.default_rng(n_tris)
        v = rng.uniform(-1, 1, (3 * n_tris, 3)).astype(np.float32)
        t = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
        mesh = api.TriangleMesh(Transform.identity(), t, v)
        sa = api.Scene([api.GeometricPrimitive(mesh)], [], backend=sim_backend)
        sb = api.Scene([api.GeometricPrimitive(mesh)], [], backend=orc_backend)
        rays = parity.random_ray_batch(3000, 300 + n_tris, extent=1.0, far=4.0)
        parity.compare_hits(sa.intersect(rays), sb.intersect(rays), "%s %d" % (builder, n_tris))
    # coincident triangles + a geometric size progression
    base = np.array([[-1, -1, 0], [1, -1, 0], [0, 2, 0]], dtype=np.float32)
    v = np.concatenate([base] * 40 + [base * np.float32(1.1 ** k) + np.float32([0, 0, -0.01 * k]) for k in range(1, 60)]).astype(np.float32)
    t = np.arange(len(v), dtype=np.uint32).reshape(-1, 3)
    mesh = api.TriangleMesh(Transform.identity(), t, v)
    sa = api.Scene([api.GeometricPrimitive(mesh)], [], backend=sim_backend)
    sb = api.Scene([api.GeometricPrimitive(mesh)], [], backend=orc_backend)
    rays = parity.random_ray_batch(3000, 23, extent=2.0, far=6.0)
    ha, hb = sa.intersect(rays), sb.intersect(rays)
    assert np.array_equal(ha["prim"] == A.FTN_NO_HIT, hb["prim"] == A.FTN_NO_HIT) and np.array_equal(ha["t"], hb["t"])


def test_ploc_depth_fallback_to_radix_tree(sim_backend, orc_backend, rounded_cube_path, monkeypatch):
    """A PLOC tree deeper than the traversal stack allows is discarded for the depth-bounded radix tree
    (forced here through the test hook FTN_PLOC_MAX_DEPTH)."""
    monkeypatch.setenv("FTN_BVH_BUILDER", "ploc")
    monkeypatch.setenv("FTN_PLOC_MAX_DEPTH", "3")
    a, b = parity.cube_scenes(sim_backend, orc_backend, rounded_cube_path)
    monkeypatch.setenv("FTN_BVH_BUILDER", "lbvh")
    r, _ = parity.cube_scenes(sim_backend, orc_backend, rounded_cube_path)
    assert a.stats()["bvh_nodes"] == r.stats()["bvh_nodes"]          # the radix tree's node count for this mesh
    parity.check_morton(a, b)
    parity.check_ray_batch(a, b, parity.random_ray_batch(20000, 15), "fallback")


# ---- the device's MIPMap lookup (ftn_shade.cuh mip_lookup_trilinear) against the oracle's (mipmap.rs:245-311) ------
@pytest.mark.parametrize("wrap", ["repeat", "black", "clamp"])
def test_mipmap_lookup_matches_oracle(sim, oracle, wrap):
    rng = np.random.default_rng(17)
    for shape in ((6, 8), (16, 16), (5, 33), (1, 1)):
        mp = api.MIPMap(rng.random(shape + (3,)).astype(np.float32), wrap)
        ptr = mp.packed.ctypes.data_as(C.POINTER(C.c_float))
        for _ in range(400):
            s, t = (float(x) for x in rng.uniform(-1.5, 2.5, 2))
            w = float(rng.choice([0.0, 1e-9, 1e-3, 0.05, 0.13, 0.26, 0.51, 0.99, 1.0, 2.0, 7.0]))
            a, b = (C.c_float * 3)(), (C.c_float * 3)()
            assert sim.library().sim_kat_mipmap_lookup(ptr, mp.width, mp.height, len(mp.levels), mp.wrap, s, t, w, a) == 0
            assert oracle.library().orc_kat_mipmap_lookup(ptr, mp.width, mp.height, len(mp.levels), mp.wrap, s, t, w, b) == 0
            assert np.allclose(np.array(list(a)), np.array(list(b)), rtol=1e-5, atol=1e-7), (shape, s, t, w, list(a), list(b))


def test_guided_cdf_search_equals_binary_search(sim):
    """The env light's two cdf searches go through guide tables on the device (ftn_shade.cuh search_sorted_le_guided):
    same index as Distribution1D's binary search (sampling.rs:66-81) for flat, peaked, zero-run and tiny rows, at bucket
    edges and for u just below 1."""
    rng = np.random.default_rng(23)
    lib = sim.library()
    lib.sim_kat_guided_search.restype = C.c_int
    lib.sim_kat_guided_search.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    for n in (1, 2, 3, 7, 64, 1000, 2048):
        rows = [np.ones(n), rng.random(n), rng.random(n) ** 12, np.where(rng.random(n) < 0.7, 0.0, rng.random(n)), np.zeros(n)]
        peak = np.full(n, 1e-6); peak[n // 3] = 1e4
        rows.append(peak)
        for f in rows:
            f = np.ascontiguousarray(f, dtype=np.float32)
            u = np.concatenate([rng.random(20000), np.arange(n + 1) / max(n, 1), np.nextafter(np.arange(1, n + 1) / n, 0), [0.0, np.nextafter(np.float32(1.0), np.float32(0.0))]]).astype(np.float32)
            u = np.ascontiguousarray(np.clip(u, 0.0, np.nextafter(np.float32(1.0), np.float32(0.0))))
            bad_at = C.c_int(-1)
            bad = lib.sim_kat_guided_search(f.ctypes.data_as(C.POINTER(C.c_float)), n, u.ctypes.data_as(C.POINTER(C.c_float)), len(u), C.byref(bad_at))
            assert bad == 0, (n, bad, float(u[bad_at.value]))
