"""Parity checks shared by the CPU host-sim tier and the `-m gpu` tier: every function takes the
backend under test and the oracle backend and compares them on the same seeded inputs."""
import ctypes as C

import numpy as np

from fountain_b200 import _abi as A
from fountain_b200 import api
from workloads import scenes
from tests.conftest import unit_sphere_dirs

NO_HIT = A.FTN_NO_HIT


def ulps(a, b):
    """|a-b| in units of the spacing at b (f32)."""
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    return np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.spacing(np.abs(b)).astype(np.float64)


def compare_hits(test_hits, ref_hits, what=""):
    """SURVEY 8c tolerance: hit/miss identical; t bit-identical, except where candidates tie
    within rounding (then either primitive is accepted and t may differ by a few ulps)."""
    tm, rm = test_hits["prim"] == NO_HIT, ref_hits["prim"] == NO_HIT
    assert np.array_equal(tm, rm), "%s: hit/miss differs on %d rays" % (what, int((tm != rm).sum()))
    h = ~rm
    same_prim = h & (test_hits["prim"] == ref_hits["prim"])
    diff_prim = h & ~same_prim
    # identical primitive => identical arithmetic => identical bits
    assert np.array_equal(test_hits["t"][same_prim], ref_hits["t"][same_prim]), "%s: t differs for the same primitive" % what
    assert np.array_equal(test_hits["b1"][same_prim], ref_hits["b1"][same_prim])
    assert np.array_equal(test_hits["b2"][same_prim], ref_hits["b2"][same_prim])
    if diff_prim.any():
        u = ulps(test_hits["t"][diff_prim], ref_hits["t"][diff_prim])
        assert u.max() <= 8, "%s: different primitive with t %g ulps apart" % (what, u.max())
    frac_exact = float(same_prim.sum()) / max(1, int(h.sum()))
    assert frac_exact > 0.995, "%s: only %.4f of hits pick the same primitive" % (what, frac_exact)
    return {"rays": int(len(rm)), "hits": int(h.sum()), "same_prim": int(same_prim.sum()), "tie_prims": int(diff_prim.sum())}


def cube_scenes(test_backend, ref_backend, ply):
    mesh = api.TriangleMesh.from_ply(ply)
    a = api.Scene([api.GeometricPrimitive(mesh)], [], backend=test_backend)
    b = api.Scene([api.GeometricPrimitive(mesh)], [], backend=ref_backend)
    return a, b


def check_morton(test_scene, ref_scene):
    """Integer work: Morton codes and the sorted primitive order are bit-exact."""
    tc, to = test_scene.morton_codes_and_order()
    rc, ro = ref_scene.morton_codes_and_order()
    assert np.array_equal(tc, rc), "morton codes differ on %d triangles" % int((tc != rc).sum())
    assert np.array_equal(to, ro), "sorted order differs at %d positions" % int((to != ro).sum())
    assert np.array_equal(np.sort(to), np.arange(len(to), dtype=np.uint32))
    return len(tc)


def random_ray_batch(n, seed, extent=9.0, far=30.0):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-far, far, (n, 3)).astype(np.float32)
    tgt = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    rays = api.make_rays(o, tgt - o)
    # a third of the rays get a finite t_max around the hit distance (shadow-ray-like)
    k = n // 3
    rays["t_max"][:k] = rng.uniform(0.2, 1.5, k).astype(np.float32)
    return rays


def check_ray_batch(test_scene, ref_scene, rays, what=""):
    stats = compare_hits(test_scene.intersect(rays), ref_scene.intersect(rays), what)
    assert np.array_equal(test_scene.intersect_test(rays), ref_scene.intersect_test(rays)), what + ": any-hit differs"
    return stats


def check_watertight(scene, n=100_000, seed=7):
    """tests/tri_watertight.rs: every direction from the origin hits, both queries."""
    dirs = unit_sphere_dirs(n, seed)
    rays = api.make_rays(np.zeros((n, 3)), dirs)
    assert scene.intersect_test(rays).all()
    hits = scene.intersect(rays)
    assert (hits["prim"] != NO_HIT).all()
    return hits


def render(backend, scene_fn, integrator, spp, seed=0, **kw):
    scene, camera, film = scene_fn(backend=backend, **kw)
    stats = api.SamplerIntegrator(camera, integrator).render_parallel(scene, film, api.RandomSampler.new_with_seed(spp, seed))
    rgb, (w, h) = film.into_spectrum_buffer()
    return rgb.reshape(h, w, 3), film.pixels.copy(), stats


def image_diff(test_rgb, ref_rgb):
    """Counter-sampler A/B: both arms trace the same paths, so images agree except where an ulp
    of transcendental-function difference flips a discrete event (hit/miss at a silhouette, an RR
    decision).  Returns (mean relative error, fraction of pixels off by > 1e-3 relative)."""
    d = np.abs(test_rgb.astype(np.float64) - ref_rgb.astype(np.float64))
    rel = d / np.maximum(np.abs(ref_rgb.astype(np.float64)), 1e-3)
    per_px = rel.max(axis=-1)
    return float(rel.mean()), float((per_px > 1e-3).mean())


def rel_mse(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def quad_light_order(orc_backend, **kw):
    """The triangle permutation of scenes.quad_light_scene's emissive quads for which the reference's light list
    (area lights in ITS BVH's primitive order, scene/mod.rs:40-44 -- read back from the oracle) equals the ABI's
    (primitive order): with it the oracle and the CUDA path pick the same light for the same random number."""
    from oracle import orc
    from workloads import scenes
    for order in ((0, 1), (1, 0)):
        sc, _, _ = scenes.quad_light_scene(backend=orc_backend, light_order=order, **kw)
        prims = orc.light_prims(sc)
        sc.close()
        if prims == sorted(prims):
            return order
    return None    # (with several emissive meshes the reference's BVH leaves list them in an order no layout reproduces)
