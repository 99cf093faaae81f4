"""Design experiment (CPU, tests/hostsim): nodes / triangles visited per ray with the LBVH the GPU
builds vs a binned top-down SAH tree in the same node format, plus the SIMT issue-slot model."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fountain_b200 import _abi as A, api
from workloads import scenes  # noqa: E402
from tests.hostsim import sim  # noqa: E402

be = sim.backend()
lib = sim.library()
lib.sim_warp_model.restype = C.c_int
lib.sim_warp_model.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(A.FtnRay), C.POINTER(C.c_int), C.POINTER(C.c_double)]
lib.sim_bvh_rebuild_sah.restype = C.c_int
lib.sim_bvh_rebuild_sah.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]


def model(scene, batch, vote):
    batch = np.ascontiguousarray(batch)
    ip = (C.c_int * 7)(16, 1, 1, 1, 64, 1 if vote else 0, 14)
    out = (C.c_double * 8)()
    lib.sim_warp_model(scene.handle, len(batch), batch.ctypes.data_as(C.POINTER(A.FtnRay)), ip, out)
    ns, nw, ts, tw, ls, lw, rf, n = out[:]
    return dict(nodes=nw / n, tris=tw / n, node_slots=ns / n, tri_slots=ts / n, cost=(84 * ns + 72 * ts + 44 * ls + 250 * rf) / n)


which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    scene, camera = scenes.synthetic_mesh_scene(1000, 500, backend=be, resolution=(256, 256))
    prim = scenes.primary_ray_batch(camera, (256, 256))
    vote = True
elif which == "c4":
    scene, camera, film = scenes.logo_style_scene(backend=be, resolution=(256, 144))
    prim = scenes.primary_ray_batch(camera, (256, 144))
    vote = True
else:
    scene, camera, film = scenes.rounded_cube_scene(backend=be, resolution=(256, 256))
    prim = scenes.primary_ray_batch(camera, (256, 256))
    vote = False
hits = scene.intersect(prim)
inc = scenes.diffuse_bounce_batch(prim, hits, scene._positions, scene._indices, seed=2)
batches = {"primary": prim, "diffuse": inc}
if which == "c3":
    rng = np.random.default_rng(4)
    n_int = 1 << 15
    o = rng.normal(size=(n_int, 3)); o *= (rng.random((n_int, 1)) ** (1 / 3) * 9.0) / np.linalg.norm(o, axis=1, keepdims=True)
    d = rng.normal(size=(n_int, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    batches["interior"] = api.make_rays(o.astype(np.float32), d.astype(np.float32))
print("tris", scene.n_triangles, {k: len(v) for k, v in batches.items()}, flush=True)
ref_hits = {k: scene.intersect(v) for k, v in batches.items()}
print("LBVH nodes", scene.stats()["bvh_nodes"])
for k, v in batches.items():
    print("  LBVH %-9s" % k, {a: round(b, 2) for a, b in model(scene, v, vote).items()}, flush=True)
for leaf_max, ct, ci in ((4, 1.0, 1.0), (4, 1.0, 0.6), (2, 1.0, 1.0), (8, 1.0, 0.5)):
    t = time.time()
    assert lib.sim_bvh_rebuild_sah(scene.handle, leaf_max, ct, ci) == 0
    print("SAH leaf<=%d ct %.1f ci %.1f: built %.1fs" % (leaf_max, ct, ci, time.time() - t), flush=True)
    for k, v in batches.items():
        h = scene.intersect(v)
        same = np.array_equal(h["prim"] == 0xFFFFFFFF, ref_hits[k]["prim"] == 0xFFFFFFFF) and np.array_equal(h["t"], ref_hits[k]["t"])
        print("  SAH  %-9s" % k, {a: round(b, 2) for a, b in model(scene, v, vote).items()}, "same hits:", same, flush=True)
