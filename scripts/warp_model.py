"""Design tool: SIMT issue-slot model of the persistent traversal loop (tests/hostsim sim_warp_model)
on the C3 scene, for different loop policies.  CPU only."""
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

tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_lon = int(round(tris ** 0.5))
t0 = time.time()
scene, camera = scenes.synthetic_mesh_scene(n_lon, n_lon // 2, backend=be, resolution=(384, 384))
print("scene built %.1fs" % (time.time() - t0), flush=True)
prim = scenes.primary_ray_batch(camera, (384, 384))
hits = scene.intersect(prim)
inc = scenes.diffuse_bounce_batch(prim, hits, scene._positions, scene._indices, seed=2)
rng = np.random.default_rng(4)
n_int = 1 << 16
o = rng.normal(size=(n_int, 3)); o *= (rng.random((n_int, 1)) ** (1 / 3) * 9.0) / np.linalg.norm(o, axis=1, keepdims=True)
d = rng.normal(size=(n_int, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
interior = api.make_rays(o.astype(np.float32), d.astype(np.float32))
print("rays:", len(prim), len(inc), len(interior), flush=True)

C_NODE, C_TRI, C_LEAF, C_REFILL = 50.0, 60.0, 14.0, 250.0


def model(batch, refill, node_exit, leaf_exit, postpone, mode=0, alpha16=16, warps=64):
    batch = np.ascontiguousarray(batch)
    ip = (C.c_int * 7)(refill, node_exit, leaf_exit, postpone, warps, mode, alpha16)
    out = (C.c_double * 8)()
    lib.sim_warp_model(scene.handle, len(batch), batch.ctypes.data_as(C.POINTER(A.FtnRay)), ip, out)
    ns, nw, ts, tw, ls, lw, rf, n = out[:]
    cost = C_NODE * ns + C_TRI * ts + C_LEAF * ls + C_REFILL * rf
    return dict(cost_per_ray=cost / n, node_eff=nw / max(ns, 1), tri_eff=tw / max(ts, 1), nodes_per_ray=nw / n, tris_per_ray=tw / n,
                node_slots_per_ray=ns / n, tri_slots_per_ray=ts / n, refills_per_ray=rf / n)


for label, batch in (("incoherent_diffuse", inc), ("interior", interior), ("primary", prim[: 1 << 16])):
    print("==", label)
    cfgs = [(16, 1, 1, 1, 0, 16), (16, 8, 8, 1, 0, 16), (16, 0, 0, 1, 1, 16), (16, 0, 0, 1, 1, 12), (16, 0, 0, 1, 2, 16), (16, 0, 0, 1, 2, 12),
            (16, 0, 0, 1, 2, 20), (16, 0, 0, 1, 2, 8), (12, 0, 0, 1, 2, 12), (20, 0, 0, 1, 2, 12)]
    for cfg in cfgs:
        r = model(batch, *cfg)
        print("  refill<%2d node_exit<%2d leaf_exit<%2d postpone %d mode %d a %2d : cost/ray %7.1f  node eff %5.1f  tri eff %5.1f  nodes/ray %5.1f tris/ray %4.1f  slots n %.2f t %.2f  refills/ray %.3f"
              % (*cfg, r["cost_per_ray"], r["node_eff"], r["tri_eff"], r["nodes_per_ray"], r["tris_per_ray"], r["node_slots_per_ray"], r["tri_slots_per_ray"], r["refills_per_ray"]), flush=True)
