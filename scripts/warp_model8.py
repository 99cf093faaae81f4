"""Design tool: SIMT issue-slot model of the BVH8q persistent traversal loop (tests/hostsim sim_warp_model8) on the C3
scene, for different loop policies.  CPU only.  usage: warp_model8.py [triangles]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FTN_BVH_LAYOUT"] = "bvh8"
from fountain_b200 import _abi as A, api  # noqa: E402
from workloads import scenes  # noqa: E402
from tests.hostsim import sim  # noqa: E402

be = sim.backend()
lib = sim.library()
lib.sim_warp_model8.restype = C.c_int
lib.sim_warp_model8.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(A.FtnRay), C.POINTER(C.c_int), C.POINTER(C.c_double)]

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

C_NODE, C_TRI, C_ITER, C_REFILL = 230.0, 150.0, 30.0, 300.0


def model(batch, refill, bias, one_tri, postpone, coop, warps=64):
    batch = np.ascontiguousarray(batch)
    ip = (C.c_int * 6)(refill, bias, warps, one_tri, postpone, coop)
    out = (C.c_double * 8)()
    assert lib.sim_warp_model8(scene.handle, len(batch), batch.ctypes.data_as(C.POINTER(A.FtnRay)), ip, out) == 0
    ns, nw, ts, tw, it, rf, n, msp = out[:]
    cost = C_NODE * ns + C_TRI * ts * (1.25 if coop else 1.0) + C_ITER * it + C_REFILL * rf
    return dict(cost=cost / n * 32, node_eff=nw / max(ns, 1), tri_eff=tw / max(ts, 1), nodes=nw / n, tris=tw / n, iters=it / n * 32, refills=rf / n * 32, msp=msp)


for label, batch in (("incoherent_diffuse", inc[: 1 << 16]), ("interior", interior), ("primary", prim[: 1 << 16])):
    print("==", label)
    for cfg in [(16, 14, 0, 0, 0), (16, 24, 0, 0, 0), (24, 24, 0, 0, 0), (16, 24, 1, 0, 0), (16, 24, 0, 1, 0), (16, 14, 0, 1, 0), (16, 40, 0, 1, 0), (16, 24, 1, 1, 0),
                (16, 24, 0, 0, 1), (16, 24, 0, 1, 1), (16, 48, 0, 1, 1), (24, 48, 0, 1, 1)]:
        r = model(batch, *cfg)
        print("  refill<%2d bias %2d one_tri %d postpone %d coop %d : cost/ray %7.1f  node eff %5.1f  tri eff %5.1f  nodes/ray %5.1f tris/ray %4.1f  iters/ray %.2f refills/ray %.3f  max sp %d"
              % (*cfg, r["cost"], r["node_eff"], r["tri_eff"], r["nodes"], r["tris"], r["iters"], r["refills"], r["msp"]), flush=True)
