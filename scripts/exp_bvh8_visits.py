"""Design tool (CPU only, host-compiled kernels): node visits and triangle tests per ray of the two node layouts on the C3
scene.  usage: exp_bvh8_visits.py [triangles]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fountain_b200 import _abi as A, api  # noqa: E402
from workloads import scenes  # noqa: E402
from tests.hostsim import sim  # noqa: E402

be = sim.backend()
lib = sim.library()
tris = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
n_lon = int(round(tris ** 0.5))
for layout in ("bvh2", "bvh8"):
    os.environ["FTN_BVH_LAYOUT"] = layout
    t0 = time.time()
    scene, camera = scenes.synthetic_mesh_scene(n_lon, n_lon // 2, backend=be, resolution=(256, 256))
    st = scene.stats()
    print("%s: %d tris, %d nodes x %d B, built in %.1fs" % (layout, scene.n_triangles, st["bvh_nodes"], st["bvh_node_bytes"], time.time() - t0), flush=True)
    prim = scenes.primary_ray_batch(camera, (256, 256))
    hits = scene.intersect(prim)
    inc = scenes.diffuse_bounce_batch(prim, hits, scene._positions, scene._indices, seed=2)
    rng = np.random.default_rng(4)
    n_int = 1 << 15
    o = rng.normal(size=(n_int, 3)); o *= (rng.random((n_int, 1)) ** (1 / 3) * 9.0) / np.linalg.norm(o, axis=1, keepdims=True)
    d = rng.normal(size=(n_int, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    interior = api.make_rays(o.astype(np.float32), d.astype(np.float32))
    for label, batch in (("primary", prim), ("incoherent_diffuse", inc), ("interior", interior)):
        batch = np.ascontiguousarray(batch)
        ctr = (C.c_uint64 * 2)(0, 0)
        lib.sim_intersect_count(scene.handle, len(batch), batch.ctypes.data_as(C.POINTER(A.FtnRay)), None, ctr)
        n = len(batch)
        nodes, tr = ctr[0] / n, ctr[1] / n
        loads = nodes * (3 if layout == "bvh8" else 2) + tr * 2
        print("   %-20s nodes/ray %6.2f  tris/ray %5.2f  lane-loads/ray %6.1f" % (label, nodes, tr, loads), flush=True)
    scene.close()
