"""Wall and device time of scene upload + BVH build, per iteration (builder from FTN_BVH_BUILDER)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fountain_b200 import api
from workloads import scenes
gpu = api.default_backend()
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
for it in range(8):
    t0 = time.perf_counter()
    if which == "c2":
        scene = scenes.rounded_cube_scene(backend=gpu, resolution=(64, 64))[0]
    else:
        scene = scenes.synthetic_mesh_scene(1000, 500, backend=gpu, resolution=(64, 64))[0]
    dt = time.perf_counter() - t0
    st = scene.stats()
    print("%s iter %d: wall %.2f ms, device build %.2f ms, nodes %d" % (os.environ.get("FTN_BVH_BUILDER", "default"), it, dt * 1e3, st["bvh_build_seconds"] * 1e3, st["bvh_nodes"]), flush=True)
    scene.close()
