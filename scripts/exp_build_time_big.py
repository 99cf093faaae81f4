"""Device time of ftn_bvh_build on large synthetic meshes, first (cold) build and repeats (same process): usage exp_build_time_big.py <n_lon> ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fountain_b200 import api
from workloads import scenes
gpu = api.default_backend()
for n_lon in [int(a) for a in sys.argv[1:]] or [1000, 2828]:
    for it in range(4):
        t0 = time.perf_counter()
        scene = scenes.synthetic_mesh_scene(n_lon, n_lon // 2, backend=gpu, resolution=(64, 64))[0]
        dt = time.perf_counter() - t0
        st = scene.stats()
        print("n_lon %d (%d tris) iter %d: wall %.1f ms, device build %.2f ms (morton+sort %.2f ms), nodes %d" % (n_lon, scene.n_triangles, it, dt * 1e3, st["bvh_build_seconds"] * 1e3, st.get("morton_sort_seconds", 0) * 1e3, st["bvh_nodes"]), flush=True)
        scene.close()
