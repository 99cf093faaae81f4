"""PLOC tree depth as a function of scene size (FTN_DEBUG_BUILD prints it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FTN_DEBUG_BUILD"] = "1"
os.environ["FTN_BVH_BUILDER"] = "ploc"
from fountain_b200 import api
from workloads import scenes
gpu = api.default_backend()
for n_lon in (1000, 2000, 4000, 7071):
    scene = scenes.synthetic_mesh_scene(n_lon, n_lon // 2, backend=gpu, resolution=(64, 64))[0]
    print("tris", scene.n_triangles, "nodes", scene.stats()["bvh_nodes"], flush=True)
    scene.close()
scene = scenes.logo_style_scene(backend=gpu, resolution=(64, 64))[0]
print("c4 tris", scene.n_triangles, "nodes", scene.stats()["bvh_nodes"], flush=True)
