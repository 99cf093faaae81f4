"""Where does the host-side time of an e2e step go?  cProfile of workload_scene (scene description -> ftn_scene_create -> ftn_bvh_build)."""
import cProfile
import os
import pstats
import sys
import argparse

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fountain_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
args = argparse.Namespace(spp=0, detail=1.0, tris=1_000_000)
gpu = api.default_backend()
gpu.call("set_device", 0)
for _ in range(3):
    sc = bench.workload_scene(name, gpu, args)[0]; sc.close()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    sc = bench.workload_scene(name, gpu, args)[0]; sc.close()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
