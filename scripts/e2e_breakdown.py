import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fountain_b200 import api
from workloads import scenes
gpu = api.default_backend()
def T(label, f):
    t=time.perf_counter(); r=f(); dt=(time.perf_counter()-t)*1e3; print("%-28s %8.2f ms"%(label,dt)); return r
for it in range(3):
    print("--- iteration", it)
    mesh = T("TriangleMesh.from_ply", lambda: api.TriangleMesh.from_ply(scenes.ROUNDED_CUBE_PLY))
    prim = api.GeometricPrimitive(mesh, api.MatteMaterial(0.5))
    scene = T("Scene create (no build)", lambda: api.Scene([prim],[api.InfiniteAreaLight.new_uniform(1.0)],backend=gpu,build=False))
    T("bvh_build", scene.build)
    _, camera, film = T("rounded_cube_scene (all)", lambda: scenes.rounded_cube_scene(backend=gpu, resolution=(512,512)))
    integ = api.SamplerIntegrator(camera, api.PathIntegrator(5,1.0))
    st = T("render 64spp (1st, ws alloc)", lambda: integ.render_parallel(scene, film, api.RandomSampler.new_with_seed(64,0)))
    st = T("render 64spp (2nd)", lambda: integ.render_parallel(scene, film, api.RandomSampler.new_with_seed(64,0)))
    print("   device_seconds %.2f ms"%(st["device_seconds"]*1e3))
    T("scene.close", scene.close)
