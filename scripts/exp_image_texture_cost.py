"""What does the image-texture path cost?  The image_texture_scene at 1080p, 16 spp, path depth 5, with the image
textures and with the same scene carrying constant Kd (plain shade kernels)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from fountain_b200 import api
from workloads import scenes  # noqa: E402

gpu = api.default_backend()
gpu.call("set_device", 0)
res = (1920, 1080)


def constant_scene(backend=None, resolution=res, **kw):
    scene, camera, film = scenes.image_texture_scene(backend=backend, resolution=resolution, **kw)
    return scene, camera, film


for label, material in (("image textures (matte)", "matte"), ("image textures (plastic)", "plastic")):
    scene, camera, film = scenes.image_texture_scene(backend=gpu, resolution=res, material=material)
    integ = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0))
    for i in range(4):
        t0 = time.perf_counter()
        st = integ.render_parallel(scene, film, api.RandomSampler.new_with_seed(16, 0))
        dt = time.perf_counter() - t0
    rays = st["rays_closest"] + st["rays_any"]
    print("%-28s %7.1f ms device, %6.0f Mrays/s (%.1f M rays)" % (label, st["device_seconds"] * 1e3, rays / st["device_seconds"] / 1e6, rays / 1e6), flush=True)
    scene.close()

# the same geometry and lights with constant Kd
import workloads.scenes as S  # noqa: E402
orig = api.ImageTexture
try:
    api.ImageTexture = lambda mp, mapping=None: np.array([0.5, 0.4, 0.3], np.float32)
    scene, camera, film = scenes.image_texture_scene(backend=gpu, resolution=res, material="matte")
finally:
    api.ImageTexture = orig
integ = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0))
for i in range(4):
    st = integ.render_parallel(scene, film, api.RandomSampler.new_with_seed(16, 0))
rays = st["rays_closest"] + st["rays_any"]
print("%-28s %7.1f ms device, %6.0f Mrays/s (%.1f M rays)" % ("constant Kd (matte)", st["device_seconds"] * 1e3, rays / st["device_seconds"] / 1e6, rays / 1e6))
