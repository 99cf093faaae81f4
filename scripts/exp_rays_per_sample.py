"""Sanity: rays per camera sample must not depend on spp / number of passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fountain_b200 import api
from workloads import scenes
gpu = api.default_backend()
scene, camera, film = scenes.logo_style_scene(backend=gpu, resolution=(1920, 1080))
for spp in (8, 16, 24, 64, 128):
    st = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0)).render_parallel(scene, film, api.RandomSampler.new_with_seed(spp, 0))
    rays = st["rays_closest"] + st["rays_any"]
    print("spp %4d: samples %d rays %d rays/sample %.3f closest %d any %d device %.1f ms mean %.4f" % (
        spp, st["camera_samples"], rays, rays / st["camera_samples"], st["rays_closest"], st["rays_any"], st["device_seconds"] * 1e3, float(film.pixels[..., :3].mean())), flush=True)
