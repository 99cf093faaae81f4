"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : N-GPU sharded render == 1-GPU render."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from fountain_b200 import api
from workloads import scenes
from fountain_b200.distributed import render_sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gpu = api.default_backend(); gpu.call("set_device", local)
scene, camera, film = scenes.rounded_cube_scene(backend=gpu, resolution=(256, 256))
integ = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0))
sampler = api.RandomSampler.new_with_seed(16, 3)
st = render_sharded(integ, scene, film, sampler)
if rank == 0:
    sharded = film.pixels.copy()
    integ.render_parallel(scene, film, sampler)
    ok_w = np.array_equal(sharded[..., 3], film.pixels[..., 3])
    ok_c = np.allclose(sharded[..., :3], film.pixels[..., :3], rtol=1e-5, atol=1e-6)
    print("multi_gpu_check world=%d weights_equal=%s colours_close=%s rays=%d" % (world, ok_w, ok_c, st["rays_closest"] + st["rays_any"]))
    assert ok_w and ok_c
dist.destroy_process_group()
