"""CPU tier: the N > 1 host logic (sample-index sharding + one film reduction) with world size 2
over gloo, on the oracle backend."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fountain_b200 import api
    from workloads import scenes
    from fountain_b200.distributed import render_sharded
    from oracle import orc
    orc.set_threads(2)
    be = orc.backend()
    scene, camera, film = scenes.rounded_cube_scene(backend=be, resolution=(32, 32))
    integ = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0))
    sampler = api.RandomSampler.new_with_seed(8, 11)
    stats = render_sharded(integ, scene, film, sampler)
    if rank == 0:
        np.save(os.path.join(out_dir, "sharded.npy"), film.pixels)
        np.save(os.path.join(out_dir, "stats.npy"), np.array([stats["camera_samples"], stats["rays_closest"], stats["rays_any"]]))
    dist.destroy_process_group()


def test_two_rank_sharded_render_equals_single(tmp_path):
    from fountain_b200 import api
    from workloads import scenes
    from oracle import orc
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sharded = np.load(tmp_path / "sharded.npy")
    be = orc.backend()
    scene, camera, film = scenes.rounded_cube_scene(backend=be, resolution=(32, 32))
    st = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0)).render_parallel(scene, film, api.RandomSampler.new_with_seed(8, 11))
    assert np.array_equal(sharded[..., 3], film.pixels[..., 3])                     # every sample rendered exactly once
    assert np.allclose(sharded[..., :3], film.pixels[..., :3], rtol=1e-5, atol=1e-6)
    tot = np.load(tmp_path / "stats.npy")
    assert tot[0] == st["camera_samples"] and tot[1] == st["rays_closest"] and tot[2] == st["rays_any"]


def test_shard_helper():
    from fountain_b200.distributed import shard
    assert shard(0, 1) == (0, 1) and shard(3, 8) == (3, 8)
    covered = sorted(s for r in range(4) for s in range(r, 10, 4))
    assert covered == list(range(10))
