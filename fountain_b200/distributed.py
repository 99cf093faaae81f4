"""Multi-GPU host logic: one process per GPU, replicated scene, sample-index
sharding, ONE sum-reduction of the film -- the analogue of `merge_film_tile`'s mutex merge
(film.rs:121-132).  `torch.distributed` is plumbing only (NCCL over NVLink on GPUs, gloo in the
CPU tests)."""
import ctypes as C

import numpy as np

from . import _abi as A


def shard(rank, world):
    """(sample_begin, sample_stride) of a rank: it renders samples rank, rank + world, ..."""
    if not (0 <= rank < world):
        raise ValueError("bad rank %d of %d" % (rank, world))
    return rank, world


def render_sharded(integrator, scene, film, sampler, rank=None, world=None, dst=0):
    """SamplerIntegrator.render_parallel across the ranks of the default process group.

    Every rank must hold the same scene / camera / film / sampler.  After the call rank `dst`'s
    film.pixels holds the full image; other ranks hold their partial film.  Works with any
    backend: with the CUDA library the per-rank film stays on the device and is reduced with
    NCCL; otherwise (CPU test tier) the host film is reduced with the group's backend."""
    import torch
    import torch.distributed as dist

    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    begin, stride = shard(rank, world)
    be = scene.backend
    n_px = film.width * film.height
    if be.has("render_device") and torch.cuda.is_available():
        dev = torch.device("cuda", torch.cuda.current_device())
        d_film = torch.zeros((n_px, 4), dtype=torch.float32, device=dev)
        st = A.FtnStats()
        cam, f, it = integrator.camera.to_abi(), film.to_abi(), integrator.radiance.to_abi()
        s = sampler.to_abi(begin, stride)
        stream = torch.cuda.current_stream()
        be.call("render_device", scene.handle, C.byref(cam), C.byref(f), C.byref(s), C.byref(it),
                C.c_void_p(d_film.data_ptr()), C.byref(st), C.c_void_p(stream.cuda_stream))
        if world > 1:
            dist.reduce(d_film, dst=dst, op=dist.ReduceOp.SUM)
        film.pixels = d_film.cpu().numpy().reshape(film.height, film.width, 4)
        stats = st.as_dict()
    else:
        stats = integrator.render_parallel(scene, film, sampler, sample_begin=begin, sample_stride=stride)
        if world > 1:
            t = torch.from_numpy(np.ascontiguousarray(film.pixels))
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
            film.pixels = t.numpy()
    if world > 1:
        tot = torch.tensor([stats["camera_samples"], stats["rays_closest"], stats["rays_any"]], dtype=torch.float64)
        if be.has("render_device") and torch.cuda.is_available():
            tot = tot.cuda()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        stats["camera_samples"], stats["rays_closest"], stats["rays_any"] = (int(x) for x in tot.tolist())
    return stats
