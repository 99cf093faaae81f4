"""`-m gpu` tier, N > 1: the two multi-GPU forms of the render -- ftn_render_multi (one process, N devices, NCCL
inside the library) and the one-process-per-GPU torchrun form (fountain_b200/distributed.py) -- against the single-GPU
film of the same samples (film.rs:121-132: every sample lands exactly once, whatever the split).  Skipped on a
one-GPU box; the round's bench repeats the check at every N (`film_check` in bench.py's line).  Also here: the
boundary's concurrency contract (queries in flight on two streams, renders from two host threads)."""
import ctypes as C
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from fountain_b200 import _abi as A
from fountain_b200 import api
from workloads import scenes
from tests import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count(gpu_backend):
    n = C.c_int(0)
    gpu_backend.call("device_count", C.byref(n))
    return n.value


def _cube_on(gpu_backend, device, resolution=(96, 96)):
    mesh = api.TriangleMesh.from_ply(scenes.ROUNDED_CUBE_PLY)
    prim = api.GeometricPrimitive(mesh, api.MatteMaterial(0.5))
    scene = api.Scene([prim], [api.InfiniteAreaLight.new_uniform(1.0)], backend=gpu_backend, device=device)
    _, camera, film = scenes.rounded_cube_scene(backend=gpu_backend, resolution=resolution)
    return scene, camera, film


def test_render_multi_with_one_scene_is_render(gpu_backend):
    scene, camera, film = _cube_on(gpu_backend, 0)
    integ = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0))
    sampler = api.RandomSampler.new_with_seed(6, 5)
    st1 = integ.render_multi([scene], film, sampler)
    multi = film.pixels.copy()
    st2 = integ.render_parallel(scene, film, sampler)
    assert np.array_equal(multi, film.pixels)
    assert st1["rays_closest"] == st2["rays_closest"] and st1["rays_any"] == st2["rays_any"]


@pytest.mark.parametrize("n_dev", [2, 4, 8])
def test_render_multi_matches_single_gpu(gpu_backend, n_dev):
    if _device_count(gpu_backend) < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    built = [_cube_on(gpu_backend, d) for d in range(n_dev)]
    gpu_backend.call("set_device", 0)
    scene_list = [b[0] for b in built]
    _, camera, film = built[0]
    integ = api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0))
    sampler = api.RandomSampler.new_with_seed(16, 3)
    st = integ.render_multi(scene_list, film, sampler)
    multi = film.pixels.copy()
    st1 = integ.render_parallel(scene_list[0], film, sampler)
    assert np.array_equal(multi[..., 3], film.pixels[..., 3])                      # every sample exactly once
    assert np.allclose(multi[..., :3], film.pixels[..., :3], rtol=1e-5, atol=1e-6)  # sums in a different order
    assert st["camera_samples"] == st1["camera_samples"]
    assert st["rays_closest"] == st1["rays_closest"] and st["rays_any"] == st1["rays_any"]
    for s in scene_list:
        s.close()


def test_torchrun_sharded_render_matches_single_gpu(gpu_backend):
    """The one-process-per-GPU form bench.py uses: scripts/multi_gpu_check.py asserts weights == and colours
    allclose(1e-5) between the NCCL-reduced film and rank 0's own full render."""
    if _device_count(gpu_backend) < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "scripts", "multi_gpu_check.py")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "weights_equal=True colours_close=True" in r.stdout


def test_queries_in_flight_on_two_streams(gpu_backend):
    """include/fountain_gpu.h: ftn_intersect*_device only enqueue, and queries on one scene may overlap on different
    streams (each call owns a work counter).  Two closest-hit batches and an any-hit batch enqueued back to back on
    three streams give what the same calls give one at a time."""
    import torch
    scene, _ = scenes.synthetic_mesh_scene(400, 200, backend=gpu_backend, resolution=(64, 64))
    dev = torch.device("cuda", 0)
    batches = [parity.random_ray_batch(400_000, 31 + i, extent=10.0, far=40.0) for i in range(3)]
    d_rays = [torch.from_numpy(b.view(np.float32).reshape(-1, 8)).to(dev) for b in batches]
    ref = [scene.intersect(batches[0]), scene.intersect(batches[1]), scene.intersect_test(batches[2])]
    for _ in range(3):
        d_hits = [torch.zeros((len(batches[0]), 4), dtype=torch.float32, device=dev) for _ in range(2)]
        d_any = torch.zeros(len(batches[2]), dtype=torch.uint8, device=dev)
        streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
        torch.cuda.synchronize()
        gpu_backend.call("intersect_device", scene.handle, len(batches[0]), C.c_void_p(d_rays[0].data_ptr()), C.c_void_p(d_hits[0].data_ptr()), C.c_void_p(streams[0].cuda_stream))
        gpu_backend.call("intersect_device", scene.handle, len(batches[1]), C.c_void_p(d_rays[1].data_ptr()), C.c_void_p(d_hits[1].data_ptr()), C.c_void_p(streams[1].cuda_stream))
        gpu_backend.call("intersect_test_device", scene.handle, len(batches[2]), C.c_void_p(d_rays[2].data_ptr()), C.c_void_p(d_any.data_ptr()), C.c_void_p(streams[2].cuda_stream))
        torch.cuda.synchronize()
        for k in range(2):
            got = np.frombuffer(d_hits[k].cpu().numpy().tobytes(), dtype=api.HIT_DTYPE)
            assert np.array_equal(got.view(np.uint32), ref[k].view(np.uint32))
        assert np.array_equal(d_any.cpu().numpy().astype(bool), ref[2])


def test_renders_from_two_host_threads(gpu_backend):
    """the film slot of ftn_render is shared per device; two host threads rendering films of different
    sizes at the same time must each get their own image (the slot is held for the whole call)."""
    scene, camera_s, film_s = _cube_on(gpu_backend, 0, resolution=(64, 64))
    _, camera_l, film_l = scenes.rounded_cube_scene(backend=gpu_backend, resolution=(192, 160))
    sampler = api.RandomSampler.new_with_seed(4, 9)
    ref_s = api.Film((64, 64), backend=gpu_backend); ref_l = api.Film((192, 160), backend=gpu_backend)
    api.SamplerIntegrator(camera_s, api.PathIntegrator(5, 1.0)).render_parallel(scene, ref_s, sampler)
    api.SamplerIntegrator(camera_l, api.PathIntegrator(5, 1.0)).render_parallel(scene, ref_l, sampler)
    errors = []

    def work(camera, film, ref):
        try:
            for _ in range(6):
                api.SamplerIntegrator(camera, api.PathIntegrator(5, 1.0)).render_parallel(scene, film, sampler)
                if not np.array_equal(film.pixels, ref.pixels):
                    errors.append("film differs")
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(camera_s, film_s, ref_s)), threading.Thread(target=work, args=(camera_l, film_l, ref_l))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
