"""Experiment: how much faster does the traversal kernel run when an incoherent ray batch is
reordered first?  Sorting here is done out of band (torch.sort) -- only the trace is timed -- to
size the potential before writing a device-side reorder stage."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from fountain_b200 import api
from workloads import scenes  # noqa: E402


def expand_bits(v):
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def morton(q):   # q: (n,3) int64 in [0,1023]
    return (expand_bits(q[:, 0]) << 2) | (expand_bits(q[:, 1]) << 1) | expand_bits(q[:, 2])


def keys_for(rays, lo, hi, mode):
    o = rays[:, 0:3].double(); d = rays[:, 3:6].double()
    q = ((o - lo) / (hi - lo)).clamp(0, 1 - 1e-9)
    octant = ((d[:, 0] < 0).long() << 2) | ((d[:, 1] < 0).long() << 1) | (d[:, 2] < 0).long()
    dn = d / d.norm(dim=1, keepdim=True)
    if mode == "origin30":
        return morton((q * 1024).long())
    if mode == "octant_origin":
        return (octant << 30) | morton((q * 1024).long())
    if mode == "origin15_dir15":
        dq = ((dn + 1) * 0.5).clamp(0, 1 - 1e-9)
        return ((morton((q * 1024).long()) >> 15) << 15) | (morton((dq * 1024).long()) >> 15)
    if mode == "origin9_dir12":
        dq = ((dn + 1) * 0.5).clamp(0, 1 - 1e-9)
        return ((morton((q * 1024).long()) >> 21) << 12) | (morton((dq * 1024).long()) >> 18)
    if mode == "dir15_origin15":
        dq = ((dn + 1) * 0.5).clamp(0, 1 - 1e-9)
        return ((morton((dq * 1024).long()) >> 15) << 15) | (morton((q * 1024).long()) >> 15)
    if mode == "origin18_dir6":
        dq = ((dn + 1) * 0.5).clamp(0, 1 - 1e-9)
        return ((morton((q * 1024).long()) >> 12) << 6) | (morton((dq * 1024).long()) >> 24)
    raise ValueError(mode)


def main():
    dev = torch.device("cuda", 0)
    gpu = api.default_backend()
    gpu.call("set_device", 0)
    tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    n_lon = int(round(tris ** 0.5))
    scene, camera = scenes.synthetic_mesh_scene(n_lon, n_lon // 2, backend=gpu, resolution=(2048, 2048))
    prim = scenes.primary_ray_batch(camera, (2048, 2048))
    hits = scene.intersect(prim)
    inc = scenes.diffuse_bounce_batch(prim, hits, scene._positions, scene._indices, seed=2)
    reps = int(np.ceil((4 << 20) / max(1, len(inc))))
    inc = np.concatenate([inc] * reps)[: 4 << 20] if len(inc) < (4 << 20) else inc
    rng = np.random.default_rng(4)
    n_int = 4 << 20
    o = rng.normal(size=(n_int, 3)); o *= (rng.random((n_int, 1)) ** (1 / 3) * 9.0) / np.linalg.norm(o, axis=1, keepdims=True)
    d = rng.normal(size=(n_int, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    interior = api.make_rays(o.astype(np.float32), d.astype(np.float32))
    lo_np, hi_np = scene.world_bound()
    lo = torch.tensor(lo_np, dtype=torch.float64, device=dev); hi = torch.tensor(hi_np, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}

    def time_trace(d_rays):
        n = d_rays.shape[0]
        d_hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            gpu.call("intersect_device", scene.handle, n, C.c_void_p(d_rays.data_ptr()), C.c_void_p(d_hits.data_ptr()), C.c_void_p(stream.cuda_stream))
            e1.record(stream); e1.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return n / (float(np.median(ts)) * 1e-3) / 1e6

    for label, batch in (("incoherent_diffuse", inc), ("incoherent_interior", interior), ("coherent_primary", prim)):
        n = len(batch)
        d_rays = torch.from_numpy(batch.view(np.float32).reshape(n, 8)).to(dev)
        res = {"unsorted": time_trace(d_rays)}
        perm = torch.randperm(n, device=dev)
        res["shuffled"] = time_trace(d_rays[perm].contiguous())
        for mode in ("origin30", "octant_origin", "origin15_dir15", "origin9_dir12", "dir15_origin15", "origin18_dir6"):
            k = keys_for(d_rays, lo, hi, mode)
            order = torch.sort(k, stable=True)[1]
            res[mode] = time_trace(d_rays[order].contiguous())
        out[label] = res
        print(label, json.dumps({k: round(v, 1) for k, v in res.items()}), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/exp_sort.json", "w"), indent=1)


if __name__ == "__main__":
    main()
