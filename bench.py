#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native path (BASELINE.json metric: Mrays/s incl.
secondary rays, and camera samples/s, per scene, next to the host-CPU figure).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c2|c3|c5]

A "step" is one pass of the hot path over one batch of synthetic input:
  c4 (default, BASELINE configs[3], the config the metric's 8-GPU number is quoted on): one full render of the
      logo-style scene (thin lens, Trowbridge-Reitz copper mesh, 2048x1024 image env light), 1920x1080, a FIXED
      TOTAL of 1024 spp, path depth 5 -- split over the N GPUs by sample index (STRONG scaling: N = 8 renders 128
      spp per GPU).  At N = 1 the same line also carries C2 and the three C3 ray batches as extra keys.
  c2 (BASELINE configs[1]): data/rounded_cube.ply, Lambert, uniform env light, 512x512, 64 spp per GPU, depth 5.
  c3 (configs[2]): closest-hit ray batches on the 1M-triangle mesh (coherent primary / incoherent diffuse-bounce /
      interior); value = the incoherent diffuse-bounce batch.
  c5 (configs[4]): large synthetic mesh (--tris) at 3840x2160, --spp per GPU (weak scaling).
Prints ONE JSON line (rank 0).  `value` = whole-job Mrays/s with the scene resident in HBM; `e2e` = the same
metric through the host-buffer API (scene upload + BVH build + render + film read-back inside the timed region).
N > 1: one process per GPU (torchrun), replicated scene, sample-index sharding, one NCCL reduction of the film;
after the timed loop the NCCL-reduced film is CHECKED against a single-GPU render of the same samples (`film_check`).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fountain_b200 import _abi as A          # noqa: E402
from fountain_b200 import api                # noqa: E402
from workloads import scenes                 # noqa: E402

METRIC = "Mrays/s (closest-hit + any-hit + MIS rays)"
UNIT = "Mrays/s"
STRONG = ("c4",)            # workloads with a fixed total spp split over the GPUs
DEFAULT_SPP = {"c2": 64, "c4": 1024, "c5": 4}
SM_COUNT = 148
L1_BYTES_PER_CLK = 128      # L1 data pipe width per SM (B300_MICROARCH: smem / L1 crossbar 128 B/cyc/SM)
L2_BYTES_PER_CLK = 6300     # LTS throughput cap, full chip (B300_MICROARCH)


# ---------------------------------------------------------------------------------------------------
def workload_scene(name, backend, args, resolution=None):
    """-> (scene, camera, film, integrator, spp, description)"""
    spp = args.spp or DEFAULT_SPP.get(name, 1)
    if name == "c2":
        res = resolution or (512, 512)
        scene, camera, film = scenes.rounded_cube_scene(backend=backend, resolution=res)
        return scene, camera, film, api.PathIntegrator(5, 1.0), spp, \
            "C2 rounded_cube.ply (4332 tris) Lambert Kd .5, uniform env L=1, %dx%d, %d spp, path depth 5, rr 1.0" % (res[0], res[1], spp)
    if name == "c4":
        res = resolution or (1920, 1080)
        scene, camera, film = scenes.logo_style_scene(backend=backend, resolution=res, detail=args.detail)
        return scene, camera, film, api.PathIntegrator(5, 1.0), spp, \
            "C4 logo-style gear ring (%d tris) TR copper r=.01, thin lens, 2048x1024 sky+sun env, %dx%d, %d spp, depth 5" % (scene.n_triangles, res[0], res[1], spp)
    if name == "c5":
        res = resolution or (3840, 2160)
        n_lon = int(round(args.tris ** 0.5))
        scene, camera = scenes.synthetic_mesh_scene(n_lon, n_lon // 2, backend=backend, resolution=res)
        film = api.Film(res, backend=backend)
        return scene, camera, film, api.PathIntegrator(5, 1.0), spp, \
            "C5 displaced sphere (%d tris) Lambert, uniform env, %dx%d, %d spp, depth 5" % (scene.n_triangles, res[0], res[1], spp)
    raise SystemExit("unknown render workload %r" % name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for t, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def git_commit():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=5).stdout.strip() or None
    except Exception:
        return None


def ncu_record(workload, kernel):
    """What ncu measured for `kernel` of `workload` (profiles/traffic.json: one entry per workload, stamped with the
    commit / build it was taken on): DRAM bytes per launch and the pipe utilisations that name the limiter.  The
    entry is returned with its own provenance; a workload without an entry gets None -- never another workload's."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(workload, {}).get(kernel)


def hierarchy_roofline(workload, kernel, bytes_per_launch, avg_launch_s, sm_mhz, extra, rays_per_launch=None):
    """Roofline of a traversal kernel from the live CUDA-event launch time.

    Which roof: ncu (profiles/, one record per workload and kernel in profiles/traffic.json, stamped with the commit it was
    taken on) says the BVH8q kernels are bound by the SM's instruction issue -- the ALU pipe inside it -- and the BVH2x64
    kernels by the L1 data pipe; no scene measured is HBM-bound (DRAM <= 15 % of peak).  So when a record with the
    kernel's warp instructions per ray exists, `achieved` = those instructions x the rays one launch traces / the measured
    launch time, against the issue peak (SMs x 4 schedulers x SM clock sampled under load).  Without a record the L1
    data pipe stands in: ALGORITHMIC bytes (node + triangle + ray / hit bytes the rays request) per launch time against
    SMs x 128 B/clk.  The L1, L2 and HBM figures of the same bytes always stand beside the stated roof; `traffic` is the DRAM
    traffic ncu measured for this workload (None when there is no record -- never another workload's)."""
    peak_hbm, peak_src = measured_peak()
    clk = (sm_mhz or 1965.0) * 1e6
    achieved_b = bytes_per_launch / max(avg_launch_s, 1e-12) / 1e9
    l1_peak = SM_COUNT * L1_BYTES_PER_CLK * clk / 1e9
    l2_peak = L2_BYTES_PER_CLK * clk / 1e9
    rec = ncu_record(workload, kernel)
    mem = {"l1": {"achieved": achieved_b, "peak": l1_peak, "frac": achieved_b / l1_peak, "unit": "GB/s",
                  "peak_source": "L1 data pipe: %d SMs x %d B/clk x %.0f MHz" % (SM_COUNT, L1_BYTES_PER_CLK, clk / 1e6)},
           "l2": {"peak": l2_peak, "frac": achieved_b / l2_peak, "peak_source": "LTS cap %d B/clk x SM clock (B300_MICROARCH)" % L2_BYTES_PER_CLK},
           "hbm": {"peak": peak_hbm, "frac_if_all_bytes_came_from_hbm": achieved_b / peak_hbm, "peak_source": peak_src,
                   "note": "algorithmic bytes / HBM peak; > 1 only says the working set is cache-resident -- the measured DRAM traffic is `traffic`"}}
    if rec and rec.get("warp_inst_per_ray") and rays_per_launch:
        issue_peak = SM_COUNT * 4 * clk / 1e9
        achieved = rec["warp_inst_per_ray"] * rays_per_launch / max(avg_launch_s, 1e-12) / 1e9
        out = {"bound": rec.get("bound", "sm-issue"), "kernel": kernel, "achieved": achieved, "peak": issue_peak, "unit": "Gwarp-inst/s",
               "frac": achieved / issue_peak,
               "peak_source": "instruction issue: %d SMs x 4 schedulers x %.0f MHz (SM clock sampled under load); warp instructions per ray from ncu (%s)"
                              % (SM_COUNT, clk / 1e6, rec.get("source")),
               "traffic": rec["dram_bytes_per_ray"] * rays_per_launch if rec.get("dram_bytes_per_ray") is not None else None}
    else:
        out = {"bound": "l1", "kernel": kernel, "achieved": achieved_b, "peak": l1_peak, "unit": "GB/s", "frac": achieved_b / l1_peak,
               "peak_source": mem["l1"]["peak_source"] + " (SM clock sampled under load)",
               "traffic": rec["dram_bytes_per_ray"] * rays_per_launch if (rec and rays_per_launch and rec.get("dram_bytes_per_ray") is not None) else None}
    out.update(mem)
    out.update({"algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_launch_s * 1e3, "rays_per_launch": rays_per_launch, "ncu": rec})
    out.update(extra)
    return out


# ---------------------------------------------------------------------------------------------------
def cpu_baseline(name, args, budget_s=15.0):
    """The oracle (CPU restatement of the reference algorithm) on the host cores, on a bounded
    sample of the same workload: the same scene / camera / resolution at reduced spp."""
    from oracle import orc
    be = orc.backend()
    cores = orc.hardware_threads()
    orc.set_threads(cores)
    scene, camera, film, integ, spp, desc = workload_scene(name, be, args)
    runner = api.SamplerIntegrator(camera, integ)
    t = time.perf_counter()
    st = runner.render_parallel(scene, film, api.RandomSampler.new_with_seed(1, 0, mode=A.FTN_SAMPLER_REFERENCE_TILE_STREAM))
    t1 = time.perf_counter() - t
    n = max(1, min(spp, int(budget_s / max(t1, 1e-3))))
    t = time.perf_counter()
    st = runner.render_parallel(scene, film, api.RandomSampler.new_with_seed(n, 0, mode=A.FTN_SAMPLER_REFERENCE_TILE_STREAM))
    dt = time.perf_counter() - t
    rays = st["rays_closest"] + st["rays_any"]
    return {"value": rays / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "samples_per_s": st["camera_samples"] / dt, "seconds": dt, "bvh_build_seconds": scene.stats()["bvh_build_seconds"],
            "sample": "%s at %d of %d spp, reference tile-stream sampler, %d threads (C++ restatement of the reference; Rust toolchain absent)"
                      % (desc.split(",")[0], n, spp, cores)}


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port) on the host
    cores; each step = BVH build + render of a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    be = orc.backend()
    cores = orc.hardware_threads()
    orc.set_threads(cores)
    name = args.workload
    if name == "c3":
        return run_reference_raybatch(args, be, cores)
    scene, camera, film, integ, spp, desc = workload_scene(name, be, args)
    runner = api.SamplerIntegrator(camera, integ)
    t = time.perf_counter()
    runner.render_parallel(scene, film, api.RandomSampler.new_with_seed(1, 0, mode=A.FTN_SAMPLER_REFERENCE_TILE_STREAM))
    t1 = time.perf_counter() - t
    total = args.steps + args.warmup
    n = max(1, min(spp, int(120.0 / total / max(t1, 1e-3))))   # whole run within ~2 minutes
    rays = samples = 0
    secs = 0.0
    for i in range(total):
        t = time.perf_counter()
        scene, camera, film, integ, _, _ = workload_scene(name, be, args)      # scene build + BVH::build
        st = api.SamplerIntegrator(camera, integ).render_parallel(scene, film, api.RandomSampler.new_with_seed(n, i, mode=A.FTN_SAMPLER_REFERENCE_TILE_STREAM))
        dt = time.perf_counter() - t
        if i >= args.warmup:
            secs += dt; rays += st["rays_closest"] + st["rays_any"]; samples += st["camera_samples"]
    v = rays / secs / 1e6
    sample = "%s at %d of %d spp per step (BVH build + render), %d host threads" % (desc.split(",")[0], n, spp, cores)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong" if name in STRONG else "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": {"workload": desc, "sample": sample},
           "samples_per_s": samples / secs,
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def run_reference_raybatch(args, be, cores):
    scene, camera, batches = c3_batches(be, args, want=("incoherent_diffuse",))
    inc = batches["incoherent_diffuse"]
    n_s = 1 << 17
    scene.intersect(inc[:4096])
    t0 = time.perf_counter(); scene.intersect(inc[:n_s]); dt = time.perf_counter() - t0
    n_s = int(min(len(inc), max(n_s, n_s * (60.0 / (args.steps + args.warmup)) / max(dt, 1e-3))))
    secs = 0.0
    for i in range(args.steps + args.warmup):
        t0 = time.perf_counter(); scene.intersect(inc[:n_s]); dt = time.perf_counter() - t0
        if i >= args.warmup:
            secs += dt
    v = n_s * args.steps / secs / 1e6
    sample = "closest hit for the first %d of the %d incoherent diffuse-bounce rays per step, %d host threads" % (n_s, len(inc), cores)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": c3_description(scene, len(inc)), "sample": sample},
                      "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing of one rank (NCCL over NVLink); world == 1 needs none of it."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL's version banner (NCCL_DEBUG=VERSION on the box) goes to stdout by default; stdout carries exactly one JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.gpu = api.default_backend()
        self.gpu.call("set_device", self.local_rank)
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()

    def max_f(self, v):
        if not self.dist:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_i(self, vals):
        if not self.dist:
            return [int(v) for v in vals]
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) for x in t.tolist()]


class ShardedRender:
    """One rank's share of a render: samples rank, rank + world, ... of `spp_total`, film on the device, ONE NCCL
    sum onto rank 0 (merge_film_tile across GPUs), Film::into_spectrum_buffer on rank 0."""

    def __init__(self, D, scene, camera, film, integ, spp_total, seed=0):
        torch = D.torch
        self.D, self.scene = D, scene
        self.cam, self.film, self.integ = camera.to_abi(), film.to_abi(), integ.to_abi()
        self.smp = api.RandomSampler.new_with_seed(spp_total, seed).to_abi(sample_begin=D.rank, sample_stride=D.world)
        self.n_px = film.width * film.height
        self.d_film = torch.zeros((self.n_px, 4), dtype=torch.float32, device=D.dev)
        self.d_rgb = torch.empty((self.n_px, 3), dtype=torch.float32, device=D.dev)

    def step(self, stats, scene=None, reduce=True, to_rgb=True):
        D = self.D
        self.d_film.zero_()
        D.gpu.call("render_device", (scene or self.scene).handle, C.byref(self.cam), C.byref(self.film), C.byref(self.smp), C.byref(self.integ),
                   C.c_void_p(self.d_film.data_ptr()), C.byref(stats), C.c_void_p(D.stream.cuda_stream))
        if D.dist and reduce:
            D.dist.reduce(self.d_film, dst=0, op=D.dist.ReduceOp.SUM)
        if D.rank == 0 and to_rgb:
            D.gpu.call("film_to_rgb_device", self.n_px, C.c_void_p(self.d_film.data_ptr()), C.c_void_p(self.d_rgb.data_ptr()), C.c_void_p(D.stream.cuda_stream))


def measure_render(D, name, args, steps, warmup, with_e2e=True, with_counts=True, sample_clocks=True):
    """Timed loop of one render workload on this rank's GPU (all ranks call it together).  Returns the pieces of the
    JSON line (rank 0 uses them)."""
    torch = D.torch
    scene, camera, film, integ, spp, desc = workload_scene(name, D.gpu, args)
    strong = name in STRONG
    spp_total = spp if strong else spp * D.world
    R = ShardedRender(D, scene, camera, film, integ, spp_total)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=D.dev)     # > 126 MB L2
    st = A.FtnStats()
    for _ in range(warmup):
        st.flags = A.FTN_STATS_TIME_KERNELS
        R.step(st)
    D.barrier()
    clocks = ClockSampler(D.local_rank) if (D.rank == 0 and sample_clocks) else None
    launches0 = int(D.gpu.fn["kernel_launch_count"]())
    ms_total = 0.0
    rays = samples = 0
    tsec = [0.0, 0.0, 0.0]; tl = [0, 0, 0]; trays = [0, 0, 0]; shade_s = 0.0; shade_l = 0
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        flush.zero_()                                                  # L2 flush between timed iterations
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(D.stream)
        st.flags = A.FTN_STATS_TIME_KERNELS                            # CUDA-event pairs around every traversal / shading launch
        R.step(st)
        e1.record(D.stream)
        e1.synchronize()
        ms_total += D.max_f(e0.elapsed_time(e1))                       # max over ranks, per step
        rays += st.rays_closest + st.rays_any
        samples += st.camera_samples
        for c in range(3):
            tsec[c] += st.trace_seconds[c]; tl[c] += st.trace_launches[c]; trays[c] += st.trace_rays[c]
        shade_s += st.shade_seconds; shade_l += st.shade_launches
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    launches = int(D.gpu.fn["kernel_launch_count"]()) - launches0
    clock_info = clocks.stop(t_wall0, t_wall1) if clocks else None
    rays_all, samples_all, launches_all = D.sum_i([rays, samples, launches])
    value = rays_all / (ms_total * 1e-3) / 1e6
    out = {"desc": desc, "spp": spp, "spp_total": spp_total, "value": value, "ms_per_step": ms_total / steps, "rays_per_step": rays_all // steps,
           "samples_per_s": samples_all / (ms_total * 1e-3), "launches": launches_all, "clocks": clock_info,
           "scene": scene, "camera": camera, "film": film, "integ": integ, "render": R}

    # ---- traversal counts for the algorithmic-bytes roofline (untimed counting render of this rank's samples) ----
    if with_counts:
        cst = A.FtnStats(); cst.flags = A.FTN_STATS_COUNT_TRAVERSAL
        R.step(cst, reduce=False, to_rgb=False)
        torch.cuda.synchronize()
        node_b, tri_b = cst.bvh_node_bytes, cst.bvh_tri_bytes
        k = 0                                                             # dominant kernel: k_extend (closest hit)
        bytes_per_step = node_b * cst.trace_nodes[k] + tri_b * cst.trace_tris[k] + 48 * cst.trace_rays[k]
        launches_per_step = max(1, tl[k] // steps)
        avg_launch_s = tsec[k] / max(1, tl[k])
        step_s = ms_total * 1e-3
        out["roofline"] = hierarchy_roofline(name, "k_extend", bytes_per_step / launches_per_step, avg_launch_s,
                                             clock_info["sm_mhz"] if clock_info else None, {
            "launches_per_step": launches_per_step, "bytes_per_ray": bytes_per_step / max(1, cst.trace_rays[k]),
            "nodes_per_ray": cst.trace_nodes[k] / max(1, cst.trace_rays[k]), "tris_per_ray": cst.trace_tris[k] / max(1, cst.trace_rays[k]),
            "node_bytes": node_b, "tri_bytes": tri_b, "kernel_share_of_step": tsec[k] / step_s,
            "all_traversal_share_of_step": sum(tsec) / step_s, "shade_share_of_step": shade_s / step_s,
            "kernel_mrays_per_s": trays[k] / max(tsec[k], 1e-12) / 1e6,
            "shadow_mrays_per_s": trays[1] / max(tsec[1], 1e-12) / 1e6, "mis_mrays_per_s": trays[2] / max(tsec[2], 1e-12) / 1e6,
            "shade_avg_launch_ms": shade_s / max(1, shade_l) * 1e3,
            "kernel_rays_per_step": [trays[0] // steps, trays[1] // steps, trays[2] // steps],
            "working_set_mb": (cst.bvh_nodes * node_b + scene.n_triangles * tri_b) / 1e6},
            rays_per_launch=trays[k] / max(1, tl[k]))

    # ---- e2e: host buffers in, host film out; scene upload + BVH build + render + read-back per step ----
    if with_e2e:
        e2e_samples = []          # (seconds, rays) per e2e step; the median step is reported
        pinned_film = None
        h2d = d2h = 0
        long_step = ms_total / steps > 500.0            # seconds-long steps: fewer repeats keep the default run within minutes
        n_e2e = 3 if long_step else max(3, min(steps, 5))
        n_e2e_warm = 1 if long_step else min(warmup, 3)  # untimed: first-use costs (pinned pools, allocator growth) are not steady state
        sampler = api.RandomSampler.new_with_seed(spp_total, 0)
        for i in range(-n_e2e_warm, n_e2e):
            D.barrier()
            t0 = time.perf_counter()
            sc2, cam2, film2, integ2, _, _ = workload_scene(name, D.gpu, args)          # host arrays -> ftn_scene_create (H2D) + ftn_bvh_build
            t_build = time.perf_counter() - t0
            if D.world == 1:
                st2 = api.SamplerIntegrator(cam2, integ2).render_parallel(sc2, film2, sampler)   # ftn_render: film D2H inside
                r = st2["rays_closest"] + st2["rays_any"]
                d2h = film2.pixels.nbytes
            else:
                s2 = A.FtnStats()
                R.step(s2, scene=sc2, to_rgb=False)
                if D.rank == 0:
                    if pinned_film is None:
                        pinned_film = torch.empty(R.d_film.shape, dtype=R.d_film.dtype, pin_memory=True)
                    pinned_film.copy_(R.d_film)            # D2H into page-locked memory
                r = s2.rays_closest + s2.rays_any
                d2h = (R.d_film.numel() * 4) if D.rank == 0 else 0
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if D.rank == 0:
                print("e2e %s step %d: scene+build %.2f ms, total %.2f ms" % (name, i, t_build * 1e3, dt * 1e3), file=sys.stderr)
            h2d = sc2.upload_bytes
            sc2.close()
            dt = D.max_f(dt)
            r = D.sum_i([r])[0]
            if i >= 0:
                e2e_samples.append((dt, r))
        med = sorted(e2e_samples)[len(e2e_samples) // 2]
        out["e2e"] = {"value": med[1] / med[0] / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                      "ms_per_step": med[0] * 1e3, "ms_all_steps": [round(x[0] * 1e3, 3) for x in e2e_samples], "warmup_steps": n_e2e_warm,
                      "includes": "ftn_scene_create (host mesh arrays -> H2D) + ftn_bvh_build + ftn_render + film D2H; PLY text parsing excluded (stays on the host side of the ABI)"}
    del flush
    return out


def film_check(D, name, args, spp_total):
    """N > 1: the NCCL-reduced film of a sharded render == rank 0's own render of ALL the samples (film.rs:121-132:
    every sample lands exactly once) -- weights equal, colours allclose(rtol 1e-5); at reduced resolution so that
    rank 0's full render stays short."""
    torch = D.torch
    res = {"c2": (256, 256), "c4": (240, 135), "c5": (240, 135)}[name]
    spp_c = min(spp_total, 64 * D.world)
    scene, camera, film, integ, _, _ = workload_scene(name, D.gpu, args, resolution=res)
    R = ShardedRender(D, scene, camera, film, integ, spp_c, seed=5)
    st = A.FtnStats()
    R.step(st, to_rgb=False)
    torch.cuda.synchronize()
    ok = None
    if D.rank == 0:
        sharded = R.d_film.cpu().numpy().reshape(film.height, film.width, 4)
        api.SamplerIntegrator(camera, integ).render_parallel(scene, film, api.RandomSampler.new_with_seed(spp_c, 5))
        w_ok = bool(np.array_equal(sharded[..., 3], film.pixels[..., 3]))
        c_ok = bool(np.allclose(sharded[..., :3], film.pixels[..., :3], rtol=1e-5, atol=1e-6))
        ok = w_ok and c_ok
        print("film_check: %dx%d, %d spp over %d ranks: weights_equal=%s colours_close=%s" % (res[0], res[1], spp_c, D.world, w_ok, c_ok), file=sys.stderr)
    D.barrier()
    scene.close()
    return ok


# ---------------------------------------------------------------------------------------------------
def c3_description(scene, n_inc):
    return "C3 displaced sphere %d tris: closest-hit batches, 2048x2048 primary + %d incoherent diffuse-bounce rays" % (scene.n_triangles, n_inc)


def c3_batches(backend, args, want=("coherent_primary", "incoherent_diffuse", "incoherent_interior")):
    n_lon = int(round(args.tris ** 0.5))
    scene, camera = scenes.synthetic_mesh_scene(n_lon, n_lon // 2, backend=backend, resolution=(2048, 2048))
    prim = scenes.primary_ray_batch(camera, (2048, 2048))
    out = {}
    if "coherent_primary" in want:
        out["coherent_primary"] = prim
    if "incoherent_diffuse" in want:
        hits = scene.intersect(prim)          # any backend: the hits are bit-identical (tests/test_gpu_parity.py)
        inc = scenes.diffuse_bounce_batch(prim, hits, scene._positions, scene._indices, seed=2)
        reps = int(np.ceil((4 << 20) / max(1, len(inc))))
        out["incoherent_diffuse"] = np.concatenate([inc] * reps)[: 4 << 20] if len(inc) < (4 << 20) else inc
    if "incoherent_interior" in want:
        # harder than the spec's batch B (whose rays mostly leave the convex mesh): uniformly random origins INSIDE the
        # closed mesh with uniformly random directions -- every ray hits, none is coherent
        rng = np.random.default_rng(4)
        n_int = 4 << 20
        o = rng.normal(size=(n_int, 3)); o *= (rng.random((n_int, 1)) ** (1 / 3) * 9.0) / np.linalg.norm(o, axis=1, keepdims=True)
        d = rng.normal(size=(n_int, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        out["incoherent_interior"] = api.make_rays(o.astype(np.float32), d.astype(np.float32))
    return scene, camera, out


def measure_raybatch(D, args, steps, warmup, with_e2e=True, with_baseline=True, build_reps=5):
    """c3: closest-hit ray batches on the 1M-triangle mesh + the integer-work figures of the build
    (Morton codes + radix sort keys/s, one reconciled BVH build time)."""
    torch = D.torch
    gpu = D.gpu
    n_lon = int(round(args.tris ** 0.5)); side = (n_lon, n_lon // 2)          # n_lon * n_lat * 2 triangles
    # ---- build: first (cold: arena + pool growth, module load) and steady-state (median of `build_reps` re-creations) ----
    t0 = time.perf_counter()
    s0, _ = scenes.synthetic_mesh_scene(side[0], side[1], backend=gpu, resolution=(2048, 2048))
    first_wall = time.perf_counter() - t0
    first = s0.stats(); s0.close()
    builds, sorts, walls = [], [], []
    for _ in range(build_reps):
        t0 = time.perf_counter()
        s1, _ = scenes.synthetic_mesh_scene(side[0], side[1], backend=gpu, resolution=(2048, 2048))
        walls.append(time.perf_counter() - t0)
        stt = s1.stats(); builds.append(stt["bvh_build_seconds"]); sorts.append(stt["morton_sort_seconds"]); s1.close()
    scene, camera, batches = c3_batches(gpu, args)
    n_keys = scene.n_triangles
    build = {"triangles": n_keys, "bvh_build_ms": float(np.median(builds)) * 1e3, "bvh_build_ms_all": [round(b * 1e3, 3) for b in builds],
             "bvh_build_first_ms": first["bvh_build_seconds"] * 1e3, "scene_create_and_build_wall_ms": float(np.median(walls)) * 1e3,
             "scene_create_and_build_first_wall_ms": first_wall * 1e3,
             "morton_sort_ms": float(np.median(sorts)) * 1e3, "morton_sort_mkeys_per_s": n_keys / max(float(np.median(sorts)), 1e-12) / 1e6,
             "note": "bvh_build_ms = CUDA-event time of ftn_bvh_build (bounds, Morton codes, radix sort, PLOC topology, wide-node collapse, triangle "
                     "re-layout), median of %d re-creations of the same scene after a first (cold) one; morton_sort = k_morton + 4-pass 8-bit LSD "
                     "radix sort of (30-bit code, primitive) pairs inside it" % build_reps}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=D.dev)
    results = {}
    clock_info = None
    stream = D.stream
    d_hits_inc_check = None
    for label in ("coherent_primary", "incoherent_diffuse", "incoherent_interior"):
        batch = batches[label]
        n = len(batch)
        d_rays = torch.from_numpy(batch.view(np.float32).reshape(n, 8)).to(D.dev)
        d_hits = torch.empty((n, 4), dtype=torch.float32, device=D.dev)
        d_cnt = torch.zeros(2, dtype=torch.int64, device=D.dev)
        gpu.call("intersect_count_device", scene.handle, n, C.c_void_p(d_rays.data_ptr()), C.c_void_p(d_hits.data_ptr()),
                 C.c_void_p(d_cnt.data_ptr()), C.c_void_p(stream.cuda_stream))
        torch.cuda.synchronize()
        nodes, tris = (int(x) for x in d_cnt.tolist())
        for _ in range(warmup):
            gpu.call("intersect_device", scene.handle, n, C.c_void_p(d_rays.data_ptr()), C.c_void_p(d_hits.data_ptr()), C.c_void_p(stream.cuda_stream))
        times = []
        l0 = int(gpu.fn["kernel_launch_count"]())
        sample_clocks = label == "incoherent_diffuse" and D.rank == 0
        if sample_clocks:   # the timed steps are ~1 ms each: repeat them (untimed extras) until nvidia-smi has sampled the loaded GPU
            clocks = ClockSampler(D.local_rank)
            t_wall0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            gpu.call("intersect_device", scene.handle, n, C.c_void_p(d_rays.data_ptr()), C.c_void_p(d_hits.data_ptr()), C.c_void_p(stream.cuda_stream))
            e1.record(stream); e1.synchronize()
            times.append(e0.elapsed_time(e1))
        launches = int(gpu.fn["kernel_launch_count"]()) - l0
        if sample_clocks:
            while time.perf_counter() - t_wall0 < 0.3:
                gpu.call("intersect_device", scene.handle, n, C.c_void_p(d_rays.data_ptr()), C.c_void_p(d_hits.data_ptr()), C.c_void_p(stream.cuda_stream))
                torch.cuda.synchronize()
            clock_info = clocks.stop(t_wall0, time.perf_counter())
        med = float(np.median(times))
        st = scene.stats()
        bytes_ = st["bvh_node_bytes"] * nodes + st["bvh_tri_bytes"] * tris + 48 * n
        if label == "incoherent_diffuse":
            d_hits_inc_check = np.frombuffer(d_hits.cpu().numpy().tobytes(), dtype=api.HIT_DTYPE)
        results[label] = {"rays": n, "ms_median": med, "mrays_per_s": n / (med * 1e-3) / 1e6, "nodes_per_ray": nodes / n, "tris_per_ray": tris / n,
                          "bytes_per_ray": bytes_ / n, "algorithmic_bytes": bytes_, "launches": launches,
                          "hit_fraction": float((d_hits[:, 0].view(torch.int32) != -1).float().mean().item())}
        del d_rays, d_hits
    inc = batches["incoherent_diffuse"]
    out = {"scene": scene, "batches": results, "clocks": clock_info, "build": build, "n_inc": len(inc), "bvh_nodes": scene.stats()["bvh_nodes"]}
    sm_mhz = clock_info["sm_mhz"] if clock_info else None
    for label, r in results.items():
        rl = hierarchy_roofline("c3", "k_intersect_batch:" + label, r["algorithmic_bytes"], r["ms_median"] * 1e-3, sm_mhz, {}, rays_per_launch=r["rays"])
        r["achieved_gbs"] = rl["l1"]["achieved"]; r["frac_of_l1_peak"] = rl["l1"]["frac"]; r["frac_of_l2_peak"] = rl["l2"]["frac"]
        if rl["unit"] != "GB/s":
            r["issue_frac"] = rl["frac"]
        r["dram_traffic_bytes"] = rl["traffic"]
        if label == "incoherent_diffuse":
            out["roofline"] = rl
    if with_e2e:
        # e2e: host rays in, host hits out through ftn_intersect, both in page-locked host memory (ftn_host_alloc)
        n_inc = len(inc)
        p_rays, p_hits = A.VOIDP(), A.VOIDP()
        gpu.call("host_alloc", n_inc * 32, C.byref(p_rays))
        gpu.call("host_alloc", n_inc * 16, C.byref(p_hits))
        C.memmove(p_rays.value, inc.ctypes.data, n_inc * 32)
        e2e_times = []
        for i in range(5):
            t0 = time.perf_counter()
            gpu.call("intersect", scene.handle, n_inc, C.cast(p_rays, C.POINTER(A.FtnRay)), C.cast(p_hits, C.POINTER(A.FtnHit)))
            e2e_times.append(time.perf_counter() - t0)
        e2e_dt = float(np.median(e2e_times[1:]))
        host_hits = np.frombuffer((C.c_char * (n_inc * 16)).from_address(p_hits.value), dtype=api.HIT_DTYPE)
        assert np.array_equal(host_hits["t"], d_hits_inc_check["t"]) and np.array_equal(host_hits["prim"], d_hits_inc_check["prim"])   # same answers as the resident path
        gpu.call("host_free", p_rays); gpu.call("host_free", p_hits)
        out["e2e"] = {"value": n_inc / e2e_dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(inc.nbytes), "d2h_bytes_per_step": int(n_inc * 16)}
    if with_baseline:
        # the oracle's closest-hit query (reference BVH + traversal restated, all host threads) on the first rays of the same batch
        from oracle import orc
        cores = orc.hardware_threads(); orc.set_threads(cores)
        o_scene, _ = scenes.synthetic_mesh_scene(side[0], side[1], backend=orc.backend(), resolution=(2048, 2048))
        n_s = min(len(inc), 1 << 18)
        o_scene.intersect(inc[:4096])
        t0 = time.perf_counter(); o_hits = o_scene.intersect(inc[:n_s]); dt = time.perf_counter() - t0
        n_s2 = int(min(len(inc), max(n_s, n_s * 10.0 / max(dt, 1e-3))))
        if n_s2 > n_s:
            t0 = time.perf_counter(); o_hits = o_scene.intersect(inc[:n_s2]); dt = time.perf_counter() - t0; n_s = n_s2
        assert np.array_equal(o_hits["t"], d_hits_inc_check["t"][:n_s])   # and the device agrees with it bit for bit
        out["cpu_baseline"] = {"value": n_s / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt,
                               "bvh_build_seconds": o_scene.stats()["bvh_build_seconds"], "checked_bit_for_bit": True,
                               "sample": "closest hit for the first %d of the %d incoherent diffuse-bounce rays, %d threads (C++ restatement of the reference; Rust toolchain absent)" % (n_s, len(inc), cores)}
    del flush
    return out


BVH_NOTE = ("30-bit Morton codes + stable radix sort; topology: PLOC (>= 65536 tris) or Karras radix tree; records: BVH8q compressed 8-wide "
            "(>= 65536 tris) or BVH2x64; see DESIGN.md section 3")


def run_ours(args):
    D = Dist()
    name = args.workload
    if name == "c3":
        if D.rank != 0:          # the ray-batch query does not shard: replicas only; rank 0 reports
            return
        m = measure_raybatch(D, args, args.steps, args.warmup, with_baseline=not args.no_cpu_baseline)
        inc_r = m["batches"]["incoherent_diffuse"]
        print(json.dumps({"metric": METRIC, "value": inc_r["mrays_per_s"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": inc_r["ms_median"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": c3_description(m["scene"], m["n_inc"]), "l2": "flushed between timed steps (256 MiB write)", "bvh": BVH_NOTE},
                          "batches": m["batches"], "gpu_launches": inc_r["launches"], "clocks": m["clocks"], "cpu_baseline": m.get("cpu_baseline"),
                          "e2e": m.get("e2e"), "roofline": m["roofline"], "build": m["build"], "bvh_build_ms": m["build"]["bvh_build_ms"], "bvh_nodes": m["bvh_nodes"],
                          "commit": git_commit()}))
        return
    m = measure_render(D, name, args, args.steps, args.warmup, with_e2e=not args.no_e2e)
    check = film_check(D, name, args, m["spp_total"]) if D.world > 1 else None
    extras = {}
    if D.world == 1 and name == "c4" and not args.no_extras:
        # the other BASELINE configs that fit one GPU, as extra keys of the same line (short runs; the C2 / C3 lines of
        # their own come from --workload c2 / c3)
        c2 = measure_render(D, "c2", argparse.Namespace(**{**vars(args), "spp": 0}), max(3, args.steps), max(3, args.warmup), sample_clocks=False)
        c2["scene"].close()
        extras["c2"] = {"workload": c2["desc"], "value": c2["value"], "unit": UNIT, "ms_per_step": c2["ms_per_step"], "samples_per_s": c2["samples_per_s"],
                        "e2e": c2["e2e"], "roofline": c2["roofline"]}
        c3 = measure_raybatch(D, args, max(3, args.steps), max(3, args.warmup), with_e2e=False, with_baseline=False, build_reps=3)
        c3["scene"].close()
        extras["c3"] = {"workload": c3_description(c3["scene"], c3["n_inc"]), "unit": UNIT,
                        "batches": {k: {kk: v[kk] for kk in ("mrays_per_s", "ms_median", "nodes_per_ray", "tris_per_ray", "bytes_per_ray", "frac_of_l1_peak", "hit_fraction")}
                                    for k, v in c3["batches"].items()},
                        "build": c3["build"]}
    if D.rank == 0:
        base = cpu_baseline(name, args) if (D.world == 1 and not args.no_cpu_baseline) else None
        strong = name in STRONG
        out = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": m["desc"], "spp_total": m["spp_total"], "spp_per_gpu": m["spp_total"] / D.world if strong else m["spp"],
                          "parallelism": "replicated scene, sample-index sharding, NCCL film reduce" if D.world > 1 else "single GPU",
                          "l2": "flushed between timed steps (256 MiB write)", "bvh": BVH_NOTE},
               "samples_per_s": m["samples_per_s"], "rays_per_step": m["rays_per_step"],
               "clocks": m["clocks"], "e2e": m.get("e2e"), "gpu_launches": m["launches"], "roofline": m.get("roofline"), "cpu_baseline": base,
               "film_check": check, "bvh_build_ms": m["scene"].stats()["bvh_build_seconds"] * 1e3, "commit": git_commit()}
        out.update(extras)
        print(json.dumps(out))
    if D.dist:
        D.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--tris", type=int, default=1_000_000)
    ap.add_argument("--detail", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
