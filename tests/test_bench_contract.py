"""The bench line committed under profiles/ (written by `python bench.py` on the B200 box) carries every key
of the driver's contract; and `bench.py --impl reference` (CPU, runs here) prints the reference-arm line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def test_committed_bench_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r02_bench_default_n1.json")
    if not os.path.exists(path):
        import pytest
        pytest.skip("no round-2 bench line committed yet")
    d = json.load(open(path))
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "cpu_baseline"} <= set(d)
    assert d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C4") and "1024 spp" in d["config"]["workload"] and "c2" in d and "c3" in d
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["e2e"]["value"] != d["value"] and d["gpu_launches"] > 0
    r = d["roofline"]
    # the stated roof is the limiter ncu measured for this workload's kernel (profiles/traffic.json): instruction issue for the
    # BVH8q traversal; the L1 / L2 / HBM figures of the same algorithmic bytes stand beside it, traffic = ncu's DRAM bytes
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "l1", "l2", "hbm", "ncu"} <= set(r)
    assert r["bound"].startswith("sm-issue") and r["unit"] == "Gwarp-inst/s" and r["traffic"] is not None and r["ncu"]["commit"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["frac"] <= 1.0
    assert r["hbm"]["peak"] > 0 and r["l1"]["frac"] <= 1.0
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port" and c["cores"] >= 1
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])


def test_reference_arm_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--spp", "2"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads(p.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
