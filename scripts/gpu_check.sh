# usage: bash scripts/gpu_check.sh [tag]   -- GPU tests + c2/c3 benches, results under gpurun_out/
TAG=${1:-check}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "rc=$?" >> gpurun_out/bench_c2_$TAG.err
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "rc=$?" >> gpurun_out/bench_c3_$TAG.err
tail -4 gpurun_out/pytest_$TAG.log
python - <<PY
import json
for w in ("c2","c3"):
    try:
        d=json.load(open("gpurun_out/bench_%s_$TAG.json"%w))
    except Exception as e:
        print(w,"FAILED",e); print(open("gpurun_out/bench_%s_$TAG.err"%w).read()[-1500:]); continue
    print(w,"value %.1f Mrays/s  ms/step %.3f  e2e %s  launches %s"%(d["value"],d["ms_per_step"],d["e2e"] and round(d["e2e"]["value"],1),d["gpu_launches"]))
    if "batches" in d:
        for k,b in d["batches"].items(): print("   %-20s %.1f Mrays/s  %.3f ms  nodes/ray %.1f tris/ray %.1f  B/ray %.0f  GB/s %.0f hit %.2f"%(k,b["mrays_per_s"],b["ms_median"],b["nodes_per_ray"],b["tris_per_ray"],b["bytes_per_ray"],b["achieved_gbs"],b["hit_fraction"]))
    else:
        r=d["roofline"]; print("   k_extend %.1f Mrays/s avg %.3f ms share %.2f all-trav share %.2f frac %.2f; cpu %s"%(r["kernel_mrays_per_s"],r["avg_launch_ms"],r["kernel_share_of_step"],r["all_traversal_share_of_step"],r["frac"],d["cpu_baseline"] and round(d["cpu_baseline"]["value"],1)))
PY
