# round 2, GPU call 1: the full -m gpu tier (with the new 1M / 8M / concurrency tests), the default bench line (C4 strong-scaled,
# C2 + C3 as extra keys), the C3 line, and one ncu --set full capture of the C3 traversal launches (baseline of the round).
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; echo "rc=$?" >> gpurun_out/bench_default_$TAG.err
tail -3 gpurun_out/bench_default_$TAG.err
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; echo "rc=$?" >> gpurun_out/bench_c3_$TAG.err
tail -2 gpurun_out/bench_c3_$TAG.err
C3="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline"
$C3 > gpurun_out/plain_c3_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect_batch -s 19 -c 6 -o gpurun_out/prof_c3_$TAG $C3 > gpurun_out/ncu_c3_$TAG.log 2>&1
tail -3 gpurun_out/ncu_c3_$TAG.log
python - <<PY
import json
for w in ("default","c3"):
    try:
        d=json.load(open("gpurun_out/bench_%s_$TAG.json"%w))
    except Exception as e:
        print(w,"FAILED",e); print(open("gpurun_out/bench_%s_$TAG.err"%w).read()[-2500:]); continue
    print(w,"value %.1f Mrays/s  ms/step %.3f  e2e %s  launches %s"%(d["value"],d["ms_per_step"],d["e2e"] and round(d["e2e"]["value"],1),d["gpu_launches"]))
    b=d.get("batches") or (d.get("c3") or {}).get("batches")
    if b:
        for k,v in b.items(): print("   %-20s %.1f Mrays/s  %.3f ms  nodes/ray %.1f tris/ray %.1f  B/ray %.0f"%(k,v["mrays_per_s"],v["ms_median"],v["nodes_per_ray"],v["tris_per_ray"],v["bytes_per_ray"]))
    if "roofline" in d and d["roofline"]:
        r=d["roofline"]; print("   roofline",{k:r[k] for k in r if k in("kernel","achieved","peak","frac","kernel_share_of_step","all_traversal_share_of_step","shade_share_of_step","kernel_mrays_per_s","nodes_per_ray","tris_per_ray")})
    if "c2" in d: print("   c2 %.1f Mrays/s %.3f ms e2e %.1f"%(d["c2"]["value"],d["c2"]["ms_per_step"],d["c2"]["e2e"]["value"]))
    print("   build",d.get("build") or (d.get("c3") or {}).get("build"))
    print("   cpu",d.get("cpu_baseline"))
PY
