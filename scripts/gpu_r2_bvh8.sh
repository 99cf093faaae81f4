# round 2: BVH8q vs BVH2x64 -- GPU parity on both layouts, then the C3 batches and the C2 / C4 renders under each layout
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_parity_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity_$TAG.log
tail -4 gpurun_out/pytest_parity_$TAG.log
for L in bvh2 bvh8; do
  FTN_BVH_LAYOUT=$L timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_${L}_$TAG.json 2> gpurun_out/bench_c3_${L}_$TAG.err
  FTN_BVH_LAYOUT=$L timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_${L}_$TAG.json 2> gpurun_out/bench_c2_${L}_$TAG.err
  FTN_BVH_LAYOUT=$L timeout 600 python bench.py --workload c4 --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c4_${L}_$TAG.json 2> gpurun_out/bench_c4_${L}_$TAG.err
done
python - <<PY
import json
for L in ("bvh2","bvh8"):
    for w in ("c3","c2","c4"):
        f="gpurun_out/bench_%s_%s_$TAG.json"%(w,L)
        try: d=json.load(open(f))
        except Exception as e:
            print(L,w,"FAILED",e); print(open(f.replace(".json",".err")).read()[-1500:]); continue
        print(L,w,"value %.1f Mrays/s  ms/step %.3f build %.2f ms nodes %s"%(d["value"],d["ms_per_step"],d.get("bvh_build_ms") or -1, d.get("bvh_nodes")))
        for k,v in (d.get("batches") or {}).items(): print("   %-20s %.1f Mrays/s  %.3f ms  nodes/ray %.1f tris/ray %.1f"%(k,v["mrays_per_s"],v["ms_median"],v["nodes_per_ray"],v["tris_per_ray"]))
        r=d.get("roofline") or {}
        print("   ", {k:r.get(k) for k in ("kernel","kernel_mrays_per_s","shadow_mrays_per_s","mis_mrays_per_s","all_traversal_share_of_step","shade_share_of_step","nodes_per_ray","tris_per_ray")})
PY
