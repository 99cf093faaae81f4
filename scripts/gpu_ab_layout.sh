# A/B of the node layouts on the three render / batch workloads (short runs)
TAG=${1:-r2e}
LAYOUTS=${2:-"bvh8"}
mkdir -p gpurun_out
for L in $LAYOUTS; do
  LAYOUT_ENV=$([ "$L" = default ] && echo "" || echo "FTN_BVH_LAYOUT=$L"); env $LAYOUT_ENV timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_c3_${L}_$TAG.json 2> gpurun_out/ab_c3_${L}_$TAG.err
  LAYOUT_ENV=$([ "$L" = default ] && echo "" || echo "FTN_BVH_LAYOUT=$L"); env $LAYOUT_ENV timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_c2_${L}_$TAG.json 2> gpurun_out/ab_c2_${L}_$TAG.err
  LAYOUT_ENV=$([ "$L" = default ] && echo "" || echo "FTN_BVH_LAYOUT=$L"); env $LAYOUT_ENV timeout 600 python bench.py --workload c4 --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/ab_c4_${L}_$TAG.json 2> gpurun_out/ab_c4_${L}_$TAG.err
done
python - <<PY
import json
for L in "$LAYOUTS".split():
    for w in ("c3","c2","c4"):
        f="gpurun_out/ab_%s_%s_$TAG.json"%(w,L)
        try: d=json.load(open(f))
        except Exception as e:
            print(L,w,"FAILED",e); print(open(f.replace(".json",".err")).read()[-1500:]); continue
        print(L,w,"value %.1f Mrays/s  ms/step %.3f build %.2f ms"%(d["value"],d["ms_per_step"],d.get("bvh_build_ms") or -1))
        for k,v in (d.get("batches") or {}).items(): print("   %-20s %.1f Mrays/s  %.3f ms  nodes/ray %.1f tris/ray %.1f"%(k,v["mrays_per_s"],v["ms_median"],v["nodes_per_ray"],v["tris_per_ray"]))
        r=d.get("roofline") or {}
        if w!="c3": print("   ", {k:(round(r[k],3) if isinstance(r.get(k),float) else r.get(k)) for k in ("kernel_mrays_per_s","shadow_mrays_per_s","mis_mrays_per_s","all_traversal_share_of_step","shade_share_of_step","shade_avg_launch_ms")})
PY
