# A/B of library builds on the render workloads:  bash scripts/gpu_r2_ab_lib.sh <tag> <lib> [<lib> ...]
# ("default" = libfountain_gpu.so; other names = fountain_b200/csrc/libfountain_gpu_<name>.so via FTN_GPU_LIB)
TAG=$1; shift
mkdir -p gpurun_out
for L in "$@"; do
  if [ "$L" = default ]; then unset FTN_GPU_LIB; else export FTN_GPU_LIB=$PWD/fountain_b200/csrc/libfountain_gpu_$L.so; fi
  for rep in 1 2; do
  timeout 600 python bench.py --workload c4 --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/lib_c4_${L}_$TAG.json 2> gpurun_out/lib_c4_${L}_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/lib_c4_${L}_$TAG.json")); r=d["roofline"]
    print("$L c4 %.1f Mrays/s %.2f ms/step  extend %.0f shadow %.0f mis %.0f Mr/s shade avg %.3f ms shade share %.3f" % (d["value"], d["ms_per_step"], r["kernel_mrays_per_s"], r["shadow_mrays_per_s"], r["mis_mrays_per_s"], r["shade_avg_launch_ms"], r["shade_share_of_step"]))
except Exception as e:
    print("$L c4 FAILED", e); print(open("gpurun_out/lib_c4_${L}_$TAG.err").read()[-1500:])
PY
  done
  timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/lib_c2_${L}_$TAG.json 2> gpurun_out/lib_c2_${L}_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/lib_c2_${L}_$TAG.json")); r=d["roofline"]
    print("$L c2 %.1f Mrays/s %.3f ms/step  shade avg %.3f ms shade share %.3f" % (d["value"], d["ms_per_step"], r["shade_avg_launch_ms"], r["shade_share_of_step"]))
except Exception as e:
    print("$L c2 FAILED", e); print(open("gpurun_out/lib_c2_${L}_$TAG.err").read()[-1500:])
PY
done
