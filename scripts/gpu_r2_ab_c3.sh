# A/B of library builds on the C3 ray batches:  bash scripts/gpu_r2_ab_c3.sh <tag> <lib> [<lib> ...]
TAG=$1; shift
mkdir -p gpurun_out
for L in "$@"; do
  if [ "$L" = default ]; then unset FTN_GPU_LIB; else export FTN_GPU_LIB=$PWD/fountain_b200/csrc/libfountain_gpu_$L.so; fi
  timeout 600 python bench.py --workload c3 --steps 7 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c3_${L}_$TAG.json 2> gpurun_out/c3_${L}_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c3_${L}_$TAG.json")); b=d["batches"]
    print("$L c3", " | ".join("%s %.0f Mr/s" % (k[:18], v["mrays_per_s"]) for k, v in b.items()), "build %.2f ms" % d["bvh_build_ms"])
except Exception as e:
    print("$L c3 FAILED", e); print(open("gpurun_out/c3_${L}_$TAG.err").read()[-1500:])
PY
done
