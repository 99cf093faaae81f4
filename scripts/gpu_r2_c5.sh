# C5 (large synthetic meshes, 4K, 4 spp) on the final code: 8 M and 50 M triangles, N = 1
TAG=${1:-r2z}
mkdir -p gpurun_out
for T in 8000000 50000000; do
  timeout 1500 python bench.py --workload c5 --tris $T --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_c5_${T}_$TAG.json 2> gpurun_out/bench_c5_${T}_$TAG.err; echo "c5 $T rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_c5_${T}_$TAG.json")); r=d["roofline"]
    print("c5 $T: %.0f Mrays/s %.2f ms/step build %.1f ms extend %.0f Mr/s nodes/ray %.1f tris/ray %.1f node bytes %d working set %.0f MB" % (d["value"], d["ms_per_step"], d["bvh_build_ms"], r["kernel_mrays_per_s"], r["nodes_per_ray"], r["tris_per_ray"], r["node_bytes"], r["working_set_mb"]))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_c5_${T}_$TAG.err").read()[-1200:])
PY
done
nvidia-smi --query-gpu=memory.used --format=csv,noheader
