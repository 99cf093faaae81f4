# C4 at 64 spp with 16 / 32 / 64 Mi paths per wavefront pass (FTN_PATHS_PER_PASS); C2 for reference
TAG=${1:-r2q}
mkdir -p gpurun_out
for P in 16777216 33554432 67108864; do
  FTN_PATHS_PER_PASS=$P python bench.py --workload c4 --spp 64 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/pass_c4_${P}_$TAG.json 2> gpurun_out/pass_c4_${P}_$TAG.err
  python - <<PY
import json
d=json.load(open("gpurun_out/pass_c4_${P}_$TAG.json")); r=d["roofline"]
print("paths/pass $P: c4 %.1f Mrays/s %.2f ms/step  extend %.0f shadow %.0f mis %.0f Mr/s shade avg %.3f ms shade share %.3f" % (d["value"], d["ms_per_step"], r["kernel_mrays_per_s"], r["shadow_mrays_per_s"], r["mis_mrays_per_s"], r["shade_avg_launch_ms"], r["shade_share_of_step"]))
PY
done
