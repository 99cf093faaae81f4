# refill threshold x vote bias on the C4 render (final code, 8 blocks per SM), 64 spp
TAG=${1:-r2ak}
mkdir -p gpurun_out
for T in 12 16 20 24; do for B in 14 20 28; do
  FTN_REFILL_THRESHOLD=$T FTN_VOTE_BIAS=$B timeout 300 python bench.py --workload c4 --spp 64 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/sw4_${T}_${B}_$TAG.json 2> gpurun_out/sw4_$TAG.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sw4_${T}_${B}_$TAG.json")); r=d["roofline"]
print("thresh $T bias $B: c4 %.1f  extend %.0f shadow %.0f mis %.0f" % (d["value"], r["kernel_mrays_per_s"], r["shadow_mrays_per_s"], r["mis_mrays_per_s"]))
PY
done; done
