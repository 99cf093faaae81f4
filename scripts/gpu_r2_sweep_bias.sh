# vote bias at refill threshold 16 on C4 (64 spp) and the C3 batches, final code
TAG=${1:-r2al}
mkdir -p gpurun_out
for B in 14 24 32 40 56 80; do
  FTN_VOTE_BIAS=$B timeout 300 python bench.py --workload c4 --spp 64 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/swb4_${B}_$TAG.json 2> gpurun_out/swb_$TAG.err
  FTN_VOTE_BIAS=$B timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/swb3_${B}_$TAG.json 2>> gpurun_out/swb_$TAG.err
  python - <<PY
import json
d=json.load(open("gpurun_out/swb4_${B}_$TAG.json")); r=d["roofline"]
c=json.load(open("gpurun_out/swb3_${B}_$TAG.json"))["batches"]
print("bias $B: c4 %.1f  extend %.0f shadow %.0f mis %.0f | c3 %s" % (d["value"], r["kernel_mrays_per_s"], r["shadow_mrays_per_s"], r["mis_mrays_per_s"], " ".join("%.0f" % v["mrays_per_s"] for v in c.values())))
PY
done
