TAG=${1:-r2d}
for T in 16 20 24 28; do for B in 10 14 18 24; do
  FTN_REFILL_THRESHOLD=$T FTN_VOTE_BIAS=$B timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sw_${T}_${B}_$TAG.json 2> gpurun_out/sw_$TAG.err
  python - <<PY
import json
d=json.load(open("gpurun_out/sw_${T}_${B}_$TAG.json"))
b=d["batches"]
print("thresh $T bias $B:", "  ".join("%s %.0f"%(k[:9],v["mrays_per_s"]) for k,v in b.items()))
PY
done; done
