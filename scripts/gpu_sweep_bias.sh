# sweep of the node/leaf vote bias (and refill threshold) on C3 ray batches and the C2 render
mkdir -p gpurun_out
for b in 8 12 14 16 20 28; do
  echo "== FTN_VOTE_BIAS=$b"
  FTN_VOTE_BIAS=$b python bench.py --workload c3 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  c3', {k: round(v['mrays_per_s'],1) for k,v in d['batches'].items()})"
  FTN_VOTE_BIAS=$b python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  c2', round(d['value'],1), 'ms', round(d['ms_per_step'],3))"
done
for t in 12 20; do
  echo "== FTN_REFILL_THRESHOLD=$t (bias default)"
  FTN_REFILL_THRESHOLD=$t python bench.py --workload c3 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  c3', {k: round(v['mrays_per_s'],1) for k,v in d['batches'].items()})"
done
