mkdir -p gpurun_out
for T in 8 12 16 20 24 28 32; do
  FTN_REFILL_THRESHOLD=$T python bench.py --workload c3 --steps 5 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); b=d['batches']
print('thresh $T', ' '.join('%s %.0f'%(k[:12],v['mrays_per_s']) for k,v in b.items()))"
done
for T in 12 20 28; do
  FTN_REFILL_THRESHOLD=$T python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); print('c2 thresh $T value %.0f ms %.3f'%(d['value'],d['ms_per_step']))"
done
