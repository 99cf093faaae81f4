# A/B of run-time switches:  bash scripts/gpu_ab_env.sh "FTN_TRAVERSE_VOTE=0" "FTN_TRAVERSE_VOTE=1" ...
mkdir -p gpurun_out
for E in "$@"; do
  env $E python bench.py --workload c3 --steps 7 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); b=d['batches']
print('[$E] c3', ' | '.join('%s %.0f Mr/s n/r %.1f t/r %.1f'%(k[:14],v['mrays_per_s'],v['nodes_per_ray'],v['tris_per_ray']) for k,v in b.items()))"
  env $E python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']; print('[$E] c2 value %.0f ms %.3f extend %.0f Mr/s share %.2f trav share %.2f'%(d['value'],d['ms_per_step'],r['kernel_mrays_per_s'],r['kernel_share_of_step'],r['all_traversal_share_of_step']))"
done
