# A/B of library variants on C3 and C2:  bash scripts/gpu_ab.sh libA.so libB.so ...
# (paths relative to fountain_b200/csrc; "default" = libfountain_gpu.so)
mkdir -p gpurun_out
for L in "$@"; do
  if [ "$L" = "default" ]; then unset FTN_GPU_LIB; else export FTN_GPU_LIB=$PWD/fountain_b200/csrc/$L; fi
  python bench.py --workload c3 --steps 7 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); b=d['batches']
print('$L c3', ' | '.join('%s %.0f Mr/s n/r %.1f t/r %.1f'%(k[:14],v['mrays_per_s'],v['nodes_per_ray'],v['tris_per_ray']) for k,v in b.items()), 'build %.1f ms nodes %d'%(d['bvh_build_ms'], d['bvh_nodes']))"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']; print('$L c2 value %.0f ms %.3f extend %.0f Mr/s share %.2f trav share %.2f'%(d['value'],d['ms_per_step'],r['kernel_mrays_per_s'],r['kernel_share_of_step'],r['all_traversal_share_of_step']))"
done
