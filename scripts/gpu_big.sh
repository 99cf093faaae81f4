mkdir -p gpurun_out
summ() { python -c "
import json,sys
d=json.load(open('$1'))
r=d['roofline']
print('$1', 'value %.0f Mr/s ms %.2f samples/s %.3g e2e %s build %.1f ms; extend %.0f Mr/s n/r %.1f t/r %.1f B/ray %.0f achieved %.0f GB/s frac %.2f share %.2f' % (d['value'], d['ms_per_step'], d['samples_per_s'], d['e2e'] and round(d['e2e']['value']), d['bvh_build_ms'], r['kernel_mrays_per_s'], r['nodes_per_ray'], r['tris_per_ray'], r['bytes_per_ray'], r['achieved'], r['frac'], r['all_traversal_share_of_step']))
print('   ', d['config']['workload'], '| cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'],1))
"; }
timeout 900 python bench.py --workload c4 --steps 3 --warmup 2 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; summ gpurun_out/bench_c4.json || tail -5 gpurun_out/bench_c4.err
timeout 900 python bench.py --workload c5 --tris 8000000 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c5_8m.json 2> gpurun_out/bench_c5_8m.err; summ gpurun_out/bench_c5_8m.json || tail -5 gpurun_out/bench_c5_8m.err
timeout 1500 python bench.py --workload c5 --tris 50000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_50m.json 2> gpurun_out/bench_c5_50m.err; summ gpurun_out/bench_c5_50m.json || tail -5 gpurun_out/bench_c5_50m.err
nvidia-smi --query-gpu=memory.used --format=csv
