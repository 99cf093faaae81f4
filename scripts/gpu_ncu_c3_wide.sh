TAG=${1:-r2c}
C3="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C3 > gpurun_out/plain_c3_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_intersect_batch -c 12 -o gpurun_out/prof_c3_$TAG $C3 > gpurun_out/ncu_c3_$TAG.log 2>&1
tail -3 gpurun_out/ncu_c3_$TAG.log
