mkdir -p gpurun_out
C3="python bench.py --workload c3 --steps 1 --warmup 1"
$C3 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_intersect_batch -s ${2:-7} -c ${3:-2} -o gpurun_out/prof_c3_${1:-x} $C3 > gpurun_out/ncu_c3_full.log 2>&1
tail -3 gpurun_out/ncu_c3_full.log
