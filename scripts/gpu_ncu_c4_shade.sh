mkdir -p gpurun_out
C4="python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C4 > gpurun_out/plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 3 -o gpurun_out/prof_c4_shade $C4 > gpurun_out/ncu_c4_full.log 2>&1
tail -2 gpurun_out/ncu_c4_full.log
