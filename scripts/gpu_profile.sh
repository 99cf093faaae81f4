# ncu evidence for the round: launch lists (shares) + one full capture of the top kernel per workload
mkdir -p gpurun_out
C2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
C3="python bench.py --workload c3 --steps 1 --warmup 1"
$C2 > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c2.csv $C2 > gpurun_out/ncu_c2.log 2>&1
$C2 > gpurun_out/plain_c2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 6 -c 3 -o gpurun_out/prof_c2_extend $C2 > gpurun_out/ncu_c2_full.log 2>&1
$C3 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_intersect_batch -s 4 -c 3 -o gpurun_out/prof_c3_intersect $C3 > gpurun_out/ncu_c3_full.log 2>&1
ls -la gpurun_out | tail -20
tail -3 gpurun_out/ncu_c2.log gpurun_out/ncu_c2_full.log gpurun_out/ncu_c3_full.log
