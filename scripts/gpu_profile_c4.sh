# launch shares of the C4 workload (gear ring, TR copper, image env map, thin lens) at 16 spp
mkdir -p gpurun_out
C4="python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C4 > gpurun_out/plain_c4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c4.csv $C4 > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/plain_c4.log | cut -c1-200
