TAG=${1:-r2h}
C4="python bench.py --workload c4 --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
$C4 > gpurun_out/plain_c4_$TAG.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 4 -o gpurun_out/prof_c4_shade_$TAG $C4 > gpurun_out/ncu_c4_shade_$TAG.log 2>&1
tail -3 gpurun_out/ncu_c4_shade_$TAG.log
