# ncu --set full of the k_extend launches of one C5 (8M triangles, 4K, 4 spp) step: the HBM-bound case
mkdir -p gpurun_out
C5="python bench.py --workload c5 --tris 8000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C5 > gpurun_out/plain_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extend -c 6 -o gpurun_out/prof_c5_extend $C5 > gpurun_out/ncu_c5_full.log 2>&1
tail -2 gpurun_out/ncu_c5_full.log; tail -1 gpurun_out/plain_c5.log | cut -c1-300
