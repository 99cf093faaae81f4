# ncu --set full of every k_extend launch of a short C2 run -> dram bytes per launch (roofline.traffic)
mkdir -p gpurun_out
C2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C2 > gpurun_out/plain_c2t.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extend -c 15 -o gpurun_out/prof_c2_extend_${1:-x} $C2 > gpurun_out/ncu_c2_traffic.log 2>&1
tail -2 gpurun_out/ncu_c2_traffic.log
