# ncu evidence for C2: launch list (shares) + full capture of one kernel class (arg 2, default k_shade)
TAG=${1:-x}
KERNEL=${2:-k_shade}
mkdir -p gpurun_out
C2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C2 > gpurun_out/plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c2_$TAG.csv $C2 > gpurun_out/ncu_c2.log 2>&1
$C2 > gpurun_out/plain_c2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 2 -c 2 -o gpurun_out/prof_c2_${KERNEL}_$TAG $C2 > gpurun_out/ncu_c2_full.log 2>&1
tail -2 gpurun_out/ncu_c2_full.log
