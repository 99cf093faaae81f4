# Final-code ncu records for bench.py's roofline (profiles/traffic.json via scripts/ncu_traffic.py): per-launch instruction,
# pipe and DRAM counters of the dominant traversal kernel of each workload; each ncu command only after the same command
# exited 0 without ncu.  usage: bash scripts/gpu_r2_final.sh <tag>
TAG=${1:-r2u}
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,l1tex__data_pipe_lsu_wavefronts.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
C4="python bench.py --workload c4 --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
C2="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
C3="python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C4 > gpurun_out/final_plain_c4_$TAG.json 2> gpurun_out/final_plain_c4_$TAG.err && \
timeout 900 ncu --metrics $M --clock-control none -k regex:k_extend --csv --log-file gpurun_out/final_ncu_c4_$TAG.csv $C4 > gpurun_out/final_ncu_c4_$TAG.log 2>&1; echo "c4 rc=$?"
$C2 > gpurun_out/final_plain_c2_$TAG.json 2> gpurun_out/final_plain_c2_$TAG.err && \
timeout 900 ncu --metrics $M --clock-control none -k regex:k_extend --csv --log-file gpurun_out/final_ncu_c2_$TAG.csv $C2 > gpurun_out/final_ncu_c2_$TAG.log 2>&1; echo "c2 rc=$?"
$C3 > gpurun_out/final_plain_c3_$TAG.json 2> gpurun_out/final_plain_c3_$TAG.err && \
timeout 900 ncu --metrics $M --clock-control none -k regex:k_intersect --csv --log-file gpurun_out/final_ncu_c3_$TAG.csv $C3 > gpurun_out/final_ncu_c3_$TAG.log 2>&1; echo "c3 rc=$?"
# launch list (durations only) of the default workload at 64 spp: the kernels' shares of the step
C4L="python bench.py --workload c4 --spp 64 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches_c4_$TAG.csv $C4L > gpurun_out/final_launches_c4_$TAG.log 2>&1; echo "launches rc=$?"
ls -la gpurun_out/final_*_$TAG.*
