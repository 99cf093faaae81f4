# round-2 checkpoint: -m gpu tier, the default bench line (+ reference arm), launch list and ncu --set full of the C4 kernels
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -6 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err; echo "ref rc=$?"
C4="python bench.py --workload c4 --spp 16 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
$C4 > gpurun_out/plain_c4_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c4_$TAG.csv $C4 > gpurun_out/ncu_launches_c4_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_extend|k_mis|k_shadow" -s 8 -c 10 -o gpurun_out/prof_c4_$TAG $C4 > gpurun_out/ncu_c4_$TAG.log 2>&1
tail -2 gpurun_out/ncu_c4_$TAG.log
cut -c1-400 gpurun_out/bench_default_$TAG.json
