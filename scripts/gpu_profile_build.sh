# launch list of a 1M-triangle scene build (PLOC): where does the build time go?
mkdir -p gpurun_out
CMD="python scripts/exp_build_time.py c3"
FTN_BVH_BUILDER=ploc $CMD > gpurun_out/plain_build.log 2>&1 && \
FTN_BVH_BUILDER=ploc ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_build.csv $CMD > gpurun_out/ncu_build.log 2>&1
tail -3 gpurun_out/plain_build.log
