# full -m gpu tier + short benches of the three workloads
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
bash scripts/gpu_ab_layout.sh $TAG "default"
