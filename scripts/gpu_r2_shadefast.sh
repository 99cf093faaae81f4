TAG=${1:-r2i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_render.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_render_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_render_$TAG.log
tail -8 gpurun_out/pytest_render_$TAG.log
bash scripts/gpu_ab_layout.sh $TAG "default"
