# N-GPU call: the multi-GPU tests, then the default bench line under torchrun at N GPUs (film_check inside)
N=${1:-2}; TAG=${2:-r2g}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/smi_$TAG.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi_$TAG.log
tail -6 gpurun_out/pytest_multi_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 2 --warmup 3 --no-extras > gpurun_out/bench_default_n${N}_$TAG.json 2> gpurun_out/bench_default_n${N}_$TAG.err; echo "rc=$?"
tail -3 gpurun_out/bench_default_n${N}_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_default_n${N}_$TAG.json").read().strip().splitlines()[-1])
    print("N=$N value %.1f Mrays/s ms/step %.1f e2e %s film_check %s scaling %s"%(d["value"],d["ms_per_step"],d["e2e"] and round(d["e2e"]["value"],1),d.get("film_check"),d.get("scaling")))
    print(d["config"])
except Exception as e: print("FAILED",e)
PY
