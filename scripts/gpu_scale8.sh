# 8-GPU measurements of the named configs (one box, torchrun, NCCL film reduce)
mkdir -p gpurun_out
run() {  # run <n> <tag> <bench args...>
  N=$1; TAG=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
      > gpurun_out/scale_$TAG.out 2> gpurun_out/scale_$TAG.err
  echo "== $TAG rc=$?"; grep '^{' gpurun_out/scale_$TAG.out | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   value %.0f Mrays/s  ms/step %.2f  e2e %s  n_gpus %d  %s' % (d['value'], d['ms_per_step'], d['e2e'] and round(d['e2e']['value']), d['n_gpus'], d['config']['workload'][:70]))"
}
run 8 c2_n8 --steps 5 --warmup 3 --no-cpu-baseline
run 8 c4_n8 --workload c4 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline
run 8 c5_8m_n8 --workload c5 --tris 8000000 --steps 3 --warmup 3 --no-cpu-baseline
run 8 c5_50m_n8 --workload c5 --tris 50000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
