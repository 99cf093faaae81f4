# 1/2/4-GPU points of the scaling sweeps (C2 and C5 at 8M triangles)
mkdir -p gpurun_out
run() {
  N=$1; TAG=$2; shift 2
  if [ "$N" = "1" ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/scale_$TAG.out 2> gpurun_out/scale_$TAG.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@" > gpurun_out/scale_$TAG.out 2> gpurun_out/scale_$TAG.err; fi
  echo "== $TAG rc=$?"; grep '^{' gpurun_out/scale_$TAG.out | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   value %.0f Mrays/s  ms/step %.2f  e2e %s  n_gpus %d  %s' % (d['value'], d['ms_per_step'], d['e2e'] and round(d['e2e']['value']), d['n_gpus'], d['config']['workload'][:70]))"
}
for N in 1 2 4; do run $N c2_n$N --steps 5 --warmup 3 --no-cpu-baseline; done
for N in 1 2 4; do run $N c5_8m_n$N --workload c5 --tris 8000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e; done
run 1 c4_n1 --workload c4 --spp 128 --steps 3 --warmup 3 --no-cpu-baseline
