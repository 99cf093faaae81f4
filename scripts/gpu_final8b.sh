# final-code check of the 8-GPU lines: C2 (default bench) and C4 at its named 1024 spp
mkdir -p gpurun_out
for W in c2 c4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --workload $W --gpus 8 --steps 5 --warmup 3 > gpurun_out/final8_$W.out 2> gpurun_out/final8_$W.err
echo "$W N=8 rc=$?"; grep '^{' gpurun_out/final8_$W.out | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.0f ms %.3f e2e %.0f clocks %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks']))"
done
