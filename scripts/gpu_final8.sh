mkdir -p gpurun_out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final_n$N.out 2> gpurun_out/final_n$N.err
echo "N=$N rc=$?"; grep '^{' gpurun_out/final_n$N.out | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   value %.0f ms %.3f e2e %.0f clocks %s cpu %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['clocks'],d['cpu_baseline']))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/final_ref_n8.out 2> gpurun_out/final_ref_n8.err; echo "ref N=8 rc=$?"; grep -c '^{' gpurun_out/final_ref_n8.out
