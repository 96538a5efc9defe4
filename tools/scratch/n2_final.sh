N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
if [ "$N" = 2 ]; then
timeout 500 $TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02e_bench_default_gpus$N.json 2>gpurun_out/err1.log; tail -n 2 gpurun_out/err1.log
else
timeout 300 $TR bench.py --gpus $N --steps 60 --warmup 6 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_bench_default_gpus$N.json 2>gpurun_out/err1.log; tail -n 2 gpurun_out/err1.log
fi
timeout 300 $TR bench.py --gpus $N --config hetero --steps 30 --warmup 5 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_bench_hetero_gpus$N.json 2>gpurun_out/err2.log; tail -n 2 gpurun_out/err2.log
for f in gpurun_out/r02e_bench_default_gpus$N.json gpurun_out/r02e_bench_hetero_gpus$N.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('slab_parity'), d['config'].get('halo_transport'), d['gpu_launches'], (d.get('e2e') or {}).get('value'))"; done
