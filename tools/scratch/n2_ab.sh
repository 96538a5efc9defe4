set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02e_gpu_multi_tests.log
timeout 300 $TR bench.py --gpus 2 --steps 60 --warmup 6 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_ab_new.json 2>gpurun_out/err1.log
OPESCI_SLAB_MIDOVERLAP=1 timeout 300 $TR bench.py --gpus 2 --steps 60 --warmup 6 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_ab_old.json 2>gpurun_out/err2.log
timeout 300 $TR bench.py --gpus 2 --steps 60 --warmup 6 --config hetero --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_ab_hetero_new.json 2>gpurun_out/err3.log
cat gpurun_out/r02e_gpu_multi_tests.log
for f in gpurun_out/r02e_ab_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('slab_parity'))"; done
tail -3 gpurun_out/err*.log
