TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
echo "== small"; timeout 200 $TR bench.py --gpus 2 --n 256 --steps 20 --warmup 4 --no-e2e --no-cpu --no-ref-arith 2>&1 | grep -a "opesci\|ms_per_step\|rror" | cut -c1-900
echo "== trace p2p"; OPESCI_STEP_TRACE=1 timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 6 --no-e2e --no-cpu --no-ref-arith 2>&1 | grep -a "opesci trace.*40 steps\|rror" | cut -c1-400
echo "== p2p"; timeout 300 $TR bench.py --gpus 2 --steps 60 --warmup 6 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_ab_p2p.json 2>gpurun_out/err1.log; tail -n 3 gpurun_out/err1.log
echo "== nccl"; OPESCI_HALO_P2P=0 timeout 300 $TR bench.py --gpus 2 --steps 60 --warmup 6 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_ab_nccl.json 2>gpurun_out/err2.log; tail -n 3 gpurun_out/err2.log
for f in gpurun_out/r02e_ab_p2p.json gpurun_out/r02e_ab_nccl.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('slab_parity'), d['config'].get('halo_transport'))"; done
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02e_gpu_multi_tests.log
