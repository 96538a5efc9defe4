TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
echo "== p2p graph"; timeout 300 $TR bench.py --gpus 2 --steps 60 --warmup 6 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_ab_p2p_graph.json 2>gpurun_out/err1.log; tail -n 3 gpurun_out/err1.log
echo "== nccl graph"; OPESCI_HALO_P2P=0 timeout 300 $TR bench.py --gpus 2 --steps 60 --warmup 6 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_ab_nccl_graph.json 2>gpurun_out/err2.log; tail -n 3 gpurun_out/err2.log
echo "== p2p nograph"; OPESCI_NO_CUDA_GRAPH=1 timeout 300 $TR bench.py --gpus 2 --steps 60 --warmup 6 --no-e2e --no-cpu --no-ref-arith > gpurun_out/r02e_ab_p2p_nograph.json 2>gpurun_out/err3.log; tail -n 3 gpurun_out/err3.log
for f in gpurun_out/r02e_ab_p2p_graph.json gpurun_out/r02e_ab_nccl_graph.json gpurun_out/r02e_ab_p2p_nograph.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d.get('slab_parity'), d['config'].get('halo_transport'), d['gpu_launches'])"; done
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_loopback.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02e_gpu_multi_tests.log
