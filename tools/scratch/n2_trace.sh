TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
OPESCI_STEP_TRACE=1 timeout 300 python bench.py --steps 40 --warmup 6 --no-e2e --no-cpu --no-ref-arith 2>&1 | grep -a "opesci trace\|ms_per_step" | cut -c1-400
OPESCI_STEP_TRACE=1 timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 6 --no-e2e --no-cpu --no-ref-arith 2>&1 | grep -a "opesci trace\|ms_per_step" | cut -c1-400
