#!/bin/bash
# usage: tools/gpu_multi_quick.sh N   (under gpurun --gpus N): multi-rank parity in the transfer modes + the bench with and without pipelining
N=$1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short ) > gpurun_out/pytest_multi_$N.log 2>&1
tail -5 gpurun_out/pytest_multi_$N.log | cut -c1-300
for K in 4 0 2 8; do
( B2F_PIPELINE=$K timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$K bench.py --gpus $N --steps 10 --warmup 3 --no-e2e ) > gpurun_out/bench1024_g${N}_pipe$K.log 2>&1
echo "K=$K $(tail -1 gpurun_out/bench1024_g${N}_pipe$K.log | cut -c1-160)"
done
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | tr '\n' ' '
