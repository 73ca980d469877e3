#!/bin/bash
# usage: tools/gpu_multi_quick.sh N   (under gpurun --gpus N): multi-rank parity in the three transfer modes + the default bench
N=$1
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short ) > gpurun_out/pytest_multi_$N.log 2>&1
tail -5 gpurun_out/pytest_multi_$N.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e ) > gpurun_out/bench1024_g${N}_fused.log 2>&1
tail -1 gpurun_out/bench1024_g${N}_fused.log | cut -c1-400
nvidia-smi --query-gpu=index,name,memory.used --format=csv
