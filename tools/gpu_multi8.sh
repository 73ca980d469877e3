#!/bin/bash
# under gpurun --gpus 8: fused-mode parity for world 8 and the default bench (kept short: 8x GPU-minutes)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short -k "8-1" ) > gpurun_out/pytest_multi_8.log 2>&1
tail -5 gpurun_out/pytest_multi_8.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e ) > gpurun_out/bench1024_g8_fused.log 2>&1
tail -1 gpurun_out/bench1024_g8_fused.log
( B2F_FUSED=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e ) > gpurun_out/bench1024_g8_put.log 2>&1
tail -1 gpurun_out/bench1024_g8_put.log | cut -c1-200
( B2F_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29615 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e ) > gpurun_out/bench1024_g8_nccl.log 2>&1
tail -1 gpurun_out/bench1024_g8_nccl.log | cut -c1-200
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader | tr '\n' ' '
