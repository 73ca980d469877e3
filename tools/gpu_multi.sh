#!/bin/bash
# usage: tools/gpu_multi.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
nvidia-smi topo -m >> gpurun_out/gpus_$N.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pfft.py -m gpu -q -x --tb=short -k "nccl or put_kernel" ) > gpurun_out/pytest_multi_$N.log 2>&1
tail -15 gpurun_out/pytest_multi_$N.log
for P2P in 1 0; do
( B2F_P2P=$P2P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$P2P bench.py --gpus $N --steps 10 --warmup 3 --no-e2e ) > gpurun_out/bench1024_g${N}_p2p$P2P.log 2>&1
tail -1 gpurun_out/bench1024_g${N}_p2p$P2P.log
done
