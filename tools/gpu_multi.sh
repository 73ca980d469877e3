#!/bin/bash
# usage: tools/gpu_multi.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
nvidia-smi topo -m >> gpurun_out/gpus_$N.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pfft.py -m gpu -q -x --tb=short -k "nccl or put_kernel or fused" ) > gpurun_out/pytest_multi_$N.log 2>&1
tail -15 gpurun_out/pytest_multi_$N.log
for MODE in fused put nccl; do
P2P=1; FUSED=1
[ $MODE = put ] && FUSED=0
[ $MODE = nccl ] && P2P=0
( B2F_P2P=$P2P B2F_FUSED=$FUSED timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e ) > gpurun_out/bench1024_g${N}_$MODE.log 2>&1
tail -1 gpurun_out/bench1024_g${N}_$MODE.log
done
