#!/bin/bash
N=$1
mkdir -p gpurun_out
for cfg in "8 60" "8 74" "8 89" "8 104" "16 74" "16 89" "6 80"; do
set -- $cfg
( B2F_PIPELINE=$1 B2F_PIPE_SMS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29619 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e ) > gpurun_out/tune_g${N}_$1_$2.log 2>&1
echo "K=$1 SMS=$2 $(tail -1 gpurun_out/tune_g${N}_$1_$2.log | cut -c60-140)"
done
