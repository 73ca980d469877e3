#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench1024_g2_full.log 2>&1
tail -1 gpurun_out/bench1024_g2_full.log | cut -c1-3000
