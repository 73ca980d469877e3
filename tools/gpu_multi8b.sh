#!/bin/bash
mkdir -p gpurun_out
( MP_ONLY=c4_r2c_16_p8_slab timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 tests/mp_worker.py ) > gpurun_out/mp8_c4.log 2>&1
grep -v "^$" gpurun_out/mp8_c4.log | grep -B2 -A25 "RANK 0 FAILED" | head -60
tail -3 gpurun_out/mp8_c4.log
