#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q -x --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python tools/sweep.py --size 512 ) > gpurun_out/sweep512.log 2>&1
( timeout 300 python tools/sweep.py --size 1024 --reps 5 ) > gpurun_out/sweep1024.log 2>&1
( timeout 300 python tools/sweep.py --shape 1024,256,512 --axes 0 --reps 5 ) > gpurun_out/sweep_c3_stage2.log 2>&1
( timeout 300 python tools/sweep.py --shape 256,1024,512 --axes 1 --reps 5 ) > gpurun_out/sweep_c3_stage1.log 2>&1
cat gpurun_out/sweep512.log gpurun_out/sweep1024.log gpurun_out/sweep_c3_stage2.log gpurun_out/sweep_c3_stage1.log
