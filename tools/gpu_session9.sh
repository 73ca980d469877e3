#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -q -x --tb=short -k "tma" ) > gpurun_out/pytest_tma.log 2>&1
tail -3 gpurun_out/pytest_tma.log
( timeout 300 python tools/sweep.py --size 1024 --reps 5 --axes 1,0 --engine tma --variants 0,6,100,102 ) > gpurun_out/sweep1024_opt.log 2>&1
( timeout 300 python tools/sweep.py --size 512 --reps 10 --axes 1,0 --engine tma --variants 2,6,8,9,101,103,105,106 ) > gpurun_out/sweep512_opt.log 2>&1
( timeout 300 python tools/sweep.py --size 256 --reps 20 --axes 1,0 --engine tma --variants 0-3,100-102 ) > gpurun_out/sweep256_opt.log 2>&1
( timeout 300 python tools/sweep.py --shape 2048,128,128 --reps 20 --axes 1,0 --engine tma --variants 0-3 ) > gpurun_out/sweep128_opt.log 2>&1
( timeout 300 python tools/sweep.py --shape 256,1024,512 --reps 5 --axes 1 --engine tma --variants 0,6,100,102 ) > gpurun_out/sweep_c3s1.log 2>&1
( timeout 300 python tools/sweep.py --shape 1024,256,512 --reps 5 --axes 0 --engine tma --variants 0,6,100,102 ) > gpurun_out/sweep_c3s2.log 2>&1
cat gpurun_out/sweep1024_opt.log gpurun_out/sweep512_opt.log gpurun_out/sweep256_opt.log gpurun_out/sweep128_opt.log gpurun_out/sweep_c3s1.log gpurun_out/sweep_c3s2.log
