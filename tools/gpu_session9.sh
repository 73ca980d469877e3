#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -q -x --tb=short -k "tma" ) > gpurun_out/pytest_tma.log 2>&1
tail -3 gpurun_out/pytest_tma.log
( timeout 300 python tools/sweep.py --size 1024 --reps 5 --axes 1,0 --engine tma --variants 100,102,106 ) > gpurun_out/sweep1024_pf.log 2>&1
( timeout 300 python tools/sweep.py --size 512 --reps 10 --axes 1,0 --engine tma --variants 2,105,107 ) > gpurun_out/sweep512_pf.log 2>&1
( timeout 300 python tools/sweep.py --size 256 --reps 20 --axes 1,0 --engine tma --variants 102,103 ) > gpurun_out/sweep256_pf.log 2>&1
( timeout 300 python tools/sweep.py --shape 1024,256,512 --reps 5 --axes 0 --engine tma --variants 100,102,106 ) > gpurun_out/sweep_c3s2_pf.log 2>&1
cat gpurun_out/sweep1024_pf.log gpurun_out/sweep512_pf.log gpurun_out/sweep256_pf.log gpurun_out/sweep_c3s2_pf.log
