#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -q -x --tb=short -k "tma" ) > gpurun_out/pytest_tma.log 2>&1
tail -3 gpurun_out/pytest_tma.log
( timeout 300 python tools/sweep.py --size 768 --reps 10 --axes 1,0 --engine tma --variants 100,101 ) > gpurun_out/sweep768.log 2>&1
( timeout 300 python tools/sweep.py --size 384 --reps 20 --axes 1,0 --engine tma --variants 100,101 ) > gpurun_out/sweep384.log 2>&1
( timeout 300 python tools/sweep.py --shape 1024,192,192 --reps 20 --axes 1 --engine tma --variants 100 ) > gpurun_out/sweep192.log 2>&1
( timeout 300 python tools/sweep.py --shape 1024,192,192 --reps 20 --axes 1 ) > gpurun_out/sweep192r.log 2>&1
cat gpurun_out/sweep768.log gpurun_out/sweep384.log gpurun_out/sweep192.log gpurun_out/sweep192r.log
