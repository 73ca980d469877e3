#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python tools/sweep.py --size 1024 --reps 5 --axes 1 --engine tma --variants 100,102 ) > gpurun_out/diag_a.log 2>&1
( timeout 200 python tools/sweep.py --size 1024 --reps 5 --axes 1 --engine tma --variants 100,102 --prealloc-gib 48 ) > gpurun_out/diag_b.log 2>&1
( timeout 200 python tools/sweep.py --size 1024 --reps 40 --axes 1 --engine tma --variants 100,102 ) > gpurun_out/diag_c.log 2>&1
grep "axis" gpurun_out/diag_a.log gpurun_out/diag_b.log gpurun_out/diag_c.log
