#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/sweep.py --size 512 --axes 0 --engine tma --variants 5-7 ) 2>&1 | tail -4
for v in 0 1 2 4; do tools/gpu_ncu.sh t1024_ax1_v$v --size 1024 --axes 1 --engine tma --variants $v >/dev/null; done
tools/gpu_ncu.sh t512_ax1_v2 --size 512 --axes 1 --engine tma --variants 2 > /dev/null
tools/gpu_ncu.sh t512_ax0_v5 --size 512 --axes 0 --engine tma --variants 5 > /dev/null
