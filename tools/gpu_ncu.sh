#!/bin/bash
# usage: tools/gpu_ncu.sh <name> <sweep args...>   -> gpurun_out/<name>.ncu-rep (kernels of the sweep, after warm-up)
name=$1; shift
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_ -s 3 -c 2 -f -o gpurun_out/$name python tools/sweep.py --reps 1 "$@" > gpurun_out/$name.log 2>&1
tail -3 gpurun_out/$name.log
