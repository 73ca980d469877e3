#!/bin/bash
mkdir -p gpurun_out
for c in c2 c4 c5 "384,384,384" ; do
( timeout 600 python tools/bench_cases.py --case $c --reps 5 ) > gpurun_out/case_$c.log 2>&1
cat gpurun_out/case_$c.log | tail -12
done
( timeout 600 python tools/bench_cases.py --case 256,256,256 --padding 1.5 --reps 5 ) > gpurun_out/case_pad.log 2>&1; tail -12 gpurun_out/case_pad.log
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench1024.log 2>&1
tail -1 gpurun_out/bench1024.log
