#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for c in "384,384,384" "768,768,768" c4; do
( timeout 600 python tools/bench_cases.py --case $c --reps 5 ) > gpurun_out/case_$c.log 2>&1
cat gpurun_out/case_$c.log | tail -9
done
( timeout 600 python tools/bench_cases.py --case 256,256,256 --padding 1.5 --reps 5 ) > gpurun_out/case_pad.log 2>&1; tail -9 gpurun_out/case_pad.log
