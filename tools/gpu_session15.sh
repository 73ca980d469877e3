#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/bench_split.log 2>&1
tail -1 gpurun_out/bench_split.log | cut -c1-900
