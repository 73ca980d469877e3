#!/bin/bash
# first GPU session: smoke, parity tests, variant sweep, bench, ncu evidence
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
( time python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 -x --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( timeout 300 python tools/sweep.py --size 512 ) > gpurun_out/sweep512.log 2>&1
( timeout 300 python tools/sweep.py --size 1024 --reps 5 ) > gpurun_out/sweep1024.log 2>&1
( timeout 600 python bench.py --size 512 --steps 10 --warmup 3 ) > gpurun_out/bench512.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench1024.log 2>&1
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --size 512 --steps 2 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/ncu_launches.log 2>&1
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_pow2 -s 6 -c 3 -f -o gpurun_out/prof_r1 python bench.py --size 512 --steps 1 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/smoke.log | tail -3
cat gpurun_out/sweep512.log
cat gpurun_out/bench512.log | tail -2
cat gpurun_out/bench1024.log | tail -2
