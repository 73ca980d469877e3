#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/bench1024.log 2>&1
tail -1 gpurun_out/bench1024.log | cut -c1-300
nvidia-smi --query-gpu=index,name,memory.used --format=csv
