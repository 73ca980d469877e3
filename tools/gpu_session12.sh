#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_pfft.py -m gpu -q --tb=short -x -k "partial or fused or played" ) > gpurun_out/pytest_partial.log 2>&1
tail -30 gpurun_out/pytest_partial.log | cut -c1-250
