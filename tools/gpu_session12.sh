#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_pfft.py -m gpu -q --tb=short -k "played" ) > gpurun_out/pytest_played.log 2>&1
tail -40 gpurun_out/pytest_played.log | cut -c1-250
