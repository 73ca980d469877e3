#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
