#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_serial.py -m gpu -q -x --tb=short -k "tma or stockham" ) > gpurun_out/pytest_tma.log 2>&1
tail -5 gpurun_out/pytest_tma.log
V=0-7,100-102
( timeout 300 python tools/sweep.py --size 512 --axes 1,0 --engine tma --variants $V ) > gpurun_out/sweep512_tma.log 2>&1
( timeout 300 python tools/sweep.py --size 1024 --reps 5 --axes 1,0 --engine tma --variants $V ) > gpurun_out/sweep1024_tma.log 2>&1
( timeout 300 python tools/sweep.py --size 512 --axes 2,1,0 ) > gpurun_out/sweep512_reg.log 2>&1
( timeout 300 python tools/sweep.py --size 1024 --reps 5 --axes 2,1,0 ) > gpurun_out/sweep1024_reg.log 2>&1
( timeout 300 python tools/sweep.py --shape 1024,256,512 --axes 0 --reps 5 --engine tma --variants $V ) > gpurun_out/sweep_c3_stage2_tma.log 2>&1
( timeout 300 python tools/sweep.py --shape 256,1024,512 --axes 1 --reps 5 --engine tma --variants $V ) > gpurun_out/sweep_c3_stage1_tma.log 2>&1
cat gpurun_out/sweep512_tma.log gpurun_out/sweep1024_tma.log gpurun_out/sweep512_reg.log gpurun_out/sweep1024_reg.log gpurun_out/sweep_c3_stage2_tma.log gpurun_out/sweep_c3_stage1_tma.log
timeout 100 tools/probe/stride_probe.bin 512 16 1 > gpurun_out/probe512_pad16.txt 2>&1
timeout 100 tools/probe/stride_probe.bin 512 0 1 > gpurun_out/probe512_pad0.txt 2>&1
timeout 100 tools/probe/stride_probe.bin 512 4112 1 > gpurun_out/probe512_pad4112.txt 2>&1
cat gpurun_out/probe512_pad0.txt gpurun_out/probe512_pad16.txt gpurun_out/probe512_pad4112.txt
