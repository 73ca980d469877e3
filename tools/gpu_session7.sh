#!/bin/bash
# full check: GPU parity suite, bench (both arms), ncu launch list of the bench command, full capture of the slowest kernel
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -x --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench1024.log 2>&1
tail -1 gpurun_out/bench1024.log
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
tail -1 gpurun_out/bench_ref.log
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1024.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_ -s 9 -c 3 -f -o gpurun_out/r1_full_1024 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
