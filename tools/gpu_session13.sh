#!/bin/bash
mkdir -p gpurun_out
for v in default 100 102; do
if [ $v = default ]; then unset B2F_VARIANT_TMA; else export B2F_VARIANT_TMA=$v; fi
( timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu ) > gpurun_out/bench_var_$v.log 2>&1
echo "variant_tma=$v $(tail -1 gpurun_out/bench_var_$v.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"],2), [round(s["ms"],2) for s in d["roofline"]["per_stage"]], d["clocks"]["sm_mhz"])')"
done
