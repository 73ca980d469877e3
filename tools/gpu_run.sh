#!/bin/bash
# One parametrised driver for everything that runs on the GPU box (under gpurun [--gpus N]):
#
#   tools/gpu_run.sh tests                 pytest -m gpu on the visible GPUs -> gpurun_out/pytest_gpu.log
#   tools/gpu_run.sh multi N               tests/test_gpu_multi.py combos of world size N -> gpurun_out/pytest_multi_N.log
#   tools/gpu_run.sh bench N [tag] [bench.py args...]   contract line at N GPUs -> gpurun_out/bench_<tag>_gN.json
#   tools/gpu_run.sh launches [args...]    ncu launch list of a short bench -> gpurun_out/launches.csv
#   tools/gpu_run.sh ncu REGEX SKIP COUNT OUT CMD...   ncu --set full of the kernels matching REGEX
#   tools/gpu_run.sh probe                 NVLink peer-store probe (tools/probe/peer_probe.cu) -> gpurun_out/peer_probe.txt
#
# Environment variables (B2F_PIPELINE, B2F_PIPE_SMS, B2F_FLAG_BARRIER, B2F_P2P, B2F_FUSED, ...) pass through.
set -u
mkdir -p gpurun_out
cmd=${1:-tests}; shift || true
run_bench() {   # N tag args...
    local N=$1 tag=$2; shift 2
    local out=gpurun_out/bench_${tag}_g${N}.json
    if [ "$N" = 1 ]; then
        timeout 900 python bench.py --gpus 1 "$@" > "$out" 2> "${out%.json}.err"
    else
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
            --master-port $((29600 + RANDOM % 300)) bench.py --gpus "$N" "$@" > "$out" 2> "${out%.json}.err"
    fi
    python - "$out" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    c = d['config']
    print("%s: %.1f GPoints/s %.2f ms/step fwd %.2f bwd %.2f ferr %.2g e2e %s | %s" % (
        sys.argv[1], d['value'], d['ms_per_step'], c['forward_ms_median'], c['backward_ms_median'], c['forward_max_err'],
        (d.get('e2e') or {}).get('ms_per_step'), ' '.join('%s:%.2fms' % (k['stage'], k['ms']) for k in d['roofline']['per_stage'])))
except Exception as exc:
    print(sys.argv[1], "no contract line:", exc)
    sys.stdout.write(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
}
case "$cmd" in
tests)
    timeout 1200 python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/pytest_gpu.log 2>&1
    tail -4 gpurun_out/pytest_gpu.log | cut -c1-300 ;;
multi)
    N=$1
    B2F_TEST_WORLD=$N timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -rs > gpurun_out/pytest_multi_$N.log 2>&1
    tail -6 gpurun_out/pytest_multi_$N.log | cut -c1-400 ;;
bench)
    N=$1; tag=${2:-run}; shift 2 || shift $#
    run_bench "$N" "$tag" "$@" ;;
launches)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu "$@" > gpurun_out/bench_under_ncu.log 2>&1
    python tools/launch_summary.py gpurun_out/launches.csv | tail -20 ;;
ncu)
    regex=$1; skip=$2; count=$3; out=$4; shift 4
    # the report stays on the box (gpurun brings back at most 64 MiB): what comes home is its summary
    ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c "$count" -f -o "/tmp/$out" "$@" > "gpurun_out/$out.log" 2>&1
    python tools/ncu_summary.py "/tmp/$out.ncu-rep" > "gpurun_out/$out.txt" 2>&1
    python tools/ncu_hot.py "/tmp/$out.ncu-rep" 24 >> "gpurun_out/$out.txt" 2>&1
    grep -E "^==|time  |dram_rd|dram_wr|occupancy|stalls" "gpurun_out/$out.txt" | cut -c1-220 ;;
probe)
    [ -x tools/probe/peer_probe.bin ] || nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/probe/peer_probe.bin tools/probe/peer_probe.cu
    timeout 300 tools/probe/peer_probe.bin > gpurun_out/peer_probe.txt 2>&1
    cat gpurun_out/peer_probe.txt ;;
*)
    echo "unknown subcommand $cmd"; exit 2 ;;
esac
