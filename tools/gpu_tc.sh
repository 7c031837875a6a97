#!/bin/bash
# tensor-core scan: parity tests, then the default bench with and without it. Usage: bash tools/gpu_tc.sh <tag>
tag=${1:-tc}; out=gpurun_out/$tag; mkdir -p $out
timeout 240 python -m pytest tests/test_tc_scan_gpu.py -x -q -s > $out/pytest_tc.log 2>&1; tail -30 $out/pytest_tc.log
if grep -q "failed\|error" $out/pytest_tc.log; then exit 0; fi
TKB_TC_SCAN=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --parity-queries 500 > $out/bench_tc.json 2> $out/bench_tc.err; tail -c 2500 $out/bench_tc.json; tail -12 $out/bench_tc.err
