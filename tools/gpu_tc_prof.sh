#!/bin/bash
# tensor-core scan: one bench line + ncu full capture of the kernel. Usage: bash tools/gpu_tc_prof.sh <tag>
tag=${1:-tcp}; out=gpurun_out/$tag; mkdir -p $out
TKB_TC_SCAN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --parity-queries 200 --no-e2e-pipeline > $out/bench_tc.json 2> $out/bench_tc.err; tail -c 1800 $out/bench_tc.json; tail -4 $out/bench_tc.err
TKB_TC_SCAN=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivf_scan_tc -c 1 \
    -o $out/tc_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 64 --no-e2e-pipeline > $out/ncu_tc.log 2>&1; tail -2 $out/ncu_tc.log
