#!/bin/bash
# ncu --set full for every kernel of ONE step. Usage: bash tools/gpu_full_all.sh <tag> [bench args]
tag=${1:-fullall}; shift
out=gpurun_out/$tag; mkdir -p $out
TKB_STREAMS=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -c 16 \
    -o $out/step python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > $out/ncu.log 2>&1
tail -2 $out/ncu.log | cut -c1-200
ls -la $out
