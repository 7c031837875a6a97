#!/bin/bash
# ncu --set full capture of one kernel inside the bench's timed region. Usage: bash tools/gpu_ncu.sh <tag> <kernel regex> [skip] [count] [bench args...]
tag=$1; kern=$2; skip=${3:-0}; cnt=${4:-2}; shift 4
out=gpurun_out/$tag
mkdir -p $out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$kern -s $skip -c $cnt \
    -o $out/$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $out/ncu.log 2>&1
tail -3 $out/ncu.log
