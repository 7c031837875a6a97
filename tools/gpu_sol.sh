#!/bin/bash
# Speed-of-light section for every kernel of one step. Usage: bash tools/gpu_sol.sh <tag> [bench args]
tag=${1:-sol}; shift
out=gpurun_out/$tag; mkdir -p $out
TKB_STREAMS=1 timeout 900 ncu --profile-from-start off --section SpeedOfLight --section LaunchStats --section Occupancy --clock-control none --csv \
    --log-file $out/sol.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > $out/sol.log 2>&1
tail -2 $out/sol.log | cut -c1-300
