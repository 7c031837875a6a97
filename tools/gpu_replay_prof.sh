#!/bin/bash
# One bench line, then an ncu full capture (with source) of the two replay launches of one timed step (probe selection, inverted lists).
# Usage: bash tools/gpu_replay_prof.sh <tag>
tag=${1:-rqp}; out=gpurun_out/$tag; mkdir -p $out
bash tools/gpu_tc_quick.sh $tag
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:replay_rq2 -c 2 \
    -o $out/replay_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 64 --recall-queries 0 --no-e2e-pipeline > $out/ncu_replay.log 2>&1; tail -1 $out/ncu_replay.log
