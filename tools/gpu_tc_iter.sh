#!/bin/bash
# tensor-core scan iteration: parity tests, one bench line, the phase probe. Usage: bash tools/gpu_tc_iter.sh <tag> [probe variants]
tag=${1:-tci}; out=gpurun_out/$tag; mkdir -p $out
bash tools/gpu_tc_quick.sh $tag
timeout 600 python tools/tc_probe.py --variants ${2:-0,32,1,4,5,13,61} > $out/tc_probe.jsonl 2> $out/tc_probe.err; cat $out/tc_probe.jsonl; tail -3 $out/tc_probe.err
