#!/bin/bash
# compute-sanitizer runs + the ncu full capture of the LISTS launch of the tensor-core scan (the second launch of a timed step; the
# first one scans the encoded centroids). Usage: bash tools/gpu_final.sh <tag>
tag=${1:-final}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ivf_scan_tc -s 1 -c 1 \
    -o $out/tc_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 64 --recall-queries 0 --no-e2e-pipeline > $out/ncu_tc.log 2>&1; tail -1 $out/ncu_tc.log
timeout 300 python tools/c1_example.py > $out/c1.json 2> $out/c1.err; tail -c 700 $out/c1.json
bash tools/gpu_sanitizer.sh $tag
