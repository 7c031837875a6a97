#!/bin/bash
# tensor-core scan: parity tests, one bench line, then an ncu full capture (with source) of ONE launch of the kernel: the launch of the
# stage-timing pass = the whole 10 000-query batch. Usage: bash tools/gpu_tc_prof.sh <tag>
tag=${1:-tcp}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python tools/tc_probe.py > $out/tc_probe.jsonl 2> $out/tc_probe.err; cat $out/tc_probe.jsonl; tail -3 $out/tc_probe.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ivf_scan_tc -s 1 -c 1 \
    -o $out/tc_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 64 --recall-queries 0 --no-e2e-pipeline > $out/ncu_tc.log 2>&1; tail -2 $out/ncu_tc.log
ls -la $out
