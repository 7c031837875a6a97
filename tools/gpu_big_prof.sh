#!/bin/bash
# 100M (or 10M) workload on one GPU: bench line, ncu launch list, ncu --set full of the scan kernel.
# Usage: bash tools/gpu_big_prof.sh <tag> <workload> <n_probes>
tag=$1; wl=$2; np_=$3
out=gpurun_out/$tag; mkdir -p $out
timeout 1500 python bench.py --workload $wl --n-probes $np_ --steps 10 --warmup 3 --no-cpu-baseline > $out/bench.json 2> $out/bench.err
tail -c 2500 $out/bench.json; tail -3 $out/bench.err
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/launches.csv python bench.py --workload $wl --n-probes $np_ --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launch.log 2>&1
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivf_scan -c 2 \
    -o $out/scan_full python bench.py --workload $wl --n-probes $np_ --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1
ls -la $out
