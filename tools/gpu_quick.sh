#!/bin/bash
# Quick GPU check: parity tests + one bench line. Usage: bash tools/gpu_quick.sh <tag> [bench args...]
tag=${1:-q}; shift
out=gpurun_out/$tag
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > $out/bench.json 2> $out/bench.err; tail -c 2500 $out/bench.json; tail -5 $out/bench.err
