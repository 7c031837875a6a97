#!/bin/bash
# The round's evidence on one GPU: GPU tests, smoke, bench (both arms). Usage: bash tools/gpu_round.sh <tag>
tag=${1:-round}; out=gpurun_out/$tag; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; tail -c 4000 $out/bench.json; tail -14 $out/bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 1500 $out/bench_ref.json; tail -3 $out/bench_ref.err
