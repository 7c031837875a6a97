#!/bin/bash
# One GPU session: parity tests, bench (both arms), ncu launch list of the bench command, ncu --set full of the scan kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 3000 $out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 1500 $out/bench_ref.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launch.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ivf_scan -c 2 \
    -o $out/scan_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_full.log 2>&1
ls -la $out
