#!/bin/bash
# N-GPU default bench (100M, lists sharded). Usage under `gpurun --gpus N`: bash tools/gpu_multi2.sh <tag> <N> [bench args]
tag=${1:-multi}; N=${2:-2}; shift 2; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 "$@" > $out/bench_n$N.json 2> $out/bench_n$N.err
tail -c 3000 $out/bench_n$N.json; grep -v "^W\|^\*\*\*" $out/bench_n$N.err | tail -40
