#!/bin/bash
# Multi-GPU check (run under `gpurun --gpus N`): NCCL parity test + bench at 1..N GPUs. Usage: bash tools/gpu_multi.sh <tag> <N> [bench args]
tag=$1; N=$2; shift 2
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q > $out/pytest_sharded.log 2>&1; tail -3 $out/pytest_sharded.log
for n in 1 $N; do
  for mode in lists replicas; do
    [ $n == 1 ] && [ $mode == replicas ] && continue
    if [ $n == 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py"; fi
    timeout 900 $cmd --gpus $n --steps 50 --warmup 3 --no-cpu-baseline --shard $mode "$@" > $out/bench_${n}_$mode.json 2> $out/bench_${n}_$mode.err
    python - "$out/bench_${n}_$mode.json" "gpus=$n shard=$mode" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), d["parity"], {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
    tail -3 $out/bench_${n}_$mode.err
  done
done
