#!/bin/bash
# Multi-GPU check (run under `gpurun --gpus N`): NCCL/peer parity test + bench at N GPUs, both exchanges and replicas.
# Usage: bash tools/gpu_multi.sh <tag> <N> [bench args]
tag=$1; N=$2; shift 2
EXTRA="$*"
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi -L > $out/gpus.txt; nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q > $out/pytest_sharded.log 2>&1; tail -3 $out/pytest_sharded.log
for mode in "lists push" "lists nccl" "replicas push"; do
  shard=${mode% *}; ex=${mode#* }
  name=${N}_${shard}_${ex}
  cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py"
  timeout 600 $cmd --gpus $N --steps 50 --warmup 3 --no-cpu-baseline --shard $shard --exchange $ex $EXTRA > $out/bench_$name.json 2> $out/bench_$name.err
  python - "$out/bench_$name.json" "gpus=$N shard=$shard exchange=$ex" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), d["parity"], d["config"]["parallelism"][:60], {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  tail -3 $out/bench_$name.err
done
