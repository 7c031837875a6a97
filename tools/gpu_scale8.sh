#!/bin/bash
# N-GPU scaling points (run under `gpurun --gpus N`). Usage: bash tools/gpu_scale8.sh <tag> <N>
tag=$1; N=$2
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --no-cpu-baseline "$@" > $out/$name.json 2> $out/$name.err
  python - "$out/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "q/s=%.0f e2e=%.0f ms/step=%.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"]), d["parity"], d["config"]["parallelism"][:50], {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  grep -v "^\*\|OMP_NUM_THREADS\|^$" $out/$name.err | tail -3; }
run glove_lists_push --steps 50 --warmup 3 --shard lists --exchange push
run ivf100m_lists_push --workload ivf100m --n-probes 32 --steps 10 --warmup 3 --shard lists --exchange push
run ivf100m_lists_nccl --workload ivf100m --n-probes 32 --steps 10 --warmup 3 --shard lists --exchange nccl
