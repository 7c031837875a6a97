#!/bin/bash
# Pull vs push exchange on N GPUs (run under `gpurun --gpus N`). Usage: bash tools/gpu_pull.sh <tag> <N> [pytest: 0|1]
tag=${1:-pull}; N=${2:-2}; T=${3:-1}; out=gpurun_out/$tag; mkdir -p $out
if [ "$T" = 1 ]; then
  timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_gpu_build_and_batch.py -x -q -m gpu -k "sharded or pull or push" > $out/pytest_pull.log 2>&1; tail -3 $out/pytest_pull.log
fi
for ex in ${EXCHANGES:-pull push}; do
  timeout ${RUN_TIMEOUT:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --exchange $ex > $out/bench_n${N}_$ex.json 2> $out/bench_n${N}_$ex.err
  python - $out/bench_n${N}_$ex.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
    print("N=%d q/s=%.0f ms/step=%.2f e2e=%.0f %s"%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['parallelism'][:60]), {k:round(v,2) for k,v in r['stage_ms'].items()}, {k:v for k,v in d['parity'].items() if 'mismatch' in k and v})
except Exception as e: print("FAILED", e)
PY
  grep -v "^W\|^\*\*\*" $out/bench_n${N}_$ex.err | grep -i "error\|Traceback" | head -5
done
