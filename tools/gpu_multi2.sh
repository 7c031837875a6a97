#!/bin/bash
# N-GPU default bench (100M, lists sharded) + the 2-process NCCL / peer-memory parity test. Usage under `gpurun --gpus N`: bash tools/gpu_multi2.sh <tag> <N> [bench args]
tag=${1:-multi}; N=${2:-2}; shift 2; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_tc_scan_gpu.py -x -q > $out/pytest_multi.log 2>&1; tail -3 $out/pytest_multi.log
timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 "$@" > $out/bench_n$N.json 2> $out/bench_n$N.err
python - $out/bench_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print("N=%d q/s=%.0f ms/step=%.2f e2e=%.0f %s"%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['kernel']), {k:round(v,2) for k,v in r['stage_ms'].items()}, {k:v for k,v in d['parity'].items() if 'mismatch' in k and v}, r.get('tc_role_cycles_per_tile'))
PY
grep -v "^W\|^\*\*\*" $out/bench_n$N.err | grep -i "error\|Traceback" | head -5
