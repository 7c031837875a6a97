#!/bin/bash
# The whole GPU suite, then the default bench line (with the CPU arm). Usage: bash tools/gpu_full.sh <tag> [bench args]
tag=${1:-full}; shift; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/ -x -q -m gpu > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
timeout 900 python bench.py "$@" > $out/bench.json 2> $out/bench.err
python - $out/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print("q/s=%.0f ms/step=%.2f e2e=%.0f"%(d['value'],d['ms_per_step'],d['e2e']['value']), {k:round(v,2) for k,v in r['stage_ms'].items()}, {k:v for k,v in d['parity'].items() if 'mismatch' in k and v}, (d.get('cpu_baseline') or {}).get('value'))
PY
tail -3 $out/bench.err
