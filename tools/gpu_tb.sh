#!/bin/bash
# tests + one bench line. Usage: bash tools/gpu_tb.sh <tag> [bench args]
tag=${1:-tb}; shift
out=gpurun_out/$tag; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline "$@" > $out/bench.json 2> $out/bench.err
python - "$out/bench.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print("q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), d["parity"], "frac=%.3f"%r["frac"], {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print("FAILED", e)
PY
tail -3 $out/bench.err
