#!/bin/bash
# other workloads. Usage: bash tools/gpu_wl.sh <tag>
tag=${1:-wl}; out=gpurun_out/$tag; mkdir -p $out
run() { name=$1; shift
  timeout 900 python bench.py --steps 20 --warmup 3 "$@" > $out/$name.json 2> $out/$name.err
  python - "$out/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]; cb=d.get("cpu_baseline") or {}
    print(sys.argv[2], "q/s=%.0f e2e=%.0f cpu=%s"%(d["value"],d["e2e"]["value"],cb.get("value")), d["parity"], "frac=%.3f"%r["frac"], {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  tail -2 $out/$name.err; }
run sift_avx --workload sift --no-cpu-baseline
run sift_sse --workload sift --order sse --no-cpu-baseline
run glove_sse --workload glove --order sse --no-cpu-baseline
