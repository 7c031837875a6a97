#!/bin/bash
# A/B one environment variable over the bench. Usage: bash tools/gpu_ab.sh <tag> <VAR> <v1> <v2> ... [-- bench args]
tag=$1; var=$2; shift 2
vals=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do vals+=("$1"); shift; done; [ "$1" == "--" ] && shift
out=gpurun_out/$tag; mkdir -p $out
for v in "${vals[@]}"; do
  env $var=$v timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline "$@" > $out/bench_$v.json 2> $out/bench_$v.err
  python - "$out/bench_$v.json" "$var=$v" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[2], "q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), "parity_bad=%s"%d["parity"]["id_set_mismatch"], "frac=%.3f"%r["frac"], {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  tail -2 $out/bench_$v.err
done
