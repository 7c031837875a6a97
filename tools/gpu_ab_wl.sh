#!/bin/bash
# A/B of environment switches on another workload. Usage: bash tools/gpu_ab_wl.sh <tag> "<bench args>" "<VAR=..>" "<VAR=..>" ...
tag=$1; wl=$2; shift 2; out=gpurun_out/$tag; mkdir -p $out
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py $wl --steps 20 --warmup 3 --no-cpu-baseline --parity-queries 200 --recall-queries 0 --no-e2e-pipeline > $out/ab.json 2> $out/ab.err
  python - "$out/ab.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st=d["roofline"]["stage_ms"]
    print(sys.argv[2], "q/s=%.0f"%d["value"], {k:v for k,v in d["parity"].items() if "mismatch" in k and v}, " ".join("%s=%.3f"%(k,v) for k,v in st.items()))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done
