#!/bin/bash
# large synthetic indexes on one GPU. Usage: bash tools/gpu_big.sh <tag> <workload> <n_probes> [more bench args]
tag=$1; wl=$2; np_=$3; shift 3
out=gpurun_out/$tag; mkdir -p $out
( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader >> $out/mem.txt; sleep 5; done ) &
mon=$!
timeout 1500 python bench.py --workload $wl --n-probes $np_ --steps 10 --warmup 3 --no-cpu-baseline "$@" > $out/bench.json 2> $out/bench.err
kill $mon
python - "$out/bench.json" "$wl p=$np_" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[2], "q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), d["parity"], "frac=%.3f flagged=%s scanned=%s"%(r["frac"],r.get("flagged_chunks"),r.get("scanned_vectors_per_launch")), {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
grep -v "^\s" $out/bench.err | tail -5; sort -n $out/mem.txt | tail -1
