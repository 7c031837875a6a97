#!/bin/bash
# GPU tests, then A/B of the chunk-minimum replay on a large index. Usage: bash tools/gpu_cm_ab.sh <tag> <workload> <n_probes>
tag=$1; wl=$2; np_=$3
out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -12 $out/pytest_gpu.log
for cm in 0 8192; do
  TKB_CMIN_CHUNKS=$cm timeout 900 python bench.py --workload $wl --n-probes $np_ --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_cm$cm.json 2> $out/bench_cm$cm.err
  python - $out/bench_cm$cm.json cm=$cm <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "q/s=%.0f e2e=%.0f frac=%.3f"%(d["value"],d["e2e"]["value"],r["frac"]), d["parity"], {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  tail -2 $out/bench_cm$cm.err
done
