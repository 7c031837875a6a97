#!/bin/bash
out=gpurun_out/exp3; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -15 $out/pytest_gpu.log
for cfg in "0" "1"; do
  TKB_RQ2=$cfg timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $out/bench_rq2_$cfg.json 2> $out/bench_rq2_$cfg.err
  python - "$out/bench_rq2_$cfg.json" "rq2=$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[2], "q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), d["parity"], "frac=%.3f"%r["frac"], {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  tail -3 $out/bench_rq2_$cfg.err
done
