#!/bin/bash
# fused kernel: parity tests, then bench A/B
out=gpurun_out/exp2; mkdir -p $out
timeout 900 python -m pytest tests/test_fused_gpu.py -x -q > $out/pytest_fused.log 2>&1; tail -15 $out/pytest_fused.log
for cfg in "0 8 4" "1 8 4" "1 8 8" "1 4 4" "1 16 4" "1 8 2"; do
  set -- $cfg
  TKB_FUSED=$1 TKB_FUSED_G=$2 TKB_FUSED_LPW=$3 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $out/bench_$1_$2_$3.json 2> $out/bench_$1_$2_$3.err
  python - "$out/bench_$1_$2_$3.json" "fused=$1 G=$2 LPW=$3" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[2], "q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), d["parity"], "flagged=%s"%r.get("flagged_chunks"), "frac=%.3f"%r["frac"], {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  tail -3 $out/bench_$1_$2_$3.err
done
