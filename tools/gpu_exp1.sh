#!/bin/bash
# experiment: stream overlap A/B + ncu of the replay kernel
out=gpurun_out/exp1; mkdir -p $out
for cfg in "1 2500" "2 5000" "2 2500" "3 2500" "2 1250" "4 1250"; do
  set -- $cfg
  TKB_STREAMS=$1 TKB_SUB_QUERIES=$2 timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $out/bench_$1_$2.json 2> $out/bench_$1_$2.err
  python - "$out/bench_$1_$2.json" "streams=$1 sub=$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), "parity_bad=%s"%d["parity"]["id_set_mismatch"], "flagged=%s"%d["roofline"].get("flagged_chunks"), {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  tail -3 $out/bench_$1_$2.err
done
TKB_STREAMS=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:replay_rq -s 1 -c 1 \
    -o $out/rq_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_rq.log 2>&1
tail -2 $out/ncu_rq.log
