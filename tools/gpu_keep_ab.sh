#!/bin/bash
# A/B of the flagged-chunk L2 pinning (TKB_SCAN_KEEP) on a large index: stage times + DRAM bytes of the scan launches.
# Usage: bash tools/gpu_keep_ab.sh <tag> <workload> <n_probes>
tag=$1; wl=$2; np_=$3
out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_glove.json 2> $out/bench_glove.err
python - $out/bench_glove.json glove <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[2], "q/s=%.0f e2e=%.0f frac=%.3f"%(d["value"],d["e2e"]["value"],r["frac"]), d["parity"], {k:round(v,3) for k,v in r["stage_ms"].items()})
PY
for keep in 0 1; do
  TKB_SCAN_KEEP=$keep timeout 900 python bench.py --workload $wl --n-probes $np_ --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_keep$keep.json 2> $out/bench_keep$keep.err
  python - $out/bench_keep$keep.json keep=$keep <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "q/s=%.0f e2e=%.0f frac=%.3f flagged=%s"%(d["value"],d["e2e"]["value"],r["frac"],r.get("flagged_chunks")), d["parity"], {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
  TKB_SCAN_KEEP=$keep timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv -k regex:ivf_scan -c 2 \
    --log-file $out/ncu_keep$keep.csv python bench.py --workload $wl --n-probes $np_ --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_keep$keep.log 2>&1
  grep -E "dram__|gpu__time" $out/ncu_keep$keep.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | head -6
done
