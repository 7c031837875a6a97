#!/bin/bash
# tests + glove bench + sift (both orders) + scan sweep (short)
tag=${1:-tb}; shift
out=gpurun_out/$tag; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d["roofline"]
    print(sys.argv[2], "q/s=%.0f e2e=%.0f"%(d["value"],d["e2e"]["value"]), "bad=%s/%s"%(d["parity"]["id_set_mismatch"],d["parity"].get("device_order_id_set_mismatch")), "frac=%.3f flagged=%s"%(r["frac"],r.get("flagged_chunks")), {k:round(v,3) for k,v in r["stage_ms"].items()})
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $out/bench.json 2> $out/bench.err; show $out/bench.json glove; tail -2 $out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload sift > $out/sift.json 2> $out/sift.err; show $out/sift.json sift_avx; tail -2 $out/sift.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload sift --order sse > $out/sift_sse.json 2> $out/sift_sse.err; show $out/sift_sse.json sift_sse; tail -2 $out/sift_sse.err
timeout 900 python tools/scan_sweep.py --batches 1,16,256 > $out/sweep.jsonl 2> $out/sweep.err; tail -2 $out/sweep.err
python - $out/sweep.jsonl <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l)
    if "batch" in d: print(d["batch"], "%.1f Gcodes/s"%(d["value"]/1e9), "ms=%.3f"%d["ms"], "q/s=%.0f"%d["queries_per_s"], "top q/s=%.0f"%d["top_queries_per_s"], "frac=%.3f"%d["roofline"]["frac"], "flag=",d["roofline"]["flagged_chunks"])
    else: print(d)
PY
