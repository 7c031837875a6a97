#!/bin/bash
# GPU tests + encoder throughput + a short default bench. Usage: bash tools/gpu_enc.sh <tag>
tag=${1:-enc}; out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -15 $out/pytest_gpu.log
timeout 600 python tools/encode_bench.py 10000000 > $out/encode_bench.json 2> $out/encode_bench.err; cat $out/encode_bench.json; tail -3 $out/encode_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_glove.json 2> $out/bench_glove.err
python - $out/bench_glove.json glove <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[2], "q/s=%.0f e2e=%.0f frac=%.3f"%(d["value"],d["e2e"]["value"],r["frac"]), d["parity"], {k:round(v,3) for k,v in r["stage_ms"].items()})
PY
