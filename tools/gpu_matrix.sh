#!/bin/bash
# The benchmark matrix of BASELINE.json's configs on one GPU: C1, C2 n_probes sweep with recall, C3 both orders,
# C4 batch sweep, C5 n_probes 8/128. Usage: bash tools/gpu_matrix.sh <tag>
tag=${1:-matrix}; out=gpurun_out/$tag; mkdir -p $out
B="--steps 10 --warmup 3 --no-cpu-baseline --parity-queries 500 --recall-queries 500 --no-e2e-pipeline"
timeout 600 python tools/c1_example.py > $out/c1.json 2> $out/c1.err; tail -c 1500 $out/c1.json; tail -2 $out/c1.err
for np_ in 1 2 4 8 16 32; do
  timeout 300 python bench.py --workload glove --n-probes $np_ $B > $out/c2_np$np_.json 2> $out/c2_np$np_.err
done
timeout 300 python bench.py --workload sift --order avx $B > $out/c3_avx.json 2> $out/c3_avx.err
timeout 300 python bench.py --workload sift --order sse $B > $out/c3_sse.json 2> $out/c3_sse.err
timeout 600 python bench.py --workload ivf100m --n-probes 8 $B > $out/c5_np8.json 2> $out/c5_np8.err
timeout 600 python bench.py --workload ivf100m --n-probes 128 --steps 4 --warmup 3 --no-cpu-baseline --parity-queries 200 --recall-queries 500 --no-e2e-pipeline > $out/c5_np128.json 2> $out/c5_np128.err
python - $out <<'PY'
import json,sys,glob,os
for f in sorted(glob.glob(sys.argv[1]+"/c[235]_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d["roofline"]; p=d["parity"]
        print(os.path.basename(f), "q/s=%.0f e2e=%.0f recall=%.3f scan_ms=%.3f kernel=%s frac=%.3f"%(d["value"],d["e2e"]["value"],(d.get("recall") or {}).get("value",-1),r["kernel_ms"],r["kernel"],r["frac"]),
              "mismatch:", {k:v for k,v in p.items() if "mismatch" in k and v})
    except Exception as e: print(os.path.basename(f),"FAILED",e)
PY
timeout 900 python tools/scan_sweep.py --batches 1,16,64,256,1024,4096 --reps 3 > $out/c4_sweep.jsonl 2> $out/c4.err; cut -c1-420 $out/c4_sweep.jsonl; tail -2 $out/c4.err
