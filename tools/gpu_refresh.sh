#!/bin/bash
# Round-end evidence on one GPU: default bench line (with the CPU arm), the reference arm, the ncu launch list and the full capture
# of the dominant kernel for the same command, the config matrix, two A/Bs. Usage: bash tools/gpu_refresh.sh <tag>
tag=${1:-refresh}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.json; tail -2 $out/bench.err
timeout 600 python bench.py --impl reference > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 600 $out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 64 --recall-queries 0 --no-e2e-pipeline > $out/launches.log 2>&1; tail -1 $out/launches.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ivf_scan_tc -s 1 -c 1 \
    -o $out/tc_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --parity-queries 64 --recall-queries 0 --no-e2e-pipeline > $out/ncu_tc.log 2>&1; tail -1 $out/ncu_tc.log
for cfg in "TKB_COARSE_FUSED=1" "TKB_RQ_INDEP=1"; do
  env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --parity-queries 200 --recall-queries 0 --no-e2e-pipeline > $out/ab.json 2> $out/ab.err
  python - "$out/ab.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st=d["roofline"]["stage_ms"]
    print(sys.argv[2], "q/s=%.0f"%d["value"], {k:v for k,v in d["parity"].items() if "mismatch" in k and v}, " ".join("%s=%.3f"%(k,v) for k,v in st.items()))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done
bash tools/gpu_matrix.sh $tag/matrix
