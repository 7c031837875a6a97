#!/bin/bash
# FIRST GPU call of round 2 (one GPU): the tests that were written after round 1's GPU budget was spent, then the regular suite,
# smoke, the default bench (100M) and the GloVe-shape A/Bs. Usage (from the repo root, under gpurun): bash tools/gpu_next.sh <tag>
tag=${1:-next}; out=gpurun_out/$tag; mkdir -p $out
G="--workload glove --parity-queries 256"
TKB_RUN_UNVALIDATED=1 timeout 900 python -m pytest tests/test_unvalidated_gpu.py -q > $out/pytest_unvalidated.log 2>&1; tail -25 $out/pytest_unvalidated.log
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -c 3500 $out/bench.json; tail -30 $out/bench.err
timeout 900 python bench.py --steps 20 --warmup 3 $G > $out/bench_glove.json 2> $out/bench_glove.err; tail -c 3500 $out/bench_glove.json; tail -3 $out/bench_glove.err
timeout 900 python bench.py --steps 20 --warmup 3 $G --no-cpu-baseline --graph > $out/bench_graph.json 2> $out/bench_graph.err; tail -c 1200 $out/bench_graph.json; tail -3 $out/bench_graph.err
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:replay_rq2 -c 4 \
    -o $out/replay_full python bench.py --steps 2 --warmup 3 $G --no-cpu-baseline --no-e2e-pipeline > $out/ncu_replay.log 2>&1; tail -2 $out/ncu_replay.log
# A/Bs prepared in round 1 (all results must stay identical: the bench's parity gate runs every time)
for cfg in "TKB_WORKSPACE_REUSE=0" "TKB_COARSE_FUSED=1" "TKB_COARSE_FUSED=1 TKB_SUB_QUERIES=2500 TKB_STREAMS=3" "TKB_RQ_MIN_CTAS=1184 TKB_RQ_QPW=1" "TKB_RQ_MIN_CTAS=592 TKB_RQ_QPW=2" "TKB_RQ_LANES=8" "TKB_SCAN_THREADS=64" "TKB_SUB_QUERIES=2500" "TKB_SUB_QUERIES=2500 TKB_STREAMS=3"; do
  env $cfg timeout 600 python bench.py --steps 20 --warmup 3 $G --no-cpu-baseline > $out/bench_rq.json 2> $out/bench_rq.err
  python - "$out/bench_rq.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); st=d["roofline"]["stage_ms"]
    print(sys.argv[2], "q/s=%.0f e2e=%.0f pipe=%.0f"%(d["value"],d["e2e"]["value"],d.get("e2e_pipelined",{}).get("value",0)), d["parity"]["id_set_mismatch"], " ".join("%s=%.3f"%(k,v) for k,v in st.items()))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done
