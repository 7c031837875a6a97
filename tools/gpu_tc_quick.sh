#!/bin/bash
# tensor-core scan: parity tests + one bench line (no CPU arm). Usage: bash tools/gpu_tc_quick.sh <tag>
tag=${1:-tcq}; out=gpurun_out/$tag; mkdir -p $out
timeout 240 python -m pytest tests/test_tc_scan_gpu.py -x -q > $out/pytest_tc.log 2>&1; tail -3 $out/pytest_tc.log
TKB_TC_SCAN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --parity-queries 200 --recall-queries 0 --no-e2e-pipeline > $out/bench_tc.json 2> $out/bench_tc.err
python - $out/bench_tc.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print("q/s=%.0f ms/step=%.2f scan=%.2f replay=%.2f"%(d['value'],d['ms_per_step'],r['stage_ms']['scan'],r['stage_ms']['replay']), {k:v for k,v in d['parity'].items() if 'mismatch' in k and v}, r.get('tc_tiles'), r.get('tc_mean_group_columns'), d['clocks']['sm_mhz'], r.get('tc_role_cycles_per_tile'))
PY
