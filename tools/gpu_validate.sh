#!/bin/bash
# Last call of a round: the whole GPU suite, smoke(), racecheck over the replay tests, the default bench line.
# Usage: bash tools/gpu_validate.sh <tag>
tag=${1:-validate}; out=gpurun_out/$tag; mkdir -p $out
timeout 1200 python -m pytest tests/ -x -q -m gpu > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
timeout 420 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $out/racecheck.log \
    python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "heap_arrays or replay_kernels or replay_cm or fast_scan_bit_exact or ivf_device_order" > $out/racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -1 $out/racecheck_pytest.log; grep "RACECHECK SUMMARY" $out/racecheck.log; grep -c "Error: Race" $out/racecheck.log
timeout 900 python bench.py --no-cpu-baseline --steps 20 > $out/bench.json 2> $out/bench.err
python - $out/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d['roofline']
print("q/s=%.0f ms/step=%.2f e2e=%.0f"%(d['value'],d['ms_per_step'],d['e2e']['value']), {k:round(v,2) for k,v in r['stage_ms'].items()}, {k:v for k,v in d['parity'].items() if 'mismatch' in k and v})
PY
