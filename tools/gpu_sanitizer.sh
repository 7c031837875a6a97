#!/bin/bash
# compute-sanitizer on hardware: memcheck over the kernel-level GPU tests, racecheck (shared-memory hazards) over the
# replay / scan tests. Usage: bash tools/gpu_sanitizer.sh <tag>
tag=${1:-san}; out=gpurun_out/$tag; mkdir -p $out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 420 $S --tool memcheck --error-exitcode 9 --log-file $out/memcheck.log \
    python -m pytest tests/test_tc_scan_gpu.py tests/test_gpu_parity.py -x -q -m gpu -k "tc_scan or fast_scan_bit_exact or heap_arrays or replay_kernels or replay_cm or lut_bytes or estimate_random or ivf_replay_fresh or plan_device or ivf_device_order" > $out/memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -2 $out/memcheck_pytest.log; tail -3 $out/memcheck.log
timeout 420 $S --tool racecheck --error-exitcode 9 --log-file $out/racecheck.log \
    python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "heap_arrays or replay_kernels or replay_cm or fast_scan_bit_exact or ivf_device_order" > $out/racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -2 $out/racecheck_pytest.log; tail -3 $out/racecheck.log
