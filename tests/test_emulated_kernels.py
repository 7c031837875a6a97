"""The library's CUDA SOURCES executed on the CPU (tests/emulate: launch syntax translated, every CUDA thread a fiber, real
__syncthreads / warp-collective semantics, "device" memory = host memory) and checked against the oracle -- what stands in for
a GPU in the container that builds this repo, and the first validation of code written after a round's GPU budget was spent.

Three layers:
  * the emulator checks itself (collectives, barriers, SIMD-in-a-word intrinsics against scalar definitions, and the
    conditions it must REPORT: a divergent barrier, a write past the end of dynamic shared memory);
  * the C ABI called directly with numpy buffers (scan, heap replay, probe planning) against the oracle;
  * the gpu-marked test files run unchanged with TKB_EMU=1 (tests/conftest.py installs tests/emulate/emu_torch.py: the package's
    host layer on CPU tensors + the emulated library), in a subprocess, minus the cases that are too slow without a GPU.
This is test infrastructure: the package itself never loads the emulated library and has no CPU path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import restate as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emulate"))


@pytest.fixture(scope="module")
def emu():
    import emu_lib
    try:
        emu_lib.load()
    except Exception as e:                                           # noqa: BLE001
        pytest.skip("cannot build the emulated library: %s" % str(e)[-400:])
    return emu_lib


# ---- the emulator itself -----------------------------------------------------------------------------------------------------

def test_emulator_warp_collectives(emu):
    L = emu.load()
    T = 96
    out = np.zeros((T, 8), np.uint32)
    assert L.emu_selftest_collectives(out.ctypes.data, T) == 0
    lane, warp = np.arange(T) % 32, np.arange(T) // 32
    ballot = sum(1 << i for i in range(32) if i % 3 == 0)
    assert np.all(out[:, 0] == ballot)
    assert np.array_equal(out[:, 1], 100 * warp + 5)
    assert np.array_equal(out[:, 2], lane ^ 16)
    assert np.array_equal(out[:, 3], np.where(lane >= 3, lane - 3, lane))
    assert np.array_equal(out[:, 4], np.where(lane + 30 < 32, lane + 30, lane))
    assert np.array_equal(out[:, 5], (warp == 1).astype(np.uint32))            # __any_sync is per warp
    assert np.all(out[:, 6] == 1)
    assert np.array_equal(out[:, 7], lane * (lane + 1) // 2)


def test_emulator_barriers_and_shared_memory(emu):
    L = emu.load()
    blocks, T, rounds = 3, 128, 3
    out = np.zeros((blocks, 2), np.int32)
    assert L.emu_selftest_barrier(out.ctypes.data, blocks, T, rounds) == 0
    assert np.all(out[:, 0] == sum(0 + r for r in range(rounds)))              # thread T-1 reads thread 0's slot after the barrier
    assert np.all(out[:, 1] == 1)                                              # __syncthreads_or: one thread's predicate reaches all
    out = np.zeros((blocks, 2), np.int32)
    assert L.emu_selftest_barrier(out.ctypes.data, blocks, T, 2) == 0
    assert np.all(out[:, 1] == 0)
    res = np.zeros(256, np.int32)
    assert L.emu_selftest_early_exit(res.ctypes.data, 256, 100) == 0           # 156 threads return before the barrier
    assert np.all(res[:100] == 100) and np.all(res[100:] == 0)


def test_emulator_reports_divergent_barrier_and_smem_overrun(emu):
    L = emu.load()
    out = np.zeros(4, np.int32)
    assert L.emu_selftest_divergent(out.ctypes.data) != 0
    assert b"deadlock" in L.emu_last_error()
    assert L.emu_selftest_overrun(256) != 0
    assert b"shared memory" in L.emu_last_error()
    assert L.emu_selftest_barrier(np.zeros((1, 2), np.int32).ctypes.data, 1, 64, 1) == 0     # and it recovers


def test_emulator_scheduling_orders_expose_a_race(emu):
    """TKB_EMU_ORDER picks which runnable thread goes next (fifo / lifo / random): a kernel with a race gives different results
    under different orders, a correct one does not -- the gpu-marked files pass under all three (run by hand, ~80 s each)."""
    emu.load()
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import emu_lib; L = emu_lib.load(); o = np.zeros(64, np.int32); "
            "assert L.emu_selftest_racy(o.ctypes.data) == 0; print(','.join(map(str, o)))" % os.path.join(ROOT, "tests", "emulate"))
    outs = {}
    for order in ("fifo", "lifo", "random:7"):
        r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, TKB_EMU_ORDER=order), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1500:]
        outs[order] = r.stdout.strip()
    assert len(set(outs.values())) == 3, outs


def test_emulator_ubsan_build_aborts_on_a_misaligned_vector_load(emu):
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import emu_lib; L = emu_lib.load(); b = emu_lib.aligned((64,), np.uint8); "
            "o = np.zeros(1, np.uint32); assert L.emu_selftest_vector_load(b.ctypes.data, o.ctypes.data) == 0; print('aligned ok', flush=True); "
            "L.emu_selftest_vector_load(b.ctypes.data + 4, o.ctypes.data); print('misaligned passed')" % os.path.join(ROOT, "tests", "emulate"))
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=dict(os.environ, TKB_EMU_UBSAN="1"), capture_output=True, text=True, timeout=900)
    assert "aligned ok" in r.stdout and "misaligned passed" not in r.stdout and r.returncode != 0, r.stdout + r.stderr[-800:]
    assert "misaligned address" in r.stderr or "alignment" in r.stderr, r.stderr[-800:]


def test_emulator_simd_intrinsics_match_scalar_definitions(emu):
    L = emu.load()
    rng = np.random.default_rng(0)
    n = 4000
    a, b, c = (rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32) for _ in range(3))
    edge = np.array([0, 0x7fff7fff, 0x80008000, 0xffffffff, 0x007f007f, 0xff80ff80, 0x7f7f7f7f, 0x80808080], np.uint32)
    a[:8], b[8:16], c[16:24] = edge, edge, edge
    out = np.zeros((n, 12), np.uint32)
    assert L.emu_selftest_simd(a.ctypes.data, b.ctypes.data, c.ctypes.data, out.ctypes.data, n) == 0

    def halves(x, signed):
        h = np.stack([x & 0xffff, x >> 16], 1).astype(np.int64)
        return np.where(h >= 0x8000, h - 0x10000, h) if signed else h

    def pack2(h):
        h = h.astype(np.int64) & 0xffff
        return (h[:, 0] | (h[:, 1] << 16)).astype(np.uint32)

    def bytes_(x, signed):
        v = np.stack([(x >> (8 * i)) & 0xff for i in range(4)], 1).astype(np.int64)
        return np.where(v >= 0x80, v - 0x100, v) if signed else v

    def pack4(v):
        v = v.astype(np.int64) & 0xff
        return (v[:, 0] | (v[:, 1] << 8) | (v[:, 2] << 16) | (v[:, 3] << 24)).astype(np.uint32)

    wrap16 = lambda h: ((h + 0x8000) % 0x10000) - 0x8000                                    # noqa: E731
    assert np.array_equal(out[:, 0], pack2(halves(a, False) + halves(b, False)))
    assert np.array_equal(out[:, 1], pack2(np.minimum(np.minimum(halves(a, True), halves(b, True)), halves(c, True))))
    assert np.array_equal(out[:, 2], pack2(np.maximum(np.maximum(halves(a, True), halves(b, True)), halves(c, True))))
    assert np.array_equal(out[:, 3], pack2(np.minimum(np.minimum(halves(a, False), halves(b, False)), halves(c, False))))
    assert np.array_equal(out[:, 4], pack2(np.maximum(wrap16(halves(a, True) + halves(b, True)), halves(c, True))))
    assert np.array_equal(out[:, 5], pack2(np.minimum((halves(a, False) + halves(b, False)) % 0x10000, halves(c, False))))
    assert np.array_equal(out[:, 6], pack4(np.minimum(bytes_(a, True), bytes_(b, True))))
    assert np.array_equal(out[:, 7], pack4(np.minimum(bytes_(a, False), bytes_(b, False))))
    assert np.array_equal(out[:, 8], pack4(np.where(bytes_(a, True) < bytes_(b, True), 0xff, 0)))
    assert np.array_equal(out[:, 9], pack4(np.where(bytes_(a, False) < bytes_(b, False), 0xff, 0)))
    s32 = lambda x: x.astype(np.int64) - ((x.astype(np.int64) >> 31) << 32)                 # noqa: E731
    wrap32 = lambda v: ((v + 2 ** 31) % 2 ** 32) - 2 ** 31                                  # noqa: E731
    assert np.array_equal(out[:, 10], (np.minimum(wrap32(s32(a) + s32(b)), s32(c)) % 2 ** 32).astype(np.uint32))
    # prmt.b32, default mode (PTX ISA): selector nibble i picks byte (n & 7) of {a, b}; bit 3 replicates that byte's sign bit
    src = np.concatenate([bytes_(a, False), bytes_(b, False)], 1)
    sel = np.stack([(c >> (4 * i)) & 0xf for i in range(4)], 1).astype(np.int64)
    picked = np.take_along_axis(src, sel & 7, 1)
    picked = np.where(sel & 8, np.where(picked & 0x80, 0xff, 0), picked)
    assert np.array_equal(out[:, 11], pack4(picked))


# ---- the C ABI on numpy buffers ------------------------------------------------------------------------------------------------

def _tables(rng, M, signd, kind):
    if kind == "full":
        return rng.integers(0, 256, size=(M, 16)).astype(np.uint8)
    if kind == "hot":                                                # small range, large values: many saturations
        t = rng.integers(8, 30, size=(M, 16)).astype(np.int16)
        return (t - (20 if signd else 0)).astype(np.int8).view(np.uint8)
    t = np.round(rng.exponential(6.0, size=(M, 16)) - 4).clip(-4, 128 / M ** 0.5).astype(np.int8)
    return t.view(np.uint8) if signd else (t + 4).astype(np.uint8)


@pytest.mark.parametrize("M", [52, 32, 8, 20])
def test_fast_scan_source_bit_exact(emu, M):
    """tkb_codes_to_native_dev + tkb_estimate_native_dev (PRMT lookups, deferred clamps, certificate, deferred exact pass) for
    the compile-time pair counts of BASELINE.json's shapes (M = 52, 32) and the generic loop, both orders, signed/unsigned."""
    L = emu.load()
    rng = np.random.default_rng(M)
    for order, signd, kind in [("avx", 1, "lut"), ("avx", 1, "hot"), ("avx", 1, "full"), ("avx", 0, "lut"), ("sse", 1, "lut"),
                               ("sse", 0, "hot")]:
        if order == "avx" and M % 4:
            continue
        n = int(rng.integers(1, 3000))
        codes = rng.integers(0, 16, size=(-(-n // 16) * 16, M), dtype=np.uint8)
        packed = np.ascontiguousarray(O.transform_data(codes))
        nch = len(packed)
        Q = 3
        tabs = emu.aligned((Q, M, 16), np.uint8)
        tabs[:] = np.stack([_tables(rng, M, signd, kind) for _ in range(Q)])
        nat = emu.aligned((-(-nch // 8) * 8 * M * 8,), np.uint8)
        emu.check(L.tkb_codes_to_native_dev(emu.ptr(packed), nch, M, emu.ptr(nat), None))
        back = np.zeros_like(packed)
        emu.check(L.tkb_codes_from_native_dev(emu.ptr(nat), nch, M, emu.ptr(back), None))
        assert np.array_equal(back, packed)
        est, ws = emu.aligned((Q, 16 * nch), np.uint8), emu.aligned((64,), np.uint8)
        emu.check(L.tkb_estimate_native_dev(emu.ptr(nat), nch, M, emu.ptr(tabs), Q, emu.ptr(est), 16 * nch,
                                            1 if order == "avx" else 0, signd, emu.ptr(ws), 64, None))
        for q in range(Q):
            exp = np.zeros(2 * nch, np.uint64)
            O.estimate_pq(packed, O.transform_tables(tabs[q]), exp, bool(signd), order)
            assert np.array_equal(est[q], exp.view(np.uint8)), (M, order, signd, kind, q)


def test_emulated_abi_reports_errors_like_the_real_one(emu):
    L = emu.load()
    assert L.tkb_estimate_native_dev(None, 4, 6, None, 1, None, 64, 1, 1, None, 0, None) == 1          # avx order needs M % 4 == 0
    assert b"M % 4" in L.tkb_last_error()


# ---- the gpu-marked test files on the emulator -------------------------------------------------------------------------------

# too slow without a GPU (minutes each); they pass when run by hand, the async batches (2.5 min) and the CUDA-graph batch (13 min:
# launches recorded with copies of their arguments and replayed, synchronising calls refused during the capture) included
_SKIP_ON_EMULATOR = ("test_replay_deep_heaps and 70000", "test_glove_shape_full_size_properties", "test_sift_shape_full_size_both_orders", "test_graphed_batch_equals_eager",
                     "test_async_results_equal_sync", "test_estimate_large_bit_exact", "test_fast_scan_large_and_patch_rate",
                     "test_query_batch_with_chunk_minima_equals_plain", "test_encode_device_large_matches_oracle_and_scan_roundtrip")


def test_gpu_test_files_pass_on_the_emulator(emu):
    """tests/test_gpu_parity.py, test_fused_gpu.py and test_gpu_build_and_batch.py (code that first ran on
    hardware yet: coarse assignment kernel, chunk minima inside the push exchange, saved-index queries), executed with
    TKB_EMU=1: the product's host layer and kernel sources against the oracle and the golden fixtures."""
    emu.load()                                                       # build once, before the child starts
    # the child uses the UBSan build of the emulated library: a misaligned access through a vector type (uint4, float4, ...: a
    # "misaligned address" fault on the GPU) or an out-of-bounds index into a fixed-size array aborts the run
    env = dict(os.environ, TKB_EMU="1", TKB_EMU_UBSAN="1", TKB_RUN_UNVALIDATED="1", OMP_NUM_THREADS="2")
    k = " and ".join("not (%s)" % s for s in _SKIP_ON_EMULATOR)
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider", "-k", k,
           os.path.join(ROOT, "tests", "test_gpu_parity.py"), os.path.join(ROOT, "tests", "test_fused_gpu.py"),
           os.path.join(ROOT, "tests", "test_gpu_build_and_batch.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    lines = r.stdout.strip().splitlines()
    last = lines[-1]
    assert " passed" in last and "failed" not in last, tail
    assert int(last.split(" passed")[0].split()[-1]) >= 140, last
    # no kernel ever ran a full-mask warp collective after some of the warp's lanes had returned
    stats = [ln for ln in lines if ln.startswith("emulator:")]
    assert stats and "'collectives_with_exited_lanes': 0" in stats[-1], stats


REFERENCE_TESTS = "/root/reference/tests"


@pytest.mark.skipif(not os.path.isdir(REFERENCE_TESTS), reason="the reference checkout is only present in the build container")
def test_reference_own_test_files_pass_unmodified_on_the_emulator(emu):
    """Drop-in check: the reference's own test files (tests/test_pq.py, test_ivf.py, test_multiprobe.py, test_heap.py,
    test_transform.py, test_utils.py), collected from the read-only checkout and run UNMODIFIED with `tinyknn` aliased to this
    package on the emulator (tests/emulate/ref_alias_plugin.py): every kernel call goes through the C ABI into the CUDA sources."""
    emu.load()
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", OMP_NUM_THREADS="2",
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "emulate"), ROOT, os.environ.get("PYTHONPATH", "")]))
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [sys.executable, "-m", "pytest", "-q", "-p", "ref_alias_plugin", "-p", "no:cacheprovider", "--rootdir", tmp,
               "-c", os.devnull, "-W", "ignore", REFERENCE_TESTS]
        r = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
    tail = r.stdout[-3000:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    last = r.stdout.strip().splitlines()[-1]
    assert " passed" in last and "failed" not in last and "error" not in last, tail
    assert int(last.split(" passed")[0].split()[-1]) >= 120, last


def test_bench_dry_run_on_the_emulator_prints_the_contract_line(emu):
    """bench.py's main() end to end (index build, parity gate, timed loop, stage timing, e2e loop, CPU arm, JSON line) with
    the library and torch.cuda emulated (tests/emulate/emu_bench.py, `tiny` workload): the numbers mean nothing, the keys and
    the parity gate do -- a Python-level mistake in bench.py must not wait for the GPU box to show up."""
    import json
    emu.load()
    cmd = [sys.executable, os.path.join(ROOT, "tests", "emulate", "emu_bench.py"), "--workload", "tiny", "--queries", "128",
           "--steps", "2", "--warmup", "3", "--cpu-seconds", "1"]
    r = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="2"), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["steps"] == 2 and line["warmup"] == 3 and line["n_gpus"] == 1 and line["gpu_launches"] > 0
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and line["e2e"]["h2d_bytes_per_step"] == 128 * 100 * 4
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and line["roofline"]["bound"] == "hbm"
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["parity"]["id_set_mismatch"] == 0 and line["parity"]["device_order_id_set_mismatch"] == 0
