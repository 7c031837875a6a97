"""The SOURCE of tkb_assign.cu's kernels executed on the CPU (tests/emulate/cuda_emu.h: one OS thread per CUDA thread, real
barriers, emulated warp shuffles) and checked against the oracle. The kernel was written after round 1's GPU budget was
spent; this is what stands in for a GPU run of its logic (tiling, barrier placement, top-k merge, edge tiles) until then."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import restate as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emulate")


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(EMU, "_build", "libassign_emu.so")
    src = [os.path.join(EMU, "assign_emu.cpp"), os.path.join(EMU, "cuda_emu.h"), os.path.join(ROOT, "tinyknn_b200", "csrc", "tkb_assign.cu")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-mfma", "-DTKB_EMULATE", "-I", EMU, "-shared", "-fPIC",
               src[0], "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("cannot compile the emulation harness: " + r.stderr[-300:])
    L = ctypes.CDLL(out)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.emu_assign_f32.argtypes = L.emu_assign_f64.argtypes = [vp, i64, i32, vp, i32, vp, vp, i32, vp]
    L.emu_row_sqnorm_f32.argtypes = [vp, i64, i32, vp]
    return L


@pytest.mark.parametrize("n,d,C,dtype", [(300, 100, 150, np.float32), (129, 20, 37, np.float32), (128, 16, 64, np.float32),
                                         (70, 7, 3, np.float32), (200, 128, 65, np.float64)])
def test_assign_kernel_source_on_cpu(emu, n, d, C, dtype):
    rng = np.random.default_rng(n + d + C)
    means = rng.standard_normal((C, d)) * 2
    X = np.ascontiguousarray(means[rng.integers(C, size=n)] + rng.standard_normal((n, d)), dtype=dtype)
    Y = np.ascontiguousarray(means + 0.1 * rng.standard_normal((C, d)), dtype=dtype)
    xn, yn = np.einsum("ij,ij->i", X, X), np.einsum("ij,ij->i", Y, Y)
    part = xn[:, None] + yn[None] - 2 * X @ Y.T                                   # utils.py:80-83
    fn = emu.emu_assign_f32 if dtype == np.float32 else emu.emu_assign_f64
    for k in (1, 2):
        if k > C:
            continue
        out = np.full((n, k), -7, np.int32)
        fn(X.ctypes.data, n, d, Y.ctypes.data, C, xn.ctypes.data, yn.ctypes.data, k, out.ctypes.data)
        if k == 1:
            assert np.array_equal(out[:, 0], O.knn_brute(X, Y, 1)[:, 0])
        else:
            assert np.array_equal(np.sort(out, axis=1), np.sort(O.knn_brute(X, Y, 2), axis=1))
            assert np.all(part[np.arange(n), out[:, 0]] <= part[np.arange(n), out[:, 1]])
            assert np.array_equal(out[:, 0], O.knn_brute(X, Y, 1)[:, 0])


def test_row_sqnorm_kernel_source_on_cpu(emu):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((300, 2)).astype(np.float32)                          # dpb-sized rows: einsum == mul-add
    out = np.empty(300, np.float32)
    emu.emu_row_sqnorm_f32(X.ctypes.data, 300, 2, out.ctypes.data)
    assert np.array_equal(out, np.einsum("ij,ij->i", X, X))
