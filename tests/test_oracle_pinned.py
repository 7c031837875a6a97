"""Pins the oracle (oracle/pq_oracle.c + oracle/restate.py) -- CPU only.

 (a) against the reference's own known-answer tests for the path (restated from
     /root/reference/tests/test_transform.py and tests/test_pq.py::test_estimate_pq_simd);
 (b) against the compiled reference kernels in oracle/_ref, when present;
 (c) against the committed golden fixtures (tests/golden/*.npz, produced by the reference);
 (d) against the reference Python package itself, when /root/reference is present (build container).
"""
import math
import random
from functools import reduce
from itertools import product

import numpy as np
import pytest

from oracle import restate as O, ref_loader


# ---- (a) known-answer tests of the reference ----------------------------------------------------

def _slow_pq(codes, tab, signed, order="sse"):
    packed = O.transform_data(codes)
    out = np.zeros(2 * len(packed), dtype=np.uint64)
    O.estimate_pq(packed, O.transform_tables(tab), out, signed, order)
    return out.view(np.uint8)


def test_simple_identity_tables():                       # ref: tests/test_transform.py:10-17
    dat = np.array([[1, 3, 7, 15]] + [[0, 0, 0, 0]] * 15, dtype=np.uint8)
    tab = np.array([list(range(16)) for _ in range(4)], dtype=np.uint8)
    for order in ("sse", "avx"):
        out = _slow_pq(dat, tab, False, order)
        assert out[0] == 1 + 3 + 7 + 15 and not out[1:].any()


def _sat8(x, y):
    return max(-128, min(127, x + y))


def test_rand_python_model():                            # ref: tests/test_transform.py:20-58
    rnd = random.Random(5)
    for i, j in product(range(1, 10, 2), range(1, 10)):
        n, d = 16 * i, 2 * j
        dat = [[rnd.randrange(16) for _ in range(d)] for _ in range(n)]
        tab = [[rnd.randrange(256 // d * 2) for _ in range(16)] for _ in range(d)]
        exp = np.minimum([sum(tab[c][dat[r][c]] for c in range(d)) for r in range(n)], 255)
        assert np.array_equal(exp, _slow_pq(np.array(dat, np.uint8), np.array(tab, np.uint8), False))
        top = int(math.floor(127 / d ** 0.5))
        tab = [[rnd.randrange(-top, top) for _ in range(16)] for _ in range(d)]
        exp = [reduce(_sat8, (tab[c][dat[r][c]] for c in range(d))) for r in range(n)]
        got = _slow_pq(np.array(dat, np.uint8), np.array(tab).astype(np.uint8), True).astype(np.int8)
        assert np.array_equal(exp, got)


@pytest.mark.parametrize("n,d,signed,order", list(product([16, 32], [4, 8], [True, False], ["sse", "avx"])))
def test_estimate_orders(n, d, signed, order):           # ref: tests/test_pq.py:12-53 (avx lane rule j & 2)
    rng = np.random.default_rng(n * d + signed)
    data = rng.integers(0, 16, size=(n, d), dtype=np.uint8)
    tables = rng.integers(0, 256, size=(d, 16)).astype(np.uint8)
    bt = np.int8 if signed else np.uint8
    lo, hi = (-128, 127) if signed else (0, 255)
    tv = tables.view(bt)
    exp = np.zeros(n, dtype=bt)
    for i, row in enumerate(data):
        acc = [0, 0]
        for j, c in enumerate(row):
            lane = 0 if order == "sse" or j & 2 == 0 else 1
            acc[lane] = int(np.clip(acc[lane] + int(tv[j][c]), lo, hi))
        exp[i] = acc[0] if order == "sse" else np.clip(acc[0] + acc[1], lo, hi)
    assert np.array_equal(_slow_pq(data, tables, signed, order).view(bt), exp)


def test_transform_layout():                             # ref: tests/test_transform.py:71-101
    rng = np.random.default_rng(10)
    n, d = 16 * 13, 2 * 7
    data0 = rng.integers(16, size=(n, d)).astype(np.uint8)
    data = O.transform_data(data0)
    assert data.shape == (n // 16, d)
    assert np.array_equal(O.unpack(data), data0)
    shifts = np.arange(15, -1, -1, dtype=np.uint64) * 4
    nib = (data[..., np.newaxis] >> shifts) & 0xF
    assert nib[0, 0, -1] == data0[0][0] and nib[0, 0, -2] == data0[0][1]
    assert nib[0, 0, -3] == data0[1][0] and nib[0, 0, -4] == data0[1][1]
    assert nib[0, 1, -1] == data0[8][0] and nib[0, 1, -2] == data0[8][1]
    assert nib[0, 2, -1] == data0[0][2] and nib[0, 2, -2] == data0[0][3]


def test_heap_known_answers():                           # ref: tests/test_heap.py:23-49
    hi, hv = np.empty(3, np.int64), np.empty(3, np.int32)
    O.init_heap(hi, hv, True)
    assert hi.tolist() == [-1] * 3 and hv.tolist() == [127] * 3
    hi, hv = np.empty(2, np.int64), np.empty(2, np.int32)
    O.init_heap(hi, hv, True)
    O.insert(hi, hv, 1, 10)
    assert hi.tolist() == [-1, 1] and hv.tolist() == [127, 10]
    O.insert(hi, hv, 1, 10)
    assert hi.tolist() == [-1, 1] and hv.tolist() == [127, 10]
    hi, hv = np.empty(4, np.int64), np.empty(4, np.int32)
    O.init_heap(hi, hv, False)
    assert hv.tolist() == [255] * 4


# ---- (b) compiled reference kernels ---------------------------------------------------------------

needs_ref = pytest.mark.skipif(not ref_loader.have_ref_kernels(), reason="oracle/_ref not built")


@needs_ref
def test_c_port_matches_compiled_reference():
    sse, avx = ref_loader.load_ref_kernels()
    rng = np.random.default_rng(0)
    for trial in range(160):
        order = ("sse", "avx")[trial % 2]
        signd = bool((trial // 2) % 2)
        M = int(rng.integers(1, 17)) * (4 if order == "avx" else 2)
        n = int(rng.integers(1, 300))
        n16 = -(-n // 16) * 16
        codes = rng.integers(0, 16, size=(n16, M), dtype=np.uint8)
        hi_ = (256, 40, 12)[trial % 3]
        tab = rng.integers(0, hi_, size=(M, 16)).astype(np.uint8)
        if signd and trial % 3:
            tab = (tab.astype(np.int16) - hi_ // 3).astype(np.int8).view(np.uint8)
        packed, T = O.transform_data(codes), O.transform_tables(tab)
        mod = avx if order == "avx" else sse
        o1, o2 = np.zeros(2 * len(packed), np.uint64), np.zeros(2 * len(packed), np.uint64)
        O.estimate_pq(packed, T, o1, signd, order)
        getattr(mod, "estimate_pq_" + order)(packed, T, o2, signd)
        assert np.array_equal(o1, o2)
        R = int(rng.integers(1, 40))
        labels = None
        if trial % 5 >= 2:
            labels = rng.permutation(10 ** 6)[:n16].astype(np.int64) + (10 ** 12 if order == "avx" else 0)
            if trial % 7 == 0:
                labels[:n16 // 2] = labels[n16 // 2:n16 // 2 * 2]
        i1, v1, i2, v2 = np.zeros(R, np.int64), np.zeros(R, np.int32), np.zeros(R, np.int64), np.zeros(R, np.int32)
        O.init_heap(i1, v1, signd)
        mod.init_heap(i2, v2, signd)
        i3, v3 = i1.copy(), v1.copy()
        for _ in range(2):
            O.query_pq(packed, n, T, i1, v1, signd, labels, order)
            getattr(mod, "query_pq_" + order)(packed, n, T, i2, v2, signd, labels)
            O.replay(o1, n, i3, v3, signd, labels)
            assert np.array_equal(i1, i2) and np.array_equal(v1, v2)
            assert np.array_equal(i1, i3) and np.array_equal(v1, v3)


# ---- (c) golden fixtures ----------------------------------------------------------------------------

def test_golden_scan(golden):
    z = golden["scan"]
    for ci, (is_avx, signd, M, n, R, with_labels) in enumerate(z["cases"]):
        p, order = "c%d_" % ci, "avx" if is_avx else "sse"
        packed = O.transform_data(z[p + "codes"])
        assert np.array_equal(packed, z[p + "packed"])
        T = O.transform_tables(z[p + "tab"])
        assert np.array_equal(T, z[p + "tables"])
        est = np.zeros(2 * len(packed), np.uint64)
        O.estimate_pq(packed, T, est, bool(signd), order)
        assert np.array_equal(est, z[p + "est"])
        labels = z[p + "labels"] if with_labels else None
        hi, hv = np.zeros(R, np.int64), np.zeros(R, np.int32)
        O.init_heap(hi, hv, bool(signd))
        for rep in range(2):
            O.query_pq(packed, int(n), T, hi, hv, bool(signd), labels, order)
            assert np.array_equal(hi, z[p + "heap_idx"][rep]) and np.array_equal(hv, z[p + "heap_val"][rep])


def test_golden_lut(golden):
    z = golden["lut"]
    for name in z["names"]:
        R = z[name + "_R"]
        pq = O.PQState(int(z[name + "_dpb"]), z[name + "_centers"], None if R.size == 0 else R, z[name + "_sqrt"][()])
        for i, q in enumerate(z[name + "_q"]):
            T, qrot, shift, scale = O.distance_table(pq, q)
            assert np.array_equal(T, z[name + "_tables"][i])
            assert float(shift) == z[name + "_shift"][i] and float(scale) == z[name + "_scale"][i]
            assert np.array_equal(O.udistance_table(pq, q)[0], z[name + "_utables"][i])


@pytest.mark.parametrize("kind", ["port", "ref"])
def test_golden_ivf(golden, kind):
    if kind == "ref" and not ref_loader.have_ref_kernels():
        pytest.skip("oracle/_ref not built")
    z = golden["ivf"]
    K = O.Kernels(kind, "avx")
    bad = total = 0
    for name in z["names"]:
        S = O.ivf_state_from_arrays(z, name + "_")
        for npr in (1, 3, 8):
            exp = z["%s_res_p%d" % (name, npr)]
            for q, e in zip(z[name + "_q"], exp):
                got = O.ivf_query(S, q, 10, n_probes=npr, kernels=K)
                total += 1
                bad += not np.array_equal(got, e[e != -1])
    # np.argpartition's order is CPU specific: on a different host than the one that
    # generated the fixtures the probe order may differ for a few queries; sets must still agree mostly.
    assert bad <= 0.02 * total, (bad, total)


# ---- (d) the reference package itself (build container only) -------------------------------------------

@pytest.mark.skipif(not ref_loader.have_ref_package(), reason="/root/reference not present")
def test_restatement_matches_reference_package():
    import warnings
    warnings.filterwarnings("ignore")
    t = ref_loader.load_ref_package()
    np.random.seed(3)
    for n, d, metric in ((1500, 128, "euclidean"), (1200, 100, "angular"), (300, 10, "euclidean")):
        X = np.random.randn(n, d).astype(np.float32) + 2 * np.random.randn(1, d).astype(np.float32)
        qs = np.random.randn(40, d).astype(np.float32)
        ivf = t.IVF(metric, int(n ** 0.5), t.FastPQ(2))
        ivf.fit(X).build(X, n_probes=2)
        S = O.IVFState.from_ivf(ivf)
        assert np.array_equal(O.transform_data(t._transform.unpack(ivf.pq_transformed_centers.packed)),
                              ivf.pq_transformed_centers.packed)
        for kind in ("port", "ref"):
            K = O.Kernels(kind, "avx")
            for q in qs:
                qq = q / np.linalg.norm(q) if metric == "angular" else q
                dt = ivf.pq.distance_table(qq)
                T, _, sh, sc = O.distance_table(S.pq, qq)
                assert np.array_equal(T, dt.tables) and sh == dt.mean and sc == dt.scale
                assert np.array_equal(O.udistance_table(S.pq, qq)[0], ivf.pq.udistance_table(qq).tables)
                for npr in (1, 4):
                    assert np.array_equal(ivf.query(q.copy(), 10, n_probes=npr),
                                          O.ivf_query(S, q, 10, n_probes=npr, kernels=K))


def _enc_pq(z, name):
    R = z[name + "_R"]
    return O.PQState(2, z[name + "_centers"], None if R.size == 0 else R)


def test_golden_encode(golden):
    """oracle.pq_transform == the reference's FastPQ.transform output (tests/golden/encode.npz), bit for bit."""
    z = golden["encode"]
    for name in z["names"]:
        n, packed = O.pq_transform(_enc_pq(z, name), z[name + "_X"])
        assert n == len(z[name + "_X"]) and np.array_equal(packed, z[name + "_packed"]), name


@pytest.mark.skipif(not ref_loader.have_ref_package(), reason="/root/reference not present")
def test_encode_restatement_matches_reference_package():
    t = ref_loader.load_ref_package()
    rng = np.random.default_rng(5)
    for d in (24, 100, 128):
        X = rng.standard_normal((300, d)).astype(np.float32)
        pq = t.FastPQ(2, use_kmeans=True).fit(X)
        td = pq.transform(X[:77])
        n, packed = O.pq_transform(O.PQState.from_pq(pq), X[:77])
        assert n == td.size and np.array_equal(packed, td.packed)


# ---- the arithmetic the build-time CUDA kernels implement, stated in scalar C and pinned against numpy itself -------

def _chain_lib():
    import ctypes
    from oracle import build_oracle
    L = ctypes.CDLL(build_oracle.build_chain())
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.tko_dchain.argtypes = L.tko_schain.argtypes = [vp, vp, vp, i64, i32, i32]
    L.tko_encode_f64.argtypes = L.tko_encode_f32.argtypes = [vp, i64, i32, i32, vp, vp, vp]
    L.tko_assign_f32.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp]
    return L


def _p(a):
    return a.ctypes.data


def test_numpy_matmul_is_a_sequential_fma_chain():
    """The assumption tkb_encode.cu / tkb_assign.cu rest on: numpy's `@` (OpenBLAS gemm) computes every output as an FMA
    chain over the inner dimension in ascending order starting from 0 (one K block: K <= 384 for f32)."""
    L = _chain_lib()
    rng = np.random.default_rng(0)
    for dt, fn, shapes in ((np.float64, L.tko_dchain, [(700, 64, 128), (100, 16, 2), (333, 24, 24)]),
                           (np.float32, L.tko_schain, [(100, 16, 2), (100, 1087, 100), (100, 300, 128), (100, 64, 384)])):
        for n, m, K in shapes:
            a, b = rng.standard_normal((n, K)).astype(dt), rng.standard_normal((m, K)).astype(dt)
            out = np.empty((n, m), dt)
            fn(_p(a), _p(b), _p(out), n, m, K)
            assert np.array_equal(out, a @ b.T), (dt.__name__, n, m, K)


def test_numpy_f32_dot_is_openblas_skylakex_sdot():
    """The assumption the LUT kernel's angular normalisation rests on (tkb_lut.cu; ref: ivf.py:127 `q /= np.linalg.norm(q)`):
    `np.linalg.norm` of an f32 vector is sqrt(x.dot(x)) and the dot product is OpenBLAS's sdot, whose SkylakeX kernels sum in 64 /
    32 lanes with a double-precision tail -- oracle.restate.sdot_openblas_skylakex. Pinned against numpy itself wherever numpy
    runs those kernels (the hosts of this pool); skipped on a host whose OpenBLAS picked another core."""
    from threadpoolctl import threadpool_info
    arch = [d.get("architecture") for d in threadpool_info() if d.get("internal_api") == "openblas"]
    if "SkylakeX" not in arch:
        pytest.skip("numpy's OpenBLAS runs %s kernels here, not SkylakeX" % arch)
    rng = np.random.default_rng(8)
    for n in (1, 7, 31, 32, 33, 64, 96, 99, 100, 104, 128, 200, 1000):
        for _ in range(60):
            x = (rng.standard_normal(n) * rng.choice([1e-3, 1.0, 50.0])).astype(np.float32)
            assert O.sdot_openblas_skylakex(x, x).tobytes() == x.dot(x).tobytes(), n
            assert np.linalg.norm(x).tobytes() == np.sqrt(O.sdot_openblas_skylakex(x, x)).tobytes(), n


def test_numpy_vector_times_matrix_is_a_four_lane_fma_chain(golden):
    """The assumption the LUT kernel's rotation rests on (tkb_lut.cu; ref: fast_pq.py:203-204 `q @ R.T`): for a float64 VECTOR
    numpy calls OpenBLAS's dgemv, which sums in four lanes (k mod 4), FMA, combined (a0 + a2) + (a1 + a3) -- pinned against
    numpy itself for random inputs and against the reference's own rotated queries in the golden LUT fixture."""
    import ctypes
    L = _chain_lib()
    L.tko_dgemv4.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    rng = np.random.default_rng(3)
    for K, m in ((16, 16), (24, 24), (64, 64), (104, 64), (128, 64), (200, 64), (256, 64), (1024, 64)):
        R = np.ascontiguousarray(rng.standard_normal((m, K)))
        for _ in range(20):
            q = rng.standard_normal(K).astype(np.float32)
            out, q64 = np.empty(m, np.float64), q.astype(np.float64)
            L.tko_dgemv4(_p(q64), _p(R), _p(out), m, K)
            assert np.array_equal(out, q @ R.T), (K, m)
    z = golden["lut"]
    for name in z["names"]:
        R = z[name + "_R"]
        if R.size == 0 or R.shape[1] % 4:
            continue
        R = np.ascontiguousarray(R)
        for q, want in zip(z[name + "_q"], z[name + "_qrot"]):
            qp = np.zeros(R.shape[1], np.float64)
            qp[:len(q)] = q
            out = np.empty(R.shape[0], np.float64)
            L.tko_dgemv4(_p(qp), _p(R), _p(out), R.shape[0], R.shape[1])
            assert np.array_equal(out, want), name


def test_scalar_encoder_arithmetic_equals_reference_codes(golden):
    """tko_encode_* (the arithmetic of tkb_encode.cu, one element at a time) reproduces the reference's packed codes."""
    L = _chain_lib()
    z = golden["encode"]
    for name in z["names"]:
        X, centers, R = z[name + "_X"], np.ascontiguousarray(z[name + "_centers"]), z[name + "_R"]
        data = O.pad2(X, 16, 8)
        if R.size:
            a = np.ascontiguousarray(data, dtype=np.float64)
            rot = np.empty((len(a), R.shape[0]), np.float64)
            L.tko_dchain(_p(a), _p(np.ascontiguousarray(R)), _p(rot), len(a), R.shape[0], R.shape[1])
            assert np.array_equal(rot, data @ R.T)
            data = rot
        Dp = centers.shape[1]
        M = Dp // 2
        cn = np.ascontiguousarray(np.stack([np.einsum("ij,ij->i", b, b) for b in centers.reshape(16, M, 2).transpose(1, 0, 2)]), dtype=np.float32)
        codes = np.empty((len(data), M), np.uint8)
        data = np.ascontiguousarray(data)
        (L.tko_encode_f64 if data.dtype == np.float64 else L.tko_encode_f32)(_p(data), len(data), Dp, 2, _p(centers), _p(cn), _p(codes))
        assert np.array_equal(O.transform_data(codes), z[name + "_packed"]), name


def test_scalar_assignment_arithmetic_equals_knn_brute():
    """tko_assign_f32 (the arithmetic of tkb_assign.cu) == the reference's knn_brute(X, Y, 1) given numpy's own norms."""
    L = _chain_lib()
    rng = np.random.default_rng(3)
    for n, d, C in ((400, 100, 1087), (300, 128, 200), (200, 20, 37)):
        means = rng.standard_normal((C, d)) * 2
        X = (means[rng.integers(C, size=n)] + rng.standard_normal((n, d))).astype(np.float32)
        Y = np.ascontiguousarray(means + 0.1 * rng.standard_normal((C, d)), dtype=np.float32)
        xn, yn = np.einsum("ij,ij->i", X, X), np.einsum("ij,ij->i", Y, Y)
        out = np.empty(n, np.int32)
        L.tko_assign_f32(_p(X), n, d, _p(Y), C, _p(xn), _p(yn), _p(out))
        assert np.array_equal(out, O.knn_brute(X, Y, 1)[:, 0])
