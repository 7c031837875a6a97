"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bar: bit-exact for estimates, heap arrays, ids; exact-distance tolerance 1e-5 relative (north star).
"""
from itertools import product

import numpy as np
import pytest

import tinyknn_b200 as tinyknn
from tinyknn_b200 import _lib, _device as D
from tinyknn_b200._fast_pq import estimate_pq_sse, query_pq_sse, init_heap
from tinyknn_b200._fast_pq_avx import estimate_pq_avx, query_pq_avx
from tinyknn_b200._transform import transform_data, transform_tables
from oracle import restate as O

pytestmark = pytest.mark.gpu

EST = {"sse": estimate_pq_sse, "avx": estimate_pq_avx}
QRY = {"sse": query_pq_sse, "avx": query_pq_avx}


def _rand_case(rng, order, signd, M, n, kind):
    n16 = -(-n // 16) * 16
    codes = rng.integers(0, 16, size=(n16, M), dtype=np.uint8)
    if kind == "full":
        tab = rng.integers(0, 256, size=(M, 16)).astype(np.uint8)
    elif kind == "lut":
        tab = np.round(rng.exponential(6.0, size=(M, 16)) - 4).clip(-4, 128 / M ** 0.5).astype(np.int8).view(np.uint8)
        if not signd:
            tab = (tab.view(np.int8) + 4).astype(np.uint8)
    else:
        tab = rng.integers(0, 28, size=(M, 16)).astype(np.int16)
        tab = (tab - 4).astype(np.int8).view(np.uint8) if signd else tab.astype(np.uint8)
    return codes, tab, transform_data(codes), transform_tables(tab)


# ---- estimates ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n,d,signed,simd", list(product([16, 32], [4, 8], [True, False], ["sse", "avx"])))
def test_estimate_pq_simd(n, d, signed, simd):            # ref: tests/test_pq.py:12-53
    rng = np.random.default_rng(10)
    data = rng.integers(0, 16, size=(n, d), dtype=np.uint8)
    tables = rng.integers(0, 256, size=(d, 16)).astype(np.uint8)
    out = np.zeros(n // 8, dtype=np.uint64)
    EST[simd](transform_data(data), transform_tables(tables), out, signed)
    exp = np.zeros(n // 8, dtype=np.uint64)
    O.estimate_pq(O.transform_data(data), O.transform_tables(tables), exp, signed, simd)
    assert np.array_equal(out, exp)


def test_estimate_random_shapes_bit_exact():
    rng = np.random.default_rng(1)
    for trial in range(60):
        order = ("sse", "avx")[trial % 2]
        signd = bool((trial // 2) % 2)
        M = int(rng.integers(1, 17)) * (4 if order == "avx" else 2)
        n = int(rng.integers(1, 5000))
        _, _, packed, T = _rand_case(rng, order, signd, M, n, ("full", "narrow", "lut")[trial % 3])
        out, exp = np.zeros(2 * len(packed), np.uint64), np.zeros(2 * len(packed), np.uint64)
        EST[order](packed, T, out, signd)
        O.estimate_pq(packed, T, exp, signd, order)
        assert np.array_equal(out, exp), (trial, order, signd, M, n)


def test_estimate_large_bit_exact():
    """2M vectors x 16 B (32 MB of codes): every estimate equal to the oracle, both orders."""
    rng = np.random.default_rng(2)
    n = 2_000_000
    for order, signd in (("avx", True), ("sse", True), ("avx", False)):
        _, _, packed, T = _rand_case(rng, order, signd, 32, n, "lut")
        out, exp = np.zeros(2 * len(packed), np.uint64), np.zeros(2 * len(packed), np.uint64)
        EST[order](packed, T, out, signd)
        O.estimate_pq(packed, T, exp, signd, order)
        assert np.array_equal(out, exp)


def test_estimate_rejects_bad_M():
    with pytest.raises(ValueError):
        estimate_pq_avx(np.zeros((1, 6), np.uint64), np.zeros(12, np.uint64), np.zeros(2, np.uint64), True)
    estimate_pq_avx(np.zeros((0, 8), np.uint64), np.zeros(16, np.uint64), np.zeros(0, np.uint64), True)   # empty ok


def test_golden_scan_and_heaps(golden):
    z = golden["scan"]
    for ci, (is_avx, signd, M, n, R, with_labels) in enumerate(z["cases"]):
        p, order = "c%d_" % ci, "avx" if is_avx else "sse"
        packed, T = z[p + "packed"], z[p + "tables"]
        est = np.zeros(2 * len(packed), np.uint64)
        EST[order](packed, T, est, bool(signd))
        assert np.array_equal(est, z[p + "est"]), ci
        labels = np.ascontiguousarray(z[p + "labels"]) if with_labels else None
        hi, hv = np.zeros(R, np.int64), np.zeros(R, np.int32)
        init_heap(hi, hv, bool(signd))
        for rep in range(2):
            QRY[order](packed, int(n), T, hi, hv, bool(signd), labels)
            assert np.array_equal(hi, z[p + "heap_idx"][rep]) and np.array_equal(hv, z[p + "heap_val"][rep]), (ci, rep)


# ---- heap replay ------------------------------------------------------------------------------------------

def test_query_pq_heap_arrays_match_oracle():
    rng = np.random.default_rng(3)
    for trial in range(80):
        order = ("sse", "avx")[trial % 2]
        signd = bool((trial // 2) % 2)
        M = int(rng.integers(1, 9)) * 4
        n = int(rng.integers(1, 3000))
        _, _, packed, T = _rand_case(rng, order, signd, M, n, ("lut", "narrow", "full")[trial % 3])
        n16 = 16 * len(packed)
        R = int(rng.integers(1, 150))
        labels = None
        if trial % 4:
            labels = rng.permutation(10 ** 6)[:n16].astype(np.int64) + 10 ** 12
            if trial % 3 == 0:
                labels[:n16 // 2] = labels[n16 // 2:n16 // 2 * 2]
        a = (np.zeros(R, np.int64), np.zeros(R, np.int32))
        b = (np.zeros(R, np.int64), np.zeros(R, np.int32))
        init_heap(*a, signd)
        O.init_heap(*b, signd)
        for rep in range(3):                               # state persists across calls (ref: ivf.py:137-150)
            QRY[order](packed, n, T, a[0], a[1], signd, labels)
            O.query_pq(packed, n, T, b[0], b[1], signd, labels, order)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (trial, rep)


@pytest.mark.parametrize("n,dpb,signed", list(product(tuple(range(1, 10)) + (20, 30, 50), [1, 2], [True, False])))
def test_topk(n, dpb, signed):                             # ref: tests/test_pq.py:86-90, 111-140
    np.random.seed(n * 7 + dpb)
    X = np.random.randn(n, 11).astype(np.float32)
    qs = np.random.randn(3, 11).astype(np.float32)
    pq = tinyknn.FastPQ(dims_per_block=dpb)
    _, data = pq.fit_transform(X)
    for q in qs:
        dtable = pq.distance_table(q)
        out = np.zeros(2 * len(data), dtype=np.uint64)
        estimate_pq_sse(data, dtable.tables, out, signed)
        est = out.view(np.int8 if signed else np.uint8)[:n]
        indices, values = np.zeros(n, np.int64), np.zeros(n, np.int32)
        init_heap(indices, values, signed)
        query_pq_sse(data, n, dtable.tables, indices, values, signed)
        maxv = 127 if signed else 255
        values = np.sort(values[values < maxv])
        est = np.sort(est)
        assert np.all(est[est < maxv] == values)


def test_topk_0():                                         # ref: tests/test_pq.py:94-97
    with pytest.raises(AssertionError):
        tinyknn.FastPQ(2).fit_transform(np.random.randn(0, 11).astype(np.float32))


def test_large_labels():                                   # ref: tests/test_pq.py:143-158
    np.random.seed(10)
    n, d, k = 100, 10, 100
    X = np.random.randn(n, d).astype(np.float32)
    q = np.random.randn(d).astype(np.float32)
    pq = tinyknn.FastPQ(2)
    _, data = pq.fit_transform(X)
    dtable = pq.distance_table(q)
    indices, values = np.empty(k, np.int64), np.empty(k, np.int32)
    labels = np.arange(n, dtype=np.int64) + 10 ** 12
    init_heap(indices, values, True)
    query_pq_sse(data, n, dtable.tables, indices, values, True, labels)
    indices.sort()
    np.testing.assert_array_equal(indices, labels)


# ---- LUT build ----------------------------------------------------------------------------------------------

def _pq_from_golden(z, name):
    pq = tinyknn.FastPQ(int(z[name + "_dpb"]))
    R = z[name + "_R"]
    pq.centers, pq.R, pq.sqrt_n_blocks = z[name + "_centers"], (None if R.size == 0 else R), z[name + "_sqrt"][()]
    return pq


def test_lut_bytes_match_reference(golden):
    z = golden["lut"]
    for name in z["names"]:
        pq = _pq_from_golden(z, name)
        qs = z[name + "_q"]
        lut = pq.distance_tables(qs, signed=True)
        got = lut["tables"].cpu().numpy().reshape(len(qs), -1).view(np.uint64)
        bad_q = int((got != z[name + "_tables"]).any(axis=1).sum())
        ulut = pq.distance_tables(qs, signed=False)["tables"].cpu().numpy().reshape(len(qs), -1).view(np.uint64)
        bad_u = int((ulut != z[name + "_utables"]).any(axis=1).sum())
        print(f"LUT parity {name}: signed {len(qs) - bad_q}/{len(qs)} queries identical, unsigned {len(qs) - bad_u}/{len(qs)}")
        assert bad_q == 0 and bad_u == 0, (name, bad_q, bad_u)
        np.testing.assert_allclose(lut["shift"].cpu().numpy(), z[name + "_shift"], rtol=1e-12 if pq.R is not None else 1e-6)
        np.testing.assert_allclose(lut["scale"].cpu().numpy(), z[name + "_scale"], rtol=1e-12 if pq.R is not None else 1e-6)
        # the rotation mirrors numpy's dgemv operation for operation (4 accumulators + FMA): q_rot is the reference's, bit for bit;
        # shift and scale follow from it through numpy-ordered sums (pairwise mean) and are then identical too
        if pq.R is not None and pq.R.shape[1] % 4 == 0:
            assert np.array_equal(lut["q_rot"].cpu().numpy(), np.asarray(z[name + "_qrot"], dtype=np.float64)), name
            if name == "d128":              # other block sizes: einsum hands numpy's mean a non-contiguous array, summed in another order
                assert np.array_equal(lut["shift"].cpu().numpy(), np.asarray(z[name + "_shift"], dtype=np.float64)), name
                assert np.array_equal(lut["scale"].cpu().numpy(), np.asarray(z[name + "_scale"], dtype=np.float64)), name
        else:
            np.testing.assert_allclose(lut["q_rot"].cpu().numpy(), z[name + "_qrot"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("d", [100, 7, 31, 64, 96, 128, 200])
def test_angular_normalisation_equals_numpy_bit_for_bit(d):
    """`q /= np.linalg.norm(q)` (ref: ivf.py:126-127) in the LUT kernel: the normalised queries equal numpy's on a host whose
    OpenBLAS runs the SkylakeX sdot kernels (this pool; tests/test_oracle_pinned.py pins that arithmetic), for vector lengths on
    both sides of the kernel's 32- and 64-element blocks."""
    from threadpoolctl import threadpool_info
    if "SkylakeX" not in [t.get("architecture") for t in threadpool_info() if t.get("internal_api") == "openblas"]:
        pytest.skip("numpy's OpenBLAS does not run the SkylakeX kernels on this host")
    rng = np.random.default_rng(d)
    pq = tinyknn.FastPQ(2, rotate_dim=None)                       # codebooks do not matter here: any fitted state will do
    Dpad = -(-d // 8) * 8
    pq.centers = rng.standard_normal((16, Dpad)).astype(np.float32)
    pq.sqrt_n_blocks = np.sqrt(Dpad // 2)
    qs = (rng.standard_normal((500, d)) * rng.choice([1e-2, 1.0, 30.0], size=(500, 1))).astype(np.float32)
    lut = pq.distance_tables(D.upload(qs), signed=True, normalize=True)
    got = lut["q"].cpu().numpy()
    exp = np.stack([q / np.linalg.norm(q) for q in qs])
    assert got.tobytes() == exp.tobytes()


def test_distance_table_object(golden):
    z = golden["lut"]
    for name in ("d128", "d100"):
        pq = _pq_from_golden(z, name)
        q = z[name + "_q"][0]
        dt = pq.distance_table(q)
        assert dt.signed and dt.tables.dtype == np.uint64 and np.array_equal(dt.tables, z[name + "_tables"][0])
        assert dt.raw_q is q or np.array_equal(dt.raw_q, q)
        assert type(dt.mean) is (np.float64 if pq.R is not None else np.float32)
        assert float(dt.scale) == pytest.approx(z[name + "_scale"][0], rel=1e-12 if pq.R is not None else 1e-6)
        ud = pq.udistance_table(q)
        assert not ud.signed and np.array_equal(ud.tables, z[name + "_utables"][0])


# ---- FastPQ brute-force API ------------------------------------------------------------------------------------

def test_estimate_distances_and_top_match_oracle():
    np.random.seed(4)
    for n, d in ((16000, 128), (777, 100), (50, 10)):
        X = np.random.randn(n, d).astype(np.float32)
        qs = np.random.randn(8, d).astype(np.float32)
        pq = tinyknn.FastPQ(2)
        td = pq.fit_transform(X[:4000]) if n > 4000 else pq.fit_transform(X)
        if n > 4000:
            td = pq.transform(X)
        S = O.PQState.from_pq(pq)
        K = O.Kernels("port", "avx")
        for q in qs:
            dt = pq.distance_table(q)
            od = O.make_dtable(S, q, K)
            assert np.array_equal(dt.tables, od.tables)
            e, oe = dt.estimate_distances(td), od.estimate_distances(td)
            assert e.dtype == np.int8 and e.shape == (n,) and np.array_equal(e, oe)
            np.testing.assert_allclose(dt.estimate_distances(td, rescale=True), od.estimate_distances(td, rescale=True), rtol=1e-6)
            hi, _ = od.top_heap(td, min(30, n))
            ghi, _ = dt._heap_dev(td, min(30, n))
            assert np.array_equal(ghi.cpu().numpy(), hi)
            assert set(dt.top(td, X, 10)) == set(od.top(td, X, 10))
            ud, oud = pq.udistance_table(q), O.make_dtable(S, q, K, signed=False)
            assert np.array_equal(ud.estimate_distances(td), oud.estimate_distances(td))
        out = np.zeros(2 * len(td.packed), dtype=np.uint64)
        r = pq.distance_table(qs[0]).estimate_distances(td, out=out)      # caller-provided buffer is reused
        assert np.shares_memory(r, out)


@pytest.mark.parametrize("i,method,signed,use_kmeans",
                         list(product(range(1, 5), ["argpartition", "top"], [True, False], [True, False])))
def test_recall(i, method, signed, use_kmeans):            # ref: tests/test_pq.py:56-82
    np.random.seed(10 + i)
    n = np.random.randint(16 * i, 16 * (i + 1))
    d, k = 8 * i, 100
    X = np.random.randn(n, d).astype(np.float32)
    qs = np.random.randn(k, d).astype(np.float32)
    trus = tinyknn.knn_brute(qs, X, k=1)[:, 0]
    pq = tinyknn.FastPQ(dims_per_block=2, use_kmeans=use_kmeans)
    data = pq.fit_transform(X)
    hits = 0
    for q, tru in zip(qs, trus):
        dtable = pq.distance_table(q) if signed else pq.udistance_table(q)
        if method == "argpartition":
            top10 = dtable.estimate_distances(data).argpartition(10)[:10]
        else:
            top10 = dtable.top(data, X, 10)
        hits += tru in top10
    assert hits / k > 0.8


# ---- IVF ----------------------------------------------------------------------------------------------------------

def _ivf_from_state(S):
    """A tinyknn_b200.IVF carrying the arrays of an oracle IVFState (as if unpickled)."""
    ivf = tinyknn.IVF(S.metric, len(S.pq_transformed_points), tinyknn.FastPQ(S.pq.dims_per_block))
    ivf.pq.centers, ivf.pq.R, ivf.pq.sqrt_n_blocks = S.pq.centers, S.pq.R, S.pq.sqrt_n_blocks
    ivf.pq_transformed_centers = tinyknn.fast_pq.TransformedData(*S.pq_transformed_centers)
    ivf.active_centers = S.active_centers
    ivf.pq_transformed_points = [tinyknn.fast_pq.TransformedData(*t) for t in S.pq_transformed_points]
    ivf.ids, ivf.data = S.ids, S.data
    return ivf


def test_ivf_query_matches_golden_and_oracle(golden):
    z = golden["ivf"]
    K = O.Kernels("port", "avx")
    for name in z["names"]:
        S = O.ivf_state_from_arrays(z, name + "_")
        ivf = _ivf_from_state(S)
        qs = z[name + "_q"]
        for npr in (1, 3, 8):
            gold = z["%s_res_p%d" % (name, npr)]
            ids, cnt = ivf.query_batch(qs, 10, n_probes=npr, order="numpy")
            last = {k: v.cpu().numpy() for k, v in ivf._last.items()}
            bad_gold = bad_or = bad_heap = bad_lut = 0
            for i, q in enumerate(qs):
                tr = {}
                exp = O.ivf_query(S, q, 10, n_probes=npr, kernels=K, trace=tr)
                got = ids[i][:cnt[i]]
                bad_or += set(got) != set(exp)
                bad_gold += set(got) != set(gold[i][gold[i] != -1])
                bad_lut += not np.array_equal(last["tables"][i].reshape(-1).view(np.uint64), tr["tables"])
                bad_heap += not (np.array_equal(last["heap_idx"][i], tr["heap_indices"])
                                 and np.array_equal(last["heap_val"][i], tr["heap_values"])
                                 and np.array_equal(last["probes"][i], tr["top"]))
            print(f"IVF {name} n_probes={npr}: lut!= {bad_lut}, heap!= {bad_heap}, ids!=oracle {bad_or}, ids!=golden {bad_gold} of {len(qs)}")
            assert bad_lut == 0 and bad_heap == 0 and bad_or == 0
            # What pins the ids is the line above: equality with the oracle -- the reference's algorithm driven by THIS host's
            # numpy -- on every query. The golden id arrays were produced by the reference on the machine that generated the
            # fixtures; np.argpartition's output order (which decides the visiting order of the probed lists and with it the
            # heap, SURVEY.md 0.5 / DESIGN.md 4.5) is specific to the numpy build and CPU, so on another machine the reference
            # itself may differ from them on a few queries. They are a drift alarm, not the acceptance test.
            assert bad_gold <= max(1, len(qs) // 16)
            # single-query API == batch of one
            one = ivf.query(qs[0], 10, n_probes=npr)
            assert one.dtype == np.int64 and set(one) == set(ids[0][:cnt[0]])


def test_ivf_device_order_matches_sorted_oracle(golden):
    """Throughput mode: selections stay on the GPU with the rule 'ascending distance, then slot'.
    The oracle is run with the same rule (restate.bottom_k_sorted) -> ids must be set-identical, and
    the distance of every returned id must match the oracle formula within 1e-5 relative."""
    z = golden["ivf"]
    K = O.Kernels("port", "avx")
    for name in z["names"]:
        S = O.ivf_state_from_arrays(z, name + "_")
        ivf = _ivf_from_state(S)
        qs = z[name + "_q"]
        for npr in (1, 4, 8):
            ids, cnt, dst = ivf.query_batch(qs, 10, n_probes=npr, order="device", return_distances=True)
            bad = 0
            for i, q in enumerate(qs):
                exp = O.ivf_query(S, q, 10, n_probes=npr, kernels=K, select=O.bottom_k_sorted)
                got = ids[i][:cnt[i]]
                bad += set(got) != set(exp)
                qq = q / np.linalg.norm(q) if S.metric == "angular" else q
                ref_d = O.exact_dists(qq.astype(np.float32), S.data[got])
                np.testing.assert_allclose(dst[i][:cnt[i]], ref_d, rtol=1e-5)
                assert np.all(np.diff(dst[i][:cnt[i]]) >= 0)
            print(f"IVF device-order {name} n_probes={npr}: ids != sorted-oracle in {bad}/{len(qs)}")
            assert bad == 0


def test_ivf_small_n():                                    # ref: tests/test_ivf.py:7-31
    np.random.seed(0)
    d = 10
    for metric in ["euclidean", "angular"]:
        for n in range(1, 5):
            for far in (False, True):
                X = np.random.randn(n, d).astype(np.float32)
                if far and n > 1:
                    X[0, :] = 10 ** 5
                q = np.random.randn(d).astype(np.float32)
                ivf = tinyknn.IVF(metric, 1, tinyknn.FastPQ(2))
                ivf.fit(X).build(X, n_probes=1)
                res = ivf.query(q, n)
                assert all(0 <= i < n for i in res)
                exp = O.ivf_query(O.IVFState.from_ivf(ivf), q, n)
                assert set(res) == set(exp)


def _recall(n, d, nq, at, metric, n_probes, build_probes=2):
    X = np.random.randn(n, d).astype(np.float32)
    qs = np.random.randn(nq, d).astype(np.float32)
    trus = tinyknn.knn_brute(qs, X, k=at, metric=metric) if at < n else np.broadcast_to(np.arange(n), (nq, n))
    ivf = tinyknn.IVF(metric, int(n ** 0.5), tinyknn.FastPQ(2))
    ivf.fit(X).build(X, n_probes=build_probes)
    hits = 0
    for q, tru in zip(qs, trus):
        hits += len(set(ivf.query(q, k=at, n_probes=n_probes)) & set(tru))
    return hits / nq / at


def test_ivf_recall_floors():                              # ref: tests/test_ivf.py:34-47,67-69 ; test_multiprobe.py:54-67
    np.random.seed(10)
    for metric, floors in (("euclidean", (0.1, 0.2, 0.35, 0.5)), ("angular", (0.09, 0.18, 0.27, 0.36))):
        for npr, fl in zip((1, 2, 4, 8), floors):
            assert _recall(100, 20, 10, 10, metric, npr) > fl
    assert _recall(15, 10, 30, 10, "euclidean", 1) > 0.05
    for metric in ("angular", "euclidean"):
        assert _recall(1000, 10, 30, 10, metric, 10, build_probes=4) >= 0.9
        assert _recall(1000, 10, 30, 10, metric, 4, build_probes=10) >= 0.9


# ---- fast scan (device-native layout) vs generic scan vs oracle --------------------------------------------

def _dev_estimates(packed, tables_u8, order, signd, impl):
    """est (Q, 16*n_chunks) uint8 through the device entry points."""
    from tinyknn_b200._lib import lib, check, ORDER_AVX, ORDER_SSE
    n_chunks, M = packed.shape
    Q = tables_u8.shape[0]
    tdev = D.upload(tables_u8)
    est = D.empty((Q, 16 * n_chunks), np.uint8)
    o = ORDER_AVX if order == "avx" else ORDER_SSE
    if impl == "fast":
        nat = D.to_native(D.upload(packed), n_chunks, M)
        back = D.from_native(nat, n_chunks, M).cpu().numpy().view(np.uint64)
        assert np.array_equal(back, packed)                      # the native layout round-trips
        ws = D.scan_workspace(Q * n_chunks)
        check(lib.tkb_estimate_native_dev(D.ptr(nat), n_chunks, M, D.ptr(tdev), Q, D.ptr(est), 16 * n_chunks, o,
                                          int(signd), D.ptr(ws), ws.numel(), D.stream_ptr()))
        npatched = int(ws[:8].cpu().numpy().view(np.uint64)[0])
    else:
        pdev = D.upload(packed)
        check(lib.tkb_estimate_dev(D.ptr(pdev), n_chunks, M, D.ptr(tdev), Q, D.ptr(est), 16 * n_chunks, o,
                                   int(signd), D.stream_ptr()))
        npatched = 0
    return est.cpu().numpy(), npatched


def test_fast_scan_bit_exact_all_table_kinds():
    """The PRMT/deferred-clamp kernel must equal the step-by-step fold for every table: tables inside the
    fast-path preconditions (certificate mostly passes), saturating tables (patch pass), and arbitrary
    full-range tables (per-query exact path)."""
    rng = np.random.default_rng(7)
    for trial in range(48):
        order = ("avx", "sse")[trial % 2]
        signd = bool((trial // 2) % 2)
        M = int(rng.integers(1, 17)) * 4
        n = int(rng.integers(1, 6000))
        kind = ("lut", "narrow", "full", "hot")[trial % 4]
        Q = 3
        tabs = []
        for _ in range(Q):
            if kind == "hot":                                    # small range but large values: many saturations
                t = rng.integers(8, 30, size=(M, 16)).astype(np.int16)
                t = (t - (20 if signd else 0)).astype(np.int8).view(np.uint8)
            else:
                t = _rand_case(rng, order, signd, M, 16, kind)[1]
            tabs.append(t)
        tabs = np.stack(tabs)
        packed = _rand_case(rng, order, signd, M, n, "lut")[2]
        fast, npatched = _dev_estimates(packed, tabs, order, signd, "fast")
        for q in range(Q):
            exp = np.zeros(2 * len(packed), np.uint64)
            O.estimate_pq(packed, O.transform_tables(tabs[q]), exp, signd, order)
            assert np.array_equal(fast[q], exp.view(np.uint8)), (trial, order, signd, M, n, kind, q, npatched)


def test_fast_scan_large_and_patch_rate():
    rng = np.random.default_rng(8)
    n, M = 1_000_000, 32
    codes = rng.integers(0, 16, size=(n, M), dtype=np.uint8)
    packed = transform_data(codes)
    tabs = np.stack([np.round(rng.exponential(6.0, size=(M, 16)) - 4).clip(-4, 23).astype(np.int8).view(np.uint8) for _ in range(2)])
    fast, npatched = _dev_estimates(packed, tabs, "avx", True, "fast")
    gen, _ = _dev_estimates(packed, tabs, "avx", True, "generic")
    assert np.array_equal(fast, gen)
    exp = np.zeros(2 * len(packed), np.uint64)
    O.estimate_pq(packed, O.transform_tables(tabs[0]), exp, True, "avx")
    assert np.array_equal(fast[0], exp.view(np.uint8))
    print(f"fast scan: {npatched} of {2 * len(packed)} chunks went through the patch pass")


def test_ivf_fast_equals_generic(golden):
    z = golden["ivf"]
    S = O.ivf_state_from_arrays(z, "euc128_")
    ivf = _ivf_from_state(S)
    qs = z["euc128_q"]
    import tinyknn_b200.fast_pq as fp
    try:
        fp.SCAN_IMPL = "generic"
        a = ivf.query_batch(qs, 10, n_probes=8, order="device", return_distances=True)
        ha = ivf._last["heap_idx"].cpu().numpy()
        fp.SCAN_IMPL = "fast"
        b = ivf.query_batch(qs, 10, n_probes=8, order="device", return_distances=True)
        hb = ivf._last["heap_idx"].cpu().numpy()
    finally:
        fp.SCAN_IMPL = "fast"
    assert np.array_equal(ha, hb)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_replay_kernels_agree_with_oracle_fresh_heap():
    """Thread-per-query replay (fresh heap, unique labels) == warp-per-query replay == oracle heap arrays."""
    from tinyknn_b200._lib import lib, check
    rng = np.random.default_rng(11)
    for trial in range(12):
        signd = bool(trial % 2)
        Q, R = int(rng.integers(1, 70)), int(rng.integers(1, 200))
        n = int(rng.integers(1, 4000))
        nck = -(-n // 16)
        est = rng.integers(0, 256, size=(Q, 16 * nck), dtype=np.uint8)
        if trial % 3 == 0:
            est = (est.astype(np.int16) // 8 + (100 if not signd else 0)).astype(np.uint8)     # many ties
        edev = D.upload(est)
        hi, hv = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)
        check(lib.tkb_replay_fresh_dev(D.ptr(edev), 16 * nck, nck, n, D.ptr(hi), D.ptr(hv), Q, R, int(signd), D.stream_ptr()))
        hi2, hv2 = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)
        check(lib.tkb_heap_fill_dev(D.ptr(hi2), D.ptr(hv2), Q * R, int(signd), D.stream_ptr()))
        check(lib.tkb_replay_dev(D.ptr(edev), 16 * nck, nck, n, D.ptr(hi2), D.ptr(hv2), Q, R, int(signd), None, D.stream_ptr()))
        a, b, a2, b2 = hi.cpu().numpy(), hv.cpu().numpy(), hi2.cpu().numpy(), hv2.cpu().numpy()
        for q in range(Q):
            oi, ov = np.zeros(R, np.int64), np.zeros(R, np.int32)
            O.init_heap(oi, ov, signd)
            O.replay(est[q], n, oi, ov, signd)
            assert np.array_equal(a[q], oi) and np.array_equal(b[q], ov), (trial, q)
            assert np.array_equal(a2[q], oi) and np.array_equal(b2[q], ov), (trial, q)


def test_replay_full_ctas_sixteen_queries_each():
    """Enough queries (>= 16 x 296) for the queue replay to pack 16 queries into every CTA, 8 per consumer warp -- the geometry
    of the benchmark batches, which the small-batch tests above never reach: heap arrays == oracle for every query, both
    replay modes (one segment per query; IVF segments through a compact plan)."""
    from tinyknn_b200._lib import lib, check, PLAN_SEND
    rng = np.random.default_rng(44)
    Q, R, n = 4800, 21, 300
    nck = -(-n // 16)
    est = rng.integers(0, 256, size=(Q, 16 * nck), dtype=np.uint8)
    est[::3] = (est[::3] // 16 + 100).astype(np.uint8)                       # every third query: few distinct values, ties
    edev = D.upload(est)
    hi, hv = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)
    check(lib.tkb_replay_fresh_dev(D.ptr(edev), 16 * nck, nck, n, D.ptr(hi), D.ptr(hv), Q, R, 1, D.stream_ptr()))
    a, b = hi.cpu().numpy(), hv.cpu().numpy()
    for q in range(Q):
        oi, ov = np.zeros(R, np.int64), np.zeros(R, np.int32)
        O.init_heap(oi, ov, True)
        O.replay(est[q], n, oi, ov, True)
        assert np.array_equal(a[q], oi) and np.array_equal(b[q], ov), q
    # IVF mode: 3 probed lists per query out of 40, compact plan
    n_lists, P = 40, 3
    sizes = rng.integers(1, 200, size=n_lists).astype(np.int32)
    nc8 = (-(-sizes.astype(np.int64) // 128)) * 8
    off = np.zeros(n_lists + 1, np.int64)
    off[1:] = np.cumsum(nc8)
    ids = rng.permutation(16 * int(off[-1])).astype(np.int64) + 10 ** 9
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
    d_off, d_sizes, d_ids, d_probes = (D.upload(x) for x in (off, sizes, ids, probes))
    d_seg, d_gb, d_ws = D.empty((Q, P), np.int64), D.empty((3,), np.int64), D.empty((Q,), np.int64)
    check(lib.tkb_ivf_plan_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), None, n_lists, PLAN_SEND, 0, 1, 0,
                               D.ptr(d_seg), D.ptr(d_gb), D.ptr(d_ws), 8 * Q, D.stream_ptr()))
    seg, total = d_seg.cpu().numpy(), int(d_gb.cpu().numpy()[1])
    pe = rng.integers(0, 256, size=max(total, 16), dtype=np.uint8)
    d_pe = D.upload(pe)
    hi, hv, fb = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32), D.empty((Q,), np.int32)
    check(lib.tkb_ivf_replay_fresh_dev(D.ptr(d_pe), 0, D.ptr(d_seg), D.ptr(d_off), D.ptr(d_sizes), n_lists, D.ptr(d_ids),
                                       D.ptr(d_probes), Q, P, D.ptr(hi), D.ptr(hv), R, 1, 1, D.ptr(fb), D.stream_ptr()))
    a, b = hi.cpu().numpy(), hv.cpu().numpy()
    for q in range(0, Q, 7):
        oi, ov = np.zeros(R, np.int64), np.zeros(R, np.int32)
        O.init_heap(oi, ov, True)
        for s in range(P):
            l = int(probes[q, s])
            ncr = -(-int(sizes[l]) // 16)
            O.replay(pe[seg[q, s]:seg[q, s] + 16 * ncr], int(sizes[l]), oi, ov, True,
                     np.ascontiguousarray(ids[16 * off[l]:16 * off[l] + 16 * ncr]))
        assert np.array_equal(a[q], oi) and np.array_equal(b[q], ov), q


@pytest.mark.parametrize("R,n", [(255, 9000), (256, 9000), (300, 20000), (1291, 60000), (4000, 30000), (70000, 90000)])
def test_replay_deep_heaps(R, n):
    """Heaps deeper than 8 levels: the pipelined queue replay runs 8 lanes per query (R <= 65535) or hands over to
    the unpipelined kernel; heap arrays == oracle slot for slot, signed and unsigned."""
    from tinyknn_b200._lib import lib, check
    rng = np.random.default_rng(R)
    nck = -(-n // 16)
    for signd in (True, False):
        Q = 5
        est = rng.integers(0, 256, size=(Q, 16 * nck), dtype=np.uint8)
        est[1] = (est[1] // 32 + 90).astype(np.uint8)                                         # few distinct values: ties everywhere
        edev = D.upload(est)
        hi, hv = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)
        check(lib.tkb_replay_fresh_dev(D.ptr(edev), 16 * nck, nck, n, D.ptr(hi), D.ptr(hv), Q, R, int(signd), D.stream_ptr()))
        a, b = hi.cpu().numpy(), hv.cpu().numpy()
        for q in range(Q):
            oi, ov = np.zeros(R, np.int64), np.zeros(R, np.int32)
            O.init_heap(oi, ov, signd)
            O.replay(est[q], n, oi, ov, signd)
            assert np.array_equal(a[q], oi) and np.array_equal(b[q], ov), (signd, q)


def test_ivf_duplicate_labels_use_dedupe_path():
    """build_probes=2 puts every point in two lists: labels repeat, the reference dedupes on insert."""
    np.random.seed(5)
    X = np.random.randn(1500, 16).astype(np.float32)
    qs = np.random.randn(24, 16).astype(np.float32)
    ivf = tinyknn.IVF("euclidean", 20, tinyknn.FastPQ(2))
    ivf.fit(X).build(X, n_probes=2)
    S = O.IVFState.from_ivf(ivf)
    assert not ivf.to_device()["unique_ids"]
    ids, cnt = ivf.query_batch(qs, 10, n_probes=6, order="numpy")
    heaps = ivf._last["heap_idx"].cpu().numpy()
    for i, q in enumerate(qs):
        tr = {}
        exp = O.ivf_query(S, q, 10, n_probes=6, trace=tr)
        assert np.array_equal(heaps[i], tr["heap_indices"]) and set(ids[i][:cnt[i]]) == set(exp)


def test_ivf_replay_fresh_random_segments():
    """IVF-mode fresh replay over random est bytes, random (also empty / tiny) lists, skipped probe slots and
    heaps small enough that the queue-replay kernel has to cut its rounds: heap arrays == oracle, per query."""
    from tinyknn_b200._lib import lib, check, PROBE_SKIP, PLAN_SEND
    from tinyknn_b200 import sharded as SH
    rng = np.random.default_rng(21)
    for trial in range(10):
        signd = bool(trial % 2)
        n_lists = int(rng.integers(3, 40))
        sizes = rng.integers(0, 700, size=n_lists).astype(np.int32)
        sizes[rng.integers(0, n_lists)] = 0
        sizes[rng.integers(0, n_lists)] = 1
        nc8 = (-(-sizes.astype(np.int64) // 128)) * 8                   # lists start on 8-chunk tiles
        off = np.zeros(n_lists + 1, np.int64)
        off[1:] = np.cumsum(nc8)
        ids = rng.permutation(16 * int(off[-1]) + 5).astype(np.int64)[:16 * int(off[-1])] + 10 ** 11
        Q, P, R = int(rng.integers(1, 60)), int(rng.integers(1, min(n_lists, 12) + 1)), int(rng.integers(1, 160))
        probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
        if trial % 3 == 0:
            probes[rng.integers(0, Q), P - 1] = PROBE_SKIP
        stride = 16 * int(max(nc8.max(), 1))
        est = rng.integers(0, 256, size=(Q, P, stride), dtype=np.uint8)
        if trial % 2 == 0:
            est = (est // 16 + (60 if not signd else 0)).astype(np.uint8)  # few distinct values: ties, long queues
        hi, hv, fb = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32), D.empty((Q,), np.int32)
        d_est, d_off, d_sizes, d_ids, d_probes = (D.upload(x) for x in (est, off, sizes, ids, probes))
        check(lib.tkb_ivf_replay_fresh_dev(D.ptr(d_est), stride, None, D.ptr(d_off), D.ptr(d_sizes), n_lists,
                                           D.ptr(d_ids), D.ptr(d_probes), Q, P, D.ptr(hi), D.ptr(hv), R,
                                           int(signd), 1, D.ptr(fb), D.stream_ptr()))
        a, b = hi.cpu().numpy(), hv.cpu().numpy()
        # the same segments packed by the device plan (compact buffer) give the same heaps, in both replay kernels
        seg, gb, _ = SH.plan_host(probes, sizes, None, PLAN_SEND, 0, 1, 0)
        d_seg, d_gb, d_ws = D.empty((Q, P), np.int64), D.empty((3,), np.int64), D.empty((Q,), np.int64)
        check(lib.tkb_ivf_plan_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), None, n_lists, PLAN_SEND, 0, 1, 0,
                                   D.ptr(d_seg), D.ptr(d_gb), D.ptr(d_ws), 8 * Q, D.stream_ptr()))
        assert np.array_equal(d_seg.cpu().numpy(), seg) and int(d_gb.cpu().numpy()[1]) == int(gb.sum())
        packed_est = np.zeros(max(int(gb.sum()), 16), np.uint8)
        for q in range(Q):
            for s_ in range(P):
                if seg[q, s_] >= 0:
                    nb = 16 * (-(-int(sizes[probes[q, s_]]) // 16))
                    packed_est[seg[q, s_]:seg[q, s_] + nb] = est[q, s_, :nb]
        d_pe = D.upload(packed_est)
        hi2, hv2 = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)
        check(lib.tkb_ivf_replay_fresh_dev(D.ptr(d_pe), 0, D.ptr(d_seg), D.ptr(d_off), D.ptr(d_sizes), n_lists,
                                           D.ptr(d_ids), D.ptr(d_probes), Q, P, D.ptr(hi2), D.ptr(hv2), R,
                                           int(signd), 1, D.ptr(fb), D.stream_ptr()))
        hi3, hv3 = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)
        check(lib.tkb_heap_fill_dev(D.ptr(hi3), D.ptr(hv3), Q * R, int(signd), D.stream_ptr()))
        check(lib.tkb_ivf_replay_dev(D.ptr(d_pe), 0, D.ptr(d_seg), D.ptr(d_off), D.ptr(d_sizes), n_lists,
                                     D.ptr(d_ids), D.ptr(d_probes), Q, P, D.ptr(hi3), D.ptr(hv3), R, int(signd), D.stream_ptr()))
        for x, y in ((hi2, hv2), (hi3, hv3)):
            assert np.array_equal(x.cpu().numpy(), a) and np.array_equal(y.cpu().numpy(), b), trial
        for q in range(Q):
            oi, ov = np.zeros(R, np.int64), np.zeros(R, np.int32)
            O.init_heap(oi, ov, signd)
            for s in range(P):
                l = int(probes[q, s])
                if l == PROBE_SKIP or sizes[l] == 0:
                    continue
                ncr = -(-int(sizes[l]) // 16)
                O.replay(est[q, s, :16 * ncr], int(sizes[l]), oi, ov, signd,
                         np.ascontiguousarray(ids[16 * off[l]:16 * off[l] + 16 * ncr]))
            assert np.array_equal(a[q], oi) and np.array_equal(b[q], ov), (trial, q)


def test_fast_scan_tiny_workspace_recomputes_inline():
    """A patch list that cannot hold the flagged chunks must not change results (the excess is folded exactly inline)."""
    from tinyknn_b200._lib import lib, check, ORDER_AVX
    rng = np.random.default_rng(31)
    M, n = 32, 40_000
    t = rng.integers(-4, 13, size=(2, M, 16)).astype(np.int8).view(np.uint8)     # eligible for the fast path, sums near 127
    packed = _rand_case(rng, "avx", True, M, n, "lut")[2]
    nck = len(packed)
    nat, tdev = D.to_native(D.upload(packed), nck, M), D.upload(t)
    outs = []
    for ws_bytes in (16 + 8 * 2 * nck, 16 + 8 * 3):
        est, ws = D.empty((2, 16 * nck), np.uint8), D.empty((ws_bytes,), np.uint8)
        check(lib.tkb_estimate_native_dev(D.ptr(nat), nck, M, D.ptr(tdev), 2, D.ptr(est), 16 * nck, ORDER_AVX, 1,
                                          D.ptr(ws), ws_bytes, D.stream_ptr()))
        outs.append((est.cpu().numpy(), int(ws[:8].cpu().numpy().view(np.uint64)[0])))
    assert outs[0][1] > 100 and outs[0][1] == outs[1][1]           # same chunks flagged; the second run had room for 3
    assert np.array_equal(outs[0][0], outs[1][0])
    for q in range(2):
        exp = np.zeros(2 * nck, np.uint64)
        O.estimate_pq(packed, O.transform_tables(t[q]), exp, True, "avx")
        assert np.array_equal(outs[1][0][q], exp.view(np.uint8))


@pytest.mark.parametrize("G", [1, 2, 4])
def test_plan_device_matches_host_restatement(G):
    from tinyknn_b200._lib import lib, check, PROBE_SKIP, PLAN_SEND, PLAN_RECV
    from tinyknn_b200 import sharded as SH
    rng = np.random.default_rng(40 + G)
    n_lists, Qh, P = 61, 37, 7
    sizes = rng.integers(0, 900, size=n_lists).astype(np.int32)
    owner = SH.assign_owners(sizes, G)
    Q = G * Qh
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
    probes[5 % Q, 1] = PROBE_SKIP
    probes[7 % Q, 0] = -1                                            # Python-wrapped: the last list
    d_probes, d_sizes, d_owner = D.upload(probes), D.upload(sizes), D.upload(owner)
    for r in range(G):
        for mode, rows in ((PLAN_SEND, Q), (PLAN_RECV, Qh)):
            seg, gb, base = SH.plan_host(probes, sizes, owner if G > 1 else None, mode, r, G, Qh)
            d_seg, d_gb, d_ws = D.empty((rows, P), np.int64), D.empty((2 * G + 1,), np.int64), D.empty((rows * G,), np.int64)
            check(lib.tkb_ivf_plan_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), D.ptr(d_owner) if G > 1 else None, n_lists,
                                       mode, r, G, Qh, D.ptr(d_seg), D.ptr(d_gb), D.ptr(d_ws), 8 * rows * G, D.stream_ptr()))
            g = d_gb.cpu().numpy()
            assert np.array_equal(d_seg.cpu().numpy(), seg), (G, r, mode)
            assert np.array_equal(g[:G], gb) and g[G] == gb.sum() and np.array_equal(g[G + 1:], base)


@pytest.mark.parametrize("Q", [129, 5000, 140_000])
def test_plan_device_many_queries_matches_numpy(Q):
    """The plan's scan over many count-kernel CTAs (and, above 131 072 queries, several CTAs per thread of the one scanning CTA):
    send layout, capacity guard of the pull exchange's owner plan, home-side absolute addresses -- against vectorised numpy."""
    from tinyknn_b200._lib import lib, check, PROBE_SKIP, PLAN_SEND
    rng = np.random.default_rng(Q)
    G, P, n_lists = 4, 3, 97
    Qh = -(-Q // G)
    sizes = rng.integers(0, 500, size=n_lists).astype(np.int32)
    owner = rng.integers(0, G, size=n_lists).astype(np.int32)
    probes = rng.integers(0, n_lists, size=(Q, P)).astype(np.int32)
    probes[rng.integers(0, Q, size=max(1, Q // 50)), rng.integers(0, P, size=max(1, Q // 50))] = PROBE_SKIP
    nbytes = np.where(probes == PROBE_SKIP, 0, 16 * ((sizes[np.maximum(probes, 0)].astype(np.int64) + 15) // 16))
    own = np.where(probes == PROBE_SKIP, -1, owner[np.maximum(probes, 0)])
    home = np.arange(Q) // Qh
    d_probes, d_sizes, d_owner = D.upload(probes), D.upload(sizes), D.upload(owner)
    r = 1
    # owner side (send layout of rank r): grouped by the home rank of the query, then (q, s)
    mine = (own == r) & (nbytes > 0)
    b = np.where(mine, nbytes, 0)
    group_bytes = np.array([b[home == g].sum() for g in range(G)], dtype=np.int64)
    base = np.concatenate([[0], np.cumsum(group_bytes)[:-1]])
    flat = b.reshape(-1)
    excl = np.cumsum(flat) - flat                                   # (q, s) order; homes are contiguous query ranges
    exp = np.where(mine.reshape(-1), excl, -1).reshape(Q, P)         # base[g] + offset inside the group == the global exclusive sum
    d_seg, d_gb, d_ws = D.empty((Q, P), np.int64), D.empty((2 * G + 1,), np.int64), D.empty((Q * G,), np.int64)
    check(lib.tkb_ivf_plan_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), D.ptr(d_owner), n_lists, PLAN_SEND, r, G, Qh,
                               D.ptr(d_seg), D.ptr(d_gb), D.ptr(d_ws), 8 * Q * G, D.stream_ptr()))
    g = d_gb.cpu().numpy()
    assert np.array_equal(g[:G], group_bytes) and g[G] == group_bytes.sum() and np.array_equal(g[G + 1:], base)
    assert np.array_equal(d_seg.cpu().numpy(), exp)
    cap = int(group_bytes.sum()) // 2 + 16                           # the guard: segments that would end past the capacity are dropped
    check(lib.tkb_ivf_plan_pull_owner_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), D.ptr(d_owner), n_lists, r, G, Qh, cap,
                                          D.ptr(d_seg), D.ptr(d_gb), D.ptr(d_ws), 8 * Q * G, D.stream_ptr()))
    assert np.array_equal(d_seg.cpu().numpy(), np.where((exp >= 0) & (exp + nbytes <= cap), exp, -1))
    assert d_gb.cpu().numpy()[G] == group_bytes.sum()                # the full total is still reported
    # home side of rank r: absolute addresses = owner's buffer + the owner's base of home group r + offset inside that group
    lo, hi = r * Qh, min(Q, (r + 1) * Qh)
    owner_base = (np.arange(G, dtype=np.int64) + 1) << 40
    owner_groups = np.zeros((G, G + 1), dtype=np.int64)
    expect = np.full((hi - lo, P), -1, dtype=np.int64)
    for o in range(G):
        bo = np.where((own == o) & (nbytes > 0), nbytes, 0)
        gbo = np.array([bo[home == g].sum() for g in range(G)], dtype=np.int64)
        owner_groups[o, 0] = gbo.sum()
        owner_groups[o, 1:] = np.concatenate([[0], np.cumsum(gbo)[:-1]])
        f = bo[lo:hi].reshape(-1)
        e = (np.cumsum(f) - f).reshape(hi - lo, P)
        sel = (own[lo:hi] == o) & (nbytes[lo:hi] > 0)
        expect[sel] = owner_base[o] + owner_groups[o, 1 + r] + e[sel]
    d_addr, d_gb2, d_ws2 = D.empty((hi - lo, P), np.int64), D.empty((2 * G + 1,), np.int64), D.empty((max(hi - lo, 1) * G,), np.int64)
    d_ob, d_og = D.upload(owner_base), D.upload(owner_groups)
    check(lib.tkb_ivf_plan_pull_home_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), D.ptr(d_owner), n_lists, r, G, Qh, D.ptr(d_ob),
                                         D.ptr(d_og), D.ptr(d_addr), D.ptr(d_gb2), D.ptr(d_ws2), 8 * max(hi - lo, 1) * G, D.stream_ptr()))
    assert np.array_equal(d_addr.cpu().numpy(), expect)


@pytest.mark.parametrize("G", [2, 3])
def test_sharded_phases_equal_single_gpu(golden, G):
    """List-sharded query path driven rank by rank on ONE GPU (buffers moved by hand instead of NCCL): heaps,
    ids and distances must equal the unsharded path's."""
    from tinyknn_b200.sharded import ShardedIVF
    import torch
    z = golden["ivf"]
    S = O.ivf_state_from_arrays(z, "euc128_")
    ivf = _ivf_from_state(S)
    qs = np.ascontiguousarray(z["euc128_q"][:48 // G * G])
    Qh = len(qs) // G
    k, n_probes = 10, 8
    ref_ids, ref_cnt, ref_d = ivf.query_batch(qs, k, n_probes=n_probes, order="device", return_distances=True)
    ref_heap = ivf._last["heap_idx"].cpu().numpy()
    shards = [ShardedIVF(ivf, rank=r, world=G, drop_full_codes=False) for r in range(G)]
    assert sum(int(sh.dev["n_chunks_total"]) for sh in shards) == ivf.to_device()["n_chunks_total"]
    homes = [sh._home(qs[r * Qh:(r + 1) * Qh], n_probes) for r, sh in enumerate(shards)]
    tables = torch.cat([h["lut"]["tables"] for h in homes])
    probes = torch.cat([h["probes"] for h in homes])
    scans = [sh._scan_owned(tables, probes, Qh, homes[0]["P"]) for sh in shards]
    for b, sh in enumerate(shards):                                 # all-to-all by hand: home b receives from every a
        parts = []
        for a in range(G):
            est_s, send_splits = scans[a][0], scans[a][1]
            o = int(np.sum(send_splits[:b]))
            parts.append(est_s[o:o + int(send_splits[b])])
            assert int(send_splits[b]) == int(scans[b][3][a])
        est_r = torch.cat(parts) if sum(p.numel() for p in parts) else D.empty((16,), np.uint8)
        ids, cnt, dst = sh._finish(homes[b], est_r, scans[b][2], k, (n_probes + 1) * k + 1)
        sl = slice(b * Qh, (b + 1) * Qh)
        assert np.array_equal(ivf._last["heap_idx"].cpu().numpy(), ref_heap[sl])
        assert np.array_equal(ids.cpu().numpy(), ref_ids[sl]) and np.array_equal(cnt.cpu().numpy(), ref_cnt[sl])
        assert np.array_equal(dst.cpu().numpy(), ref_d[sl])


@pytest.mark.parametrize("G", [1, 2, 4])
def test_push_plan_device_matches_host_restatement(G):
    """tkb_ivf_plan_push_dev: absolute addresses = home_base[home rank] + the host restatement's offsets."""
    from tinyknn_b200._lib import lib, check, PROBE_SKIP, PLAN_PUSH
    from tinyknn_b200 import sharded as SH
    rng = np.random.default_rng(50 + G)
    n_lists, Qh, P = 53, 29, 6
    sizes = rng.integers(0, 900, size=n_lists).astype(np.int32)
    owner = SH.assign_owners(sizes, G)
    Q = G * Qh
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
    probes[5 % Q, 1] = PROBE_SKIP
    probes[7 % Q, 0] = -1
    base = (rng.integers(1, 1 << 30, size=G).astype(np.int64) << 8)
    d_probes, d_sizes, d_owner, d_base = D.upload(probes), D.upload(sizes), D.upload(owner), D.upload(base)
    for r in range(G):
        seg, gb, _ = SH.plan_host(probes, sizes, owner if G > 1 else None, PLAN_PUSH, r, G, Qh)
        exp = np.where(seg >= 0, seg + base[np.arange(Q) // Qh][:, None], -1)
        d_seg, d_gb, d_ws = D.empty((Q, P), np.int64), D.empty((2 * G + 1,), np.int64), D.empty((Q * G,), np.int64)
        check(lib.tkb_ivf_plan_push_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), D.ptr(d_owner) if G > 1 else None, n_lists,
                                        r, G, Qh, D.ptr(d_base), D.ptr(d_seg), D.ptr(d_gb), D.ptr(d_ws), 8 * Q * G, D.stream_ptr()))
        assert np.array_equal(d_seg.cpu().numpy(), exp), (G, r)
        assert np.array_equal(d_gb.cpu().numpy()[:G], gb)


@pytest.mark.parametrize("G", [2, 3])
def test_sharded_push_phases_equal_single_gpu(golden, G):
    """The push exchange driven rank by rank on ONE GPU: every "rank" stores its estimates straight into the home
    ranks' receive buffers (tkb_peer_alloc memory, local addresses); heaps, ids, distances == the unsharded path."""
    from tinyknn_b200.sharded import ShardedIVF, PeerBuffers
    import torch
    z = golden["ivf"]
    S = O.ivf_state_from_arrays(z, "euc128_")
    ivf = _ivf_from_state(S)
    qs = np.ascontiguousarray(z["euc128_q"][:48 // G * G])
    Qh = len(qs) // G
    k, n_probes = 10, 8
    ref_ids, ref_cnt, ref_d = ivf.query_batch(qs, k, n_probes=n_probes, order="device", return_distances=True)
    ref_heap = ivf._last["heap_idx"].cpu().numpy()
    shards = [ShardedIVF(ivf, rank=r, world=G, drop_full_codes=False) for r in range(G)]
    homes = [sh._home(qs[r * Qh:(r + 1) * Qh], n_probes) for r, sh in enumerate(shards)]
    P = homes[0]["P"]
    tables = torch.cat([h["lut"]["tables"] for h in homes])
    probes = torch.cat([h["probes"] for h in homes])
    cap = shards[0].push_capacity(Qh, P)
    bufs = [PeerBuffers(cap, rank=0, world=1, n_buf=1) for _ in range(G)]       # one receive buffer per "rank", all local
    try:
        home_base = D.upload(np.array([b.local[0].address for b in bufs], dtype=np.int64))
        totals = [sh._scan_push(tables, probes, Qh, P, home_base).cpu().numpy() for sh in shards]
        for b, sh in enumerate(shards):
            seg_r, gb = ivf._plan(sh.dev, homes[b]["probes"], Qh, P)
            need = int(gb.cpu().numpy()[1])
            assert need <= cap and all(int(t[b]) == need for t in totals)
            ids, cnt, dst = sh._finish(homes[b], bufs[b].local[0], seg_r, k, (n_probes + 1) * k + 1)
            sl = slice(b * Qh, (b + 1) * Qh)
            assert np.array_equal(ivf._last["heap_idx"].cpu().numpy(), ref_heap[sl])
            assert np.array_equal(ids.cpu().numpy(), ref_ids[sl]) and np.array_equal(cnt.cpu().numpy(), ref_cnt[sl])
            assert np.array_equal(dst.cpu().numpy(), ref_d[sl])
    finally:
        for b in bufs:
            b.close()


# ---- PQ encoder (FastPQ.transform on the GPU) ------------------------------------------------------

def _enc_pq(z, name):
    pq = tinyknn.FastPQ(2)
    R = z[name + "_R"]
    pq.centers, pq.R = np.ascontiguousarray(z[name + "_centers"]), (None if R.size == 0 else np.ascontiguousarray(R))
    pq.sqrt_n_blocks = np.sqrt(pq.centers.shape[1] // 2)
    return pq


def _assert_codes_equal_up_to_ties(pq, X, got_packed, exp_packed):
    """Codes must equal the reference's; a difference is only tolerated where the two codewords are EXACTLY tied in the
    reference's own `part` matrix (np.argpartition may return either)."""
    if np.array_equal(got_packed, exp_packed):
        return 0
    got, exp = O.unpack(got_packed).astype(np.int64), O.unpack(exp_packed).astype(np.int64)
    parts = O.pq_encode_parts(O.PQState(2, pq.centers, pq.R), X)
    rows, cols = np.nonzero(got != exp)
    for r, m in zip(rows, cols):
        assert parts[m][r, got[r, m]] == parts[m][r, exp[r, m]], (r, m)
    return len(rows)


def test_encode_device_equals_reference_golden(golden):
    """tkb_encode_dev on the reference's own inputs: packed codes identical to FastPQ.transform of the reference
    (rotated f64 path, unrotated f32 path, f64 rows, padding of n to 16 and of d to dpad*dpb, 200-d rows)."""
    z = golden["encode"]
    for name in z["names"]:
        pq, X = _enc_pq(z, name), z[name + "_X"]
        td = pq.transform(X, device=True)
        assert td.size == len(X) and td.packed.dtype == np.uint64 and td.packed.shape == z[name + "_packed"].shape
        ties = _assert_codes_equal_up_to_ties(pq, X, td.packed, z[name + "_packed"])
        assert ties <= 2, (name, ties)


def test_encode_device_row_index_and_padding(golden):
    """Gathered encoding (what IVF.build uses): position i = rows[row_index[i]], out-of-range = zero vector."""
    z = golden["encode"]
    rng = np.random.default_rng(3)
    for name in ("rot128", "plain100"):
        pq, X = _enc_pq(z, name), z[name + "_X"]
        idx = rng.integers(-1, len(X), size=16 * 37).astype(np.int64)
        idx[5] = len(X) + 3                                            # out of range -> zero vector
        got = pq.encode_device(D.upload(X), row_index=D.upload(idx)).cpu().numpy().view(np.uint64)
        rows = np.where(((idx >= 0) & (idx < len(X)))[:, None], X[np.clip(idx, 0, len(X) - 1)], 0).astype(X.dtype)
        n, exp = O.pq_transform(O.PQState(2, pq.centers, pq.R), rows)
        assert _assert_codes_equal_up_to_ties(pq, rows, got, exp) <= 2


def test_encode_device_large_matches_oracle_and_scan_roundtrip():
    """200k x 128 rows: equal to the numpy restatement; the codes feed the scan (native layout round trip)."""
    rng = np.random.default_rng(11)
    X = rng.standard_normal((200_000, 128)).astype(np.float32)
    pq = tinyknn.FastPQ(2, use_kmeans=False).fit(X[:5000])
    td = pq.transform(X, device=True)
    n, exp = O.pq_transform(O.PQState.from_pq(pq), X)
    assert n == td.size and _assert_codes_equal_up_to_ties(pq, X, td.packed, exp) <= 4
    nat = D.to_native(D.upload(td.packed), td.packed.shape[0], td.packed.shape[1])
    assert np.array_equal(D.from_native(nat, td.packed.shape[0], td.packed.shape[1]).cpu().numpy().view(np.uint64), td.packed)


def test_ivf_build_device_encoding_equals_host_build():
    """IVF.build with the one-launch GPU encoding == the list-by-list host encoding (same ids, same packed codes)."""
    np.random.seed(4)
    X = np.random.randn(3000, 32).astype(np.float32)
    a = tinyknn.IVF("angular", 20, tinyknn.FastPQ(2)).fit(X)
    import copy
    b = copy.deepcopy(a)
    a.build(X, n_probes=2, device=True, assign_device=False)
    b.build(X, n_probes=2, device=False)
    assert a.pq_transformed_centers.size == b.pq_transformed_centers.size
    for ta, tb, ia, ib in zip(a.pq_transformed_points, b.pq_transformed_points, a.ids, b.ids):
        assert np.array_equal(ia, ib)
        if isinstance(tb, tuple):
            assert ta.size == tb.size and np.array_equal(ta.packed, tb.packed)


# ---- chunk minima: the replay of long probe lists (tkb_ivf_scan_native_cm_dev / tkb_ivf_replay_fresh_cm_dev) ---------

def _host_cmin(packed_est, signd):
    v = packed_est.reshape(-1, 16)
    return (v.view(np.int8) if signd else v).min(axis=1).astype(np.int8 if signd else np.uint8).view(np.uint8)


@pytest.mark.parametrize("signd", [True, False])
def test_replay_cm_long_streams_equal_plain_replay_and_oracle(signd):
    """Streams of several thousand chunks (so that rounds past RQ_CM_MIN use the chunk minima), value distributions with
    many ties and with rare candidates, a skipped slot, an empty list, sizes that are not multiples of 16, heaps small
    enough to cut rounds: heap arrays of the cm replay == plain replay == oracle."""
    from tinyknn_b200._lib import lib, check, PROBE_SKIP, PLAN_SEND
    from tinyknn_b200 import sharded as SH
    rng = np.random.default_rng(77 + int(signd))
    for trial in range(4):
        n_lists = 12
        sizes = rng.integers(3000, 30000, size=n_lists).astype(np.int32)
        sizes[3], sizes[7] = 0, 17
        nc8 = (-(-sizes.astype(np.int64) // 128)) * 8
        off = np.zeros(n_lists + 1, np.int64)
        off[1:] = np.cumsum(nc8)
        ids = rng.permutation(16 * int(off[-1])).astype(np.int64) + 10 ** 10
        Q, P, R = 9, 6, [331, 40, 111, 7][trial]
        probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
        probes[2, 1] = PROBE_SKIP
        seg, gb, _ = SH.plan_host(probes, sizes, None, PLAN_SEND, 0, 1, 0)
        total = int(gb.sum())
        if trial % 2 == 0:                                           # smooth: candidates get rare as the bound settles
            est = np.clip(rng.normal(60, 25, size=total), 0, 126).astype(np.uint8)
        else:                                                        # few distinct values: ties everywhere, long queues
            est = (rng.integers(0, 6, size=total) * 9 + 40).astype(np.uint8)
        if signd:
            est = (est.astype(np.int16) - 64).astype(np.int8).view(np.uint8)
        cm = np.concatenate([_host_cmin(est, signd), np.zeros(16, np.uint8)])
        d_est, d_cm, d_seg = D.upload(np.concatenate([est, np.zeros(16, np.uint8)])), D.upload(cm), D.upload(seg)
        d_off, d_sizes, d_ids, d_probes = (D.upload(x) for x in (off, sizes, ids, probes))
        fb = D.empty((Q,), np.int32)
        h = [(D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)) for _ in range(2)]
        check(lib.tkb_ivf_replay_fresh_dev(D.ptr(d_est), 0, D.ptr(d_seg), D.ptr(d_off), D.ptr(d_sizes), n_lists, D.ptr(d_ids),
                                           D.ptr(d_probes), Q, P, D.ptr(h[0][0]), D.ptr(h[0][1]), R, int(signd), 1, D.ptr(fb), D.stream_ptr()))
        check(lib.tkb_ivf_replay_fresh_cm_dev(D.ptr(d_est), D.ptr(d_seg), D.ptr(d_cm), D.ptr(d_off), D.ptr(d_sizes), n_lists, D.ptr(d_ids),
                                              D.ptr(d_probes), Q, P, D.ptr(h[1][0]), D.ptr(h[1][1]), R, int(signd), 1, D.ptr(fb), D.stream_ptr()))
        a, b = h[0][0].cpu().numpy(), h[0][1].cpu().numpy()
        assert np.array_equal(h[1][0].cpu().numpy(), a) and np.array_equal(h[1][1].cpu().numpy(), b), trial
        for q in range(0, Q, 4):
            oi, ov = np.zeros(R, np.int64), np.zeros(R, np.int32)
            O.init_heap(oi, ov, signd)
            for s_ in range(P):
                l = int(probes[q, s_])
                if l == PROBE_SKIP or sizes[l] == 0:
                    continue
                ncr = -(-int(sizes[l]) // 16)
                O.replay(est[seg[q, s_]:seg[q, s_] + 16 * ncr], int(sizes[l]), oi, ov, signd,
                         np.ascontiguousarray(ids[16 * off[l]:16 * off[l] + 16 * ncr]))
            assert np.array_equal(a[q], oi) and np.array_equal(b[q], ov), (trial, q)


def test_query_batch_with_chunk_minima_equals_plain(monkeypatch):
    """End to end on an index with long lists: the cm scan writes the same estimates plus correct minima, and query_batch
    returns the same ids, distances and heap arrays with and without the chunk-minimum path."""
    from tinyknn_b200 import synth, ivf as ivf_mod
    X = synth.clustered(400_000 + 64, 64, 50, seed=5)
    ivf = synth.build_ivf(X[:400_000], "euclidean", 12, seed=5)
    qs = X[400_000:].contiguous()
    monkeypatch.setattr(ivf_mod, "CMIN_CHUNKS", 0)
    ref = ivf.query_batch(qs, 10, n_probes=6, order="device", return_distances=True, sub_batches=1)
    ref_heap = (ivf._last["heap_idx"].cpu().numpy(), ivf._last["heap_val"].cpu().numpy())
    monkeypatch.setattr(ivf_mod, "CMIN_CHUNKS", 1)
    got = ivf.query_batch(qs, 10, n_probes=6, order="device", return_distances=True, sub_batches=1)
    assert all(np.array_equal(x, y) for x, y in zip(ref, got))
    assert np.array_equal(ivf._last["heap_idx"].cpu().numpy(), ref_heap[0])
    assert np.array_equal(ivf._last["heap_val"].cpu().numpy(), ref_heap[1])
    # the minima the scan wrote: recompute from its estimates
    dev = ivf.to_device()
    lut = ivf.pq.distance_tables(qs, signed=True)
    Q, P = qs.shape[0], 6
    probes = ivf._coarse(dev, lut, Q, P, min(2 * P + 10, dev["C"]), "device")
    seg_off, gb = ivf._plan(dev, probes, Q, P)
    total = int(gb.cpu().numpy()[1])
    est, cmin = D.empty((total + 16,), np.uint8), D.empty((total // 16 + 16,), np.uint8)
    ivf._scan(dev, lut["tables"], probes, Q, P, est, seg_off, cmin=cmin)
    e, c = est.cpu().numpy()[:total], cmin.cpu().numpy()[:total // 16]
    assert np.array_equal(c, _host_cmin(e, True))
    assert total // 16 > 20_000                                     # long streams: the cm rounds really ran
