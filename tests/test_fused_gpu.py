"""GPU parity tests of the fused per-query kernel (tkb_ivf_query_fused_dev): everything after probe selection
(ref: tinyknn/ivf.py:135-163) in one launch. Checked against the oracle and against the stage-by-stage kernels.
"""
import ctypes

import numpy as np
import pytest

import tinyknn_b200 as tinyknn
from tinyknn_b200 import _device as D
from tinyknn_b200._lib import lib, check, PROBE_SKIP, DTYPE_F32, DTYPE_F64, ORDER_AVX, ORDER_SSE
from tinyknn_b200._transform import transform_data
from oracle import restate as O

from test_gpu_parity import _ivf_from_state

pytestmark = pytest.mark.gpu


def test_fused_equals_staged_and_oracle(golden):
    """IVF.query_batch in throughput mode: fused == stage-by-stage kernels bit for bit (ids, counts, distances);
    heap arrays and ids == the oracle run with the same selection rule."""
    z = golden["ivf"]
    K = O.Kernels("port", "avx")
    for name in z["names"]:
        S = O.ivf_state_from_arrays(z, name + "_")
        ivf = _ivf_from_state(S)
        ivf._keep_heaps = True
        qs = z[name + "_q"]
        for npr in (1, 4, 8):
            a = ivf.query_batch(qs, 10, n_probes=npr, order="device", return_distances=True, fused=True)
            heaps = ivf._last["heap_idx"].cpu().numpy(), ivf._last["heap_val"].cpu().numpy()
            if not ivf.to_device()["unique_ids"]:      # labels repeat across lists: query_batch keeps the dedupe-capable staged path
                assert "fused_ws" not in ivf._last
                continue
            assert "fused_ws" in ivf._last
            b = ivf.query_batch(qs, 10, n_probes=npr, order="device", return_distances=True, fused=False)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (name, npr)
            staged = ivf._last["heap_idx"].cpu().numpy(), ivf._last["heap_val"].cpu().numpy()
            assert np.array_equal(heaps[0], staged[0]) and np.array_equal(heaps[1], staged[1])
            bad = 0
            for i, q in enumerate(qs):
                tr = {}
                exp = O.ivf_query(S, q, 10, n_probes=npr, kernels=K, select=O.bottom_k_sorted, trace=tr)
                bad += set(a[0][i][:a[1][i]]) != set(exp)
                assert np.array_equal(heaps[0][i], tr["heap_indices"]) and np.array_equal(heaps[1][i], tr["heap_values"])
            assert bad == 0


def _random_index(rng, M, n_lists, max_size, d, dtype, min_size=0):
    sizes = rng.integers(min_size, max_size, size=n_lists).astype(np.int32)
    if min_size == 0:
        sizes[rng.integers(0, n_lists)] = 0
        sizes[rng.integers(0, n_lists)] = 1
    sizes[n_lists - 1] = max(int(sizes[n_lists - 1]), 17)               # the list negative probes wrap to
    nc = -(-sizes.astype(np.int64) // 16)
    nc8 = -(-nc // 8) * 8
    off = np.zeros(n_lists + 1, np.int64)
    off[1:] = np.cumsum(nc8)
    total = int(off[-1])
    codes = rng.integers(0, 16, size=(16 * total, M), dtype=np.uint8)
    packed = transform_data(codes)                                       # (total, M) uint64, reference layout
    n_rows = int(sizes.sum()) + 3
    ids = np.full(16 * total, -1, np.int64)
    perm = rng.permutation(n_rows)
    pos = 0
    for l in range(n_lists):
        ids[16 * off[l]:16 * off[l] + sizes[l]] = perm[pos:pos + sizes[l]]
        pos += sizes[l]
    rows = rng.standard_normal((n_rows, d)).astype(dtype)
    return sizes, off, packed, ids, rows


def _tables(rng, Q, M, kind):
    if kind == "typical":                                                # like distance_table output: [-4, 23]
        t = rng.integers(-4, 24, size=(Q, M, 16))
    elif kind == "low":                                                  # sums stay well inside int8: every vector is a candidate at first
        t = rng.integers(-5, 7, size=(Q, M, 16))
    elif kind == "hot":                                                  # sums run into the +127 clamp: certificate fails often
        t = rng.integers(-2, 12, size=(Q, M, 16))
    else:                                                                # full range: the fast path is not eligible
        t = rng.integers(-128, 128, size=(Q, M, 16))
    return t.astype(np.int8).view(np.uint8)


@pytest.mark.parametrize("trial", range(8))
def test_fused_random_index_matches_oracle(trial):
    """C-ABI level: random codes / LUTs / lists (empty, one vector), skipped and Python-wrapped (negative, repeating)
    probe slots, f32 and f64 rows, both accumulation orders: heap arrays == oracle slot for slot, ids and distances ==
    the stage-by-stage kernels."""
    rng = np.random.default_rng(100 + trial)
    order = "avx" if trial % 4 != 3 else "sse"
    M = int(rng.choice([4, 8, 32, 52])) if order == "avx" else int(rng.choice([2, 6, 32]))
    n_lists = int(rng.integers(3, 30))
    d = int(rng.choice([3, 32, 100]))
    dtype = np.float32 if trial % 3 else np.float64
    sizes, off, packed, ids, rows = _random_index(rng, M, n_lists, 900 if trial % 2 else 120, d, dtype)
    Q = int(rng.integers(1, 70))
    P = int(rng.integers(1, min(n_lists, 9) + 1))
    R = int(rng.integers(1, 140))
    k = int(rng.integers(1, 12))
    kind = ["typical", "hot", "full"][trial % 3]
    tables = _tables(rng, Q, M, kind)
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
    if trial % 2 == 0:
        probes[rng.integers(0, Q), P - 1] = PROBE_SKIP
    for _ in range(4):                                                   # wrapped entries: -1 == the last list, which may also be
        q = int(rng.integers(0, Q))                                      # there under its own index, or wrapped twice (non-negative
        probes[q, rng.integers(0, P)] = -1                               # entries are distinct: they are positions of a heap)
        if P > 1 and rng.random() < 0.7:
            probes[q, rng.integers(0, P)] = -1
    queries = rng.standard_normal((Q, d)).astype(np.float32)
    o = ORDER_AVX if order == "avx" else ORDER_SSE
    ddt = DTYPE_F32 if dtype == np.float32 else DTYPE_F64
    mlc = int(max(1, (-(-sizes.astype(np.int64) // 16)).max()))

    n_chunks = packed.shape[0]
    nat = D.to_native(D.upload(packed), n_chunks, M)
    d_off, d_sizes, d_ids, d_probes, d_tab, d_rows, d_q = (D.upload(x) for x in (off, sizes, ids, probes, tables, rows, queries))
    need = ctypes.c_int64(0)
    check(lib.tkb_ivf_query_fused_workspace(Q, P, R, M, o, ddt, mlc, ctypes.byref(need)))
    ws = D.empty((need.value,), np.uint8)
    oi, od, oc = D.empty((Q, k), np.int64), D.empty((Q, k), dtype), D.empty((Q,), np.int32)
    hi, hv = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)
    check(lib.tkb_ivf_query_fused_dev(D.ptr(nat), D.ptr(d_off), D.ptr(d_sizes), n_lists, M, D.ptr(d_tab), D.ptr(d_probes), Q, P,
                                      D.ptr(d_ids), D.ptr(d_rows), ddt, rows.shape[0], d, D.ptr(d_q), R, k, o, mlc,
                                      D.ptr(oi), D.ptr(od), D.ptr(oc), D.ptr(hi), D.ptr(hv), D.ptr(ws), ws.numel(), D.stream_ptr()))
    hi_h, hv_h = hi.cpu().numpy(), hv.cpu().numpy()

    # oracle: estimates of each probed list, replay in probe order with the reference's label dedupe
    for q in range(Q):
        ei, ev = np.zeros(R, np.int64), np.zeros(R, np.int32)
        O.init_heap(ei, ev, True)
        tq = np.ascontiguousarray(tables[q].reshape(-1).view(np.uint64))
        for s in range(P):
            l = int(probes[q, s])
            if l == PROBE_SKIP:
                continue
            if l < 0:
                l += n_lists
            if sizes[l] == 0:
                continue
            ncr = -(-int(sizes[l]) // 16)
            data = np.ascontiguousarray(packed[off[l]:off[l] + ncr])
            O.query_pq(data, int(sizes[l]), tq, ei, ev, True, np.ascontiguousarray(ids[16 * off[l]:16 * off[l] + 16 * ncr]), order=order)
        assert np.array_equal(hi_h[q], ei) and np.array_equal(hv_h[q], ev), (trial, q)

    # stage-by-stage selection over the same heaps == fused outputs
    dd = D.empty((Q, R), dtype)
    check(lib.tkb_gather_dists_dev(D.ptr(d_rows), ddt, rows.shape[0], d, D.ptr(d_q), D.ptr(hi), Q, R, D.ptr(dd), D.stream_ptr()))
    si, sd, sc = D.empty((Q, k), np.int64), D.empty((Q, k), dtype), D.empty((Q,), np.int32)
    check(lib.tkb_select_topk_dev(D.ptr(hi), D.ptr(dd), ddt, Q, R, k, D.ptr(si), D.ptr(sd), D.ptr(sc), D.stream_ptr()))
    assert np.array_equal(oc.cpu().numpy(), sc.cpu().numpy())
    assert np.array_equal(oi.cpu().numpy(), si.cpu().numpy())
    assert np.array_equal(od.cpu().numpy(), sd.cpu().numpy())
    # distances follow knn_brute1's formula (ref: utils.py:89-91) within 1e-5 relative
    ids_h, cnt_h, dst_h = oi.cpu().numpy(), oc.cpu().numpy(), od.cpu().numpy()
    for q in range(Q):
        got = ids_h[q][:cnt_h[q]]
        if len(got):
            ref_d = O.exact_dists(queries[q], rows[got])
            np.testing.assert_allclose(dst_h[q][:cnt_h[q]], ref_d, rtol=1e-5)


def test_fused_large_lists_many_rounds():
    """Lists of several thousand vectors and a small heap: the queue replay has to cut rounds and resume."""
    rng = np.random.default_rng(7)
    M, n_lists, d = 32, 6, 16
    sizes, off, packed, ids, rows = _random_index(rng, M, n_lists, 6000, d, np.float32, min_size=3000)
    Q, P, R, k = 40, 4, 25, 10
    tables = _tables(rng, Q, M, "low")
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
    queries = rng.standard_normal((Q, d)).astype(np.float32)
    mlc = int(max(1, (-(-sizes.astype(np.int64) // 16)).max()))
    nat = D.to_native(D.upload(packed), packed.shape[0], M)
    d_off, d_sizes, d_ids, d_probes, d_tab, d_rows, d_q = (D.upload(x) for x in (off, sizes, ids, probes, tables, rows, queries))
    need = ctypes.c_int64(0)
    check(lib.tkb_ivf_query_fused_workspace(Q, P, R, M, ORDER_AVX, DTYPE_F32, mlc, ctypes.byref(need)))
    ws = D.empty((need.value,), np.uint8)
    oi, od, oc = D.empty((Q, k), np.int64), D.empty((Q, k), np.float32), D.empty((Q,), np.int32)
    hi, hv = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32)
    check(lib.tkb_ivf_query_fused_dev(D.ptr(nat), D.ptr(d_off), D.ptr(d_sizes), n_lists, M, D.ptr(d_tab), D.ptr(d_probes), Q, P,
                                      D.ptr(d_ids), D.ptr(d_rows), DTYPE_F32, rows.shape[0], d, D.ptr(d_q), R, k, ORDER_AVX, mlc,
                                      D.ptr(oi), D.ptr(od), D.ptr(oc), D.ptr(hi), D.ptr(hv), D.ptr(ws), ws.numel(), D.stream_ptr()))
    hi_h, hv_h = hi.cpu().numpy(), hv.cpu().numpy()
    for q in range(Q):
        ei, ev = np.zeros(R, np.int64), np.zeros(R, np.int32)
        O.init_heap(ei, ev, True)
        tq = np.ascontiguousarray(tables[q].reshape(-1).view(np.uint64))
        for s in range(P):
            l = int(probes[q, s])
            if sizes[l] == 0:
                continue
            ncr = -(-int(sizes[l]) // 16)
            O.query_pq(np.ascontiguousarray(packed[off[l]:off[l] + ncr]), int(sizes[l]), tq, ei, ev, True,
                       np.ascontiguousarray(ids[16 * off[l]:16 * off[l] + 16 * ncr]))
        assert np.array_equal(hi_h[q], ei) and np.array_equal(hv_h[q], ev), q
    assert int(oc.cpu().numpy().min()) == k

