"""CPU-side tests of the boundary and the host logic (no GPU needed, no compute calls)."""
import ctypes
import heapq
import os
import pickle
import re

import numpy as np
import pytest

import tinyknn_b200 as tinyknn
from tinyknn_b200 import _lib, _transform, utils
from tinyknn_b200._fast_pq import insert, init_heap, insert_is, estimate_pq_sse, query_pq_sse
from tinyknn_b200._fast_pq_avx import estimate_pq_avx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "tinyknn_b200.h")).read()
    declared = set(re.findall(r"TKB_API\s+[\w\s\*]*?\b(tkb_\w+)\s*\(", hdr))
    assert len(declared) >= 17
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    so = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), name
    assert _lib.lib.tkb_version() >= 100


def test_export_list_matches_reference_package():        # ref: tinyknn/__init__.py:1-6
    for name in ("_transform", "_fast_pq", "FastPQ", "avx", "IVF", "utils", "bottom_k", "bottom_k_2d", "cdist",
                 "knn_brute", "group_data_by_indices"):
        assert hasattr(tinyknn, name), name
    assert tinyknn.avx is True and tinyknn.fast_pq.dpad == 4


def test_buffer_checks_raise_like_cython():
    good = dict(data=np.zeros((1, 4), np.uint64), tables=np.zeros(8, np.uint64), out=np.zeros(2, np.uint64))
    with pytest.raises(ValueError):
        estimate_pq_sse(good["data"].astype(np.int64), good["tables"], good["out"], True)
    with pytest.raises(ValueError):
        estimate_pq_sse(good["data"], good["tables"].reshape(2, 4), good["out"], True)
    with pytest.raises(ValueError):
        estimate_pq_avx(np.zeros((2, 8), np.uint64)[:, ::2], good["tables"], good["out"], True)
    with pytest.raises(ValueError):
        query_pq_sse(good["data"], 1, good["tables"], np.zeros(3, np.int32), np.zeros(3, np.int32), True)


def test_no_cpu_fallback_without_device():
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        estimate_pq_sse(np.zeros((1, 4), np.uint64), np.zeros(8, np.uint64), np.zeros(2, np.uint64), True)
    pq = tinyknn.FastPQ(2)
    pq.fit(np.random.randn(64, 8).astype(np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pq.distance_table(np.zeros(8, np.float32))


# ---- heap host functions (ref: tests/test_heap.py) ---------------------------------------------------

class Heap:
    def __init__(self, size):
        self.indices = np.empty((size,), dtype=np.int64)
        self.vals = np.empty((size,), dtype=np.int32)
        init_heap(self.indices, self.vals, signd=True)

    def insert(self, i, v):
        if v < self.peek():
            insert(self.indices, self.vals, i, v)

    def peek(self):
        return self.vals[0]


def test_heap_init_and_known_answers():
    h = Heap(3)
    assert h.indices.tolist() == [-1] * 3 and h.vals.tolist() == [127] * 3
    h = Heap(1); h.insert(1, 10)
    assert h.indices.tolist() == [1] and h.vals.tolist() == [10]
    h = Heap(2); h.insert(1, 10)
    assert h.indices.tolist() == [-1, 1] and h.vals.tolist() == [127, 10]
    h.insert(1, 10)
    assert h.indices.tolist() == [-1, 1] and h.vals.tolist() == [127, 10]
    iu, vu = np.empty(2, np.int64), np.empty(2, np.int32)
    init_heap(iu, vu, False)
    assert vu.tolist() == [255, 255]


def test_heap_random_vs_heapq():
    rng = np.random.RandomState(10)
    heap = Heap(10)
    pyheap = [(-127, -1)] * 10
    for t in range(1000):
        top = -pyheap[0][0]
        assert top == heap.peek()
        v = rng.randint(10000 // (t + 1))
        heap.insert(t, v)
        if v < top:
            heapq.heappop(pyheap)
            heapq.heappush(pyheap, (-v, t))
        assert set(heap.vals) == {-vi for vi, _ in pyheap}


def _is_max_heap(vals, root=0):
    n = len(vals)
    return all(vals[c] <= vals[root] and _is_max_heap(vals, c) for c in (2 * root + 1, 2 * root + 2) if c < n)


def test_heap_property_and_oracle_layout():
    from oracle import restate as O
    rng = np.random.RandomState(13)
    for n in range(1, 12):
        for seq in (list(range(n)), list(reversed(range(n))), [rng.randint(n) for _ in range(n)]):
            heap = Heap(n)
            oi, ov = np.empty(n, np.int64), np.empty(n, np.int32)
            O.init_heap(oi, ov, True)
            for i, v in enumerate(seq):
                heap.insert(i, v)
                if v < ov[0]:
                    O.insert(oi, ov, i, v)
                assert v in heap.vals and i in heap.indices and _is_max_heap(heap.vals)
                assert np.array_equal(heap.indices, oi) and np.array_equal(heap.vals, ov)


def test_insert_is_matches_oracle():
    from oracle import restate as O
    rng = np.random.RandomState(2)
    a, b = (np.empty(6, np.int64), np.empty(6, np.int32)), (np.empty(6, np.int64), np.empty(6, np.int32))
    init_heap(*a, True); O.init_heap(*b, True)
    for t in range(200):
        v = int(rng.randint(0, 120))
        if v < a[1][0]:
            insert_is(a[0], a[1], t % 50, v); O.insert_is(b[0], b[1], t % 50, v)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


# ---- layouts / utils ------------------------------------------------------------------------------------

def test_transform_matches_oracle_and_roundtrips():
    from oracle import restate as O
    rng = np.random.default_rng(10)
    for n, d in ((16, 2), (16 * 13, 14), (64, 32), (48, 52)):
        codes = rng.integers(16, size=(n, d)).astype(np.uint8)
        packed = _transform.transform_data(codes)
        assert packed.dtype == np.uint64 and packed.shape == (n // 16, d) and packed.flags.c_contiguous
        assert np.array_equal(packed, O.transform_data(codes))
        assert np.array_equal(_transform.unpack(packed), codes)
    tab = rng.integers(256, size=(6, 16)).astype(np.uint8)
    assert np.array_equal(_transform.transform_tables(tab), O.transform_tables(tab))
    with pytest.raises(AssertionError):
        _transform.transform_data(np.zeros((15, 2), np.uint8))


def test_utils_match_reference_behaviour():               # ref: tests/test_utils.py
    rng = np.random.default_rng(8)
    X, Y = rng.standard_normal((30, 7)), rng.standard_normal((11, 7))
    ref = ((X[:, None, :] - Y[None]) ** 2).sum(-1)
    assert np.allclose(utils.cdist(X, Y), ref)
    nn = utils.knn_brute(X, Y, 3)
    assert all(set(nn[i]) == set(np.argsort(ref[i])[:3]) for i in range(30))
    Xn, Yn = X / np.linalg.norm(X, axis=1, keepdims=True), Y / np.linalg.norm(Y, axis=1, keepdims=True)
    assert np.array_equal(np.sort(utils.knn_brute(X, Y, 2, metric="angular")), np.sort(utils.knn_brute(Xn, Yn, 2)))
    assert utils.pad1(np.ones(5), 4).shape == (8,) and utils.pad2(np.ones((3, 5)), 16, 4).shape == (16, 8)
    assert list(utils.bottom_k(np.array([3., 1., 2.]), 5)) == [0, 1, 2]
    Xs = np.array([[1], [2], [3], [4]])
    parts, ids = utils.group_data_by_indices(Xs, np.array([[0, 1], [1, 2], [0, 2], [0, 1]]), 3)
    assert [p.ravel().tolist() for p in parts] == [[1, 3, 4], [2, 1, 4], [2, 3]]
    assert [i.tolist() for i in ids] == [[0, 2, 3], [1, 0, 3], [1, 2]]


def test_fit_transform_host_side_and_pickle():
    np.random.seed(10)
    X = np.random.randn(100, 10).astype(np.float32)
    pq = tinyknn.FastPQ(2)
    n0, t0 = pq.fit_transform(X)
    n1, t1 = pq.transform(X)
    assert n0 == n1 == 100 and np.array_equal(t0, t1) and t0.shape == (7, 8)
    assert pq.centers.shape == (16, 16) and pq.centers.dtype == np.float32 and pq.R.shape == (16, 16)
    ivf = tinyknn.IVF("euclidean", 5, tinyknn.FastPQ(2))
    ivf.fit(X).build(X, n_probes=2)
    ivf2 = pickle.loads(pickle.dumps(ivf))                # ref: examples/bench.py:88-103
    assert np.array_equal(ivf2.active_centers, ivf.active_centers)
    assert sum(t.size for t in ivf.pq_transformed_points) == 200
    with pytest.raises(AssertionError):
        tinyknn.FastPQ(2).fit(np.zeros((0, 4), np.float32))
    with pytest.raises(AssertionError):
        tinyknn.IVF("manhattan", 3)


def test_index_save_load_round_trip(tmp_path):
    """Stable on-disk format (tinyknn_b200/io.py): every attribute the query path and the oracle read survives a
    save/load, memory-mapped or not; lists are views into one codes / ids file; empty and never-filled lists keep
    their reference representation; a half-written directory (no meta.json) and a foreign dpad are refused."""
    from tinyknn_b200 import fast_pq as fp
    np.random.seed(3)
    X = np.random.randn(700, 20).astype(np.float32)
    ivf = tinyknn.IVF("angular", 9, tinyknn.FastPQ(2))
    ivf.fit(X).build(X, n_probes=2, device=False)
    ivf.pq_transformed_points.append(None)                         # a slot the build never filled
    ivf.ids.append(None)
    ivf.n_clusters += 1
    path = str(tmp_path / "idx")
    tinyknn.save_index(ivf, path)
    for mmap in (True, False):
        got = tinyknn.load_index(path, mmap=mmap)
        assert got.metric == ivf.metric and got.n_clusters == ivf.n_clusters
        assert got.pq.dims_per_block == 2 and got.pq.rotate_dim == ivf.pq.rotate_dim and got.pq.use_kmeans == ivf.pq.use_kmeans
        assert got.pq.sqrt_n_blocks == ivf.pq.sqrt_n_blocks
        for name in ("centers", "R"):
            assert np.array_equal(getattr(got.pq, name), getattr(ivf.pq, name))
        assert np.array_equal(got.all_centers, ivf.all_centers) and np.array_equal(got.active_centers, ivf.active_centers)
        assert got.active_centers.dtype == np.float32 and got.active_centers.flags.c_contiguous
        assert got.pq_transformed_centers.size == ivf.pq_transformed_centers.size
        assert np.array_equal(got.pq_transformed_centers.packed, ivf.pq_transformed_centers.packed)
        assert np.array_equal(got.data, ivf.data)
        assert len(got.pq_transformed_points) == len(ivf.pq_transformed_points)
        for a, b, ia, ib in zip(got.pq_transformed_points, ivf.pq_transformed_points, got.ids, ivf.ids):
            if b is None:
                assert a is None and ia is None
            elif not isinstance(b, tuple):
                assert not isinstance(a, tuple) and a.size == 0
            else:
                assert a.size == b.size and a.packed.dtype == np.uint64 and np.array_equal(a.packed, b.packed)
                assert np.array_equal(ia, ib) and ia.dtype == np.int64
        pickle.loads(pickle.dumps(got))                            # a loaded index still pickles like the reference's
    no_data = str(tmp_path / "idx_nodata")
    tinyknn.save_index(ivf, no_data, include_data=False)
    assert not os.path.exists(os.path.join(no_data, "data.npy"))
    assert np.array_equal(tinyknn.load_index(no_data, data=ivf.data).data, ivf.data)
    os.remove(os.path.join(no_data, "meta.json"))
    with pytest.raises(FileNotFoundError):
        tinyknn.load_index(no_data)
    fp.set_order("sse")
    try:
        with pytest.raises(ValueError):
            tinyknn.load_index(path)
    finally:
        fp.set_order("avx")


def test_loaded_index_is_queried_identically_by_the_oracle(tmp_path):
    """The on-disk format keeps the reference's attribute types: the CPU oracle answers from a memory-mapped index
    exactly as from the original."""
    from oracle import restate as O
    np.random.seed(5)
    X = np.random.randn(900, 24).astype(np.float32)
    ivf = tinyknn.IVF("euclidean", 8, tinyknn.FastPQ(2))
    ivf.fit(X).build(X, n_probes=1, device=False)
    path = tinyknn.save_index(ivf, str(tmp_path / "idx"))
    got = tinyknn.load_index(path)
    K = O.Kernels("port", "avx")
    Sa, Sb = O.IVFState.from_ivf(ivf), O.IVFState.from_ivf(got)
    for q in np.random.randn(12, 24).astype(np.float32):
        ta, tb = {}, {}
        a = O.ivf_query(Sa, q, 5, n_probes=3, kernels=K, trace=ta)
        b = O.ivf_query(Sb, q, 5, n_probes=3, kernels=K, trace=tb)
        assert np.array_equal(a, b) and np.array_equal(ta["heap_indices"], tb["heap_indices"])
