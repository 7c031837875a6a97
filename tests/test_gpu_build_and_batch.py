"""GPU tests of the build-time kernel tkb_assign_dev (IVF.build's coarse assignment), the chunk minima inside the push
exchange, the one-kernel probe selection (tkb_coarse_probes_dev), the saved-index query, CUDA-graph and asynchronous batches.
Written at the end of round 1 (first run on the CPU emulator, tests/emulate), validated on a B200 at the start of round 2:
20 passed (profiles/r2_gputest_first_call.txt)."""
import os

import numpy as np
import pytest

import tinyknn_b200 as tinyknn
from oracle import restate as O
from tinyknn_b200 import _device as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric,build_probes", [("angular", 1), ("euclidean", 1), ("angular", 2)])
def test_ivf_build_device_equals_host_build(metric, build_probes):
    """IVF.build with the GPU assignment (tkb_assign_dev) and the one-launch GPU encoding (tkb_encode_dev) == the host
    build: same lists, ids in the same order and identical packed codes for build_probes = 1; for 2 lists per point the
    same id SETS per list (np.argpartition's column order is unspecified) with the same code per id."""
    import copy
    np.random.seed(4)
    X = np.random.randn(3000, 32).astype(np.float32)
    a = tinyknn.IVF(metric, 20, tinyknn.FastPQ(2)).fit(X)
    b = copy.deepcopy(a)
    a.build(X, n_probes=build_probes, device=True, assign_device=True)
    b.build(X, n_probes=build_probes, device=False)
    assert np.array_equal(a.active_centers, b.active_centers)
    assert a.pq_transformed_centers.size == b.pq_transformed_centers.size
    assert np.array_equal(a.pq_transformed_centers.packed, b.pq_transformed_centers.packed)
    for ta, tb, ia, ib in zip(a.pq_transformed_points, b.pq_transformed_points, a.ids, b.ids):
        if not isinstance(tb, tuple):
            continue
        assert ta.size == tb.size
        if build_probes == 1:
            assert np.array_equal(ia, ib) and np.array_equal(ta.packed, tb.packed)
        else:
            oa, ob = np.argsort(ia, kind="stable"), np.argsort(ib, kind="stable")
            assert np.array_equal(np.asarray(ia)[oa], np.asarray(ib)[ob])
            ca, cb = O.unpack(ta.packed)[:ta.size], O.unpack(tb.packed)[:tb.size]
            assert np.array_equal(ca[oa], cb[ob])


@pytest.mark.parametrize("n,d,C,dtype,metric", [(5000, 100, 1087, np.float32, "angular"), (3000, 128, 300, np.float32, "euclidean"),
                                                (1000, 20, 37, np.float64, "euclidean"), (2000, 384, 64, np.float32, "euclidean"),
                                                (130, 7, 3, np.float32, "euclidean")])
def test_assign_device_equals_knn_brute(n, d, C, dtype, metric):
    """tkb_assign_dev vs the reference's knn_brute arithmetic (oracle restatement): identical nearest centroid for every
    row (a difference is only tolerated on an exact tie of the reference's own `part` values); k = 2: same pairs."""
    from tinyknn_b200.utils import knn_brute_device
    rng = np.random.default_rng(n + d)
    means = rng.standard_normal((C, d)) * 2
    X = (means[rng.integers(C, size=n)] + rng.standard_normal((n, d))).astype(dtype)
    Y = (means + 0.1 * rng.standard_normal((C, d))).astype(dtype)
    Xr, Yr = (X, Y) if metric == "euclidean" else (X / np.linalg.norm(X, axis=1, keepdims=True), Y / np.linalg.norm(Y, axis=1, keepdims=True))
    part = np.einsum("ij,ij->i", Xr, Xr)[:, None] + np.einsum("ij,ij->i", Yr, Yr)[None] - 2 * Xr @ Yr.T      # utils.py:80-83
    exp1 = O.knn_brute(Xr, Yr, 1)[:, 0]
    got1 = knn_brute_device(X, Y, 1, metric=metric)[:, 0]
    bad = np.nonzero(got1 != exp1)[0]
    assert all(part[r, got1[r]] == part[r, exp1[r]] for r in bad) and len(bad) <= 2
    if C >= 2:
        exp2 = np.sort(O.knn_brute(Xr, Yr, 2), axis=1)
        got2 = knn_brute_device(X, Y, 2, metric=metric)
        assert np.all(part[np.arange(n), got2[:, 0]] <= part[np.arange(n), got2[:, 1]])           # ascending order
        bad = np.nonzero(np.any(np.sort(got2, axis=1) != exp2, axis=1))[0]
        assert len(bad) <= 2




def _inertia(X, C):
    d2 = (np.einsum("ij,ij->i", X, X)[:, None] + np.einsum("ij,ij->i", C, C)[None, :] - 2.0 * X @ C.T)
    return float(d2.min(axis=1).sum())


def test_fit_device_pq_codebooks_quality_and_determinism():
    """FastPQ.fit(device=True): the M per-block k-means problems on the GPU (tkb_kmeans_pq_dev). Parity with the reference is
    unpinned by nature (its fit is random); the clustering must be as good as sklearn's on the same blocks (inertia within
    10 %), a deterministic function of (data, seed), and usable: transform + distance tables + top() find the true neighbours
    as often as with the host fit."""
    import sklearn.cluster
    from tinyknn_b200 import FastPQ
    rng = np.random.default_rng(5)
    means = rng.normal(size=(40, 16)) * 3
    X = (means[rng.integers(40, size=6000)] + rng.normal(size=(6000, 16))).astype(np.float32)
    pq_a = FastPQ(2, rotate_dim=None).fit(X, device=True, seed=3)
    pq_b = FastPQ(2, rotate_dim=None).fit(X, device=True, seed=3)
    assert np.array_equal(pq_a.centers, pq_b.centers) and np.isfinite(pq_a.centers).all()
    assert pq_a.centers.shape == (16, 16)
    for m in range(8):
        blk = X[:, 2 * m:2 * m + 2].astype(np.float64)
        sk = sklearn.cluster.KMeans(16, n_init=2, random_state=0).fit(blk)
        assert _inertia(blk, pq_a.centers[:, 2 * m:2 * m + 2].astype(np.float64)) <= 1.10 * sk.inertia_
    pq_h = FastPQ(2, rotate_dim=None).fit(X, device=False)
    qs = X[:200] + 0.05 * rng.normal(size=(200, 16)).astype(np.float32)
    hits = []
    for pq in (pq_a, pq_h):
        td = pq.transform(X)
        hits.append(sum(i in pq.distance_table(q).top(td, X, k=1, rescore=30) for i, q in enumerate(qs)))
    assert hits[0] >= hits[1] - 10 and hits[0] >= 150


def test_fit_device_ivf_centroids():
    """IVF.fit(device=True): Lloyd's k-means on the GPU (tkb_kmeans_dev: the exact-chain assignment kernel + fixed-point sums).
    Planted, well separated clusters are recovered, the returned assignment is the nearest centre of every row, the result is
    deterministic, the inertia is sklearn's within 5 %, and the index built from it answers queries."""
    import sklearn.cluster
    from tinyknn_b200 import IVF
    rng = np.random.default_rng(11)
    k, d, n = 24, 20, 6000
    means = rng.normal(size=(k, d)) * 6
    lab = rng.integers(k, size=n)
    X = (means[lab] + rng.normal(size=(n, d))).astype(np.float32)
    a = IVF("euclidean", k).fit(X, device=True, seed=1)
    b = IVF("euclidean", k).fit(X, device=True, seed=1)
    assert np.array_equal(a.all_centers, b.all_centers) and a.all_centers.shape == (k, d)
    assert 1 <= a.fit_iters <= 25
    d2 = ((X[:, None, :].astype(np.float64) - a.all_centers[None].astype(np.float64)) ** 2).sum(2)
    srt = np.sort(d2, axis=1)
    clear = srt[:, 1] - srt[:, 0] > 1e-3 * srt[:, 1]                      # rows whose nearest centre is not a near-tie
    assert np.array_equal(a._fit_assign[clear], d2.argmin(1)[clear])
    near = ((means[:, None, :] - a.all_centers[None].astype(np.float64)) ** 2).sum(2).min(1)
    assert (near < 1.0).all()                                             # every planted mean has a centre next to it
    sk = sklearn.cluster.KMeans(k, n_init=1, random_state=0).fit(X.astype(np.float64))
    assert _inertia(X.astype(np.float64), a.all_centers.astype(np.float64)) <= 1.05 * sk.inertia_
    h = IVF("euclidean", k).fit(X, device=False)                          # the reference's sklearn path
    found = []
    for ivf in (a, h):
        ivf.build(X, n_probes=1)
        found.append(sum(i in ivf.query(X[i], k=10, n_probes=2) for i in range(100)))
    assert found[0] >= found[1] - 10 and found[0] >= 50                   # as useful an index as the host fit's


def test_kmeans_rejects_bad_arguments():
    import ctypes
    from tinyknn_b200._lib import lib
    need = ctypes.c_int64(0)
    assert lib.tkb_kmeans_workspace(100, 8, 4, ctypes.byref(need)) == 0 and need.value > 0
    x = D.upload(np.zeros((100, 8), np.float32))
    c = D.upload(np.zeros((4, 8), np.float32))
    asg = D.empty((100,), np.int32)
    ws = D.empty((need.value,), np.uint8)
    assert lib.tkb_kmeans_dev(D.ptr(x), 100, 8, 4, D.ptr(c), 3, 1.0, D.ptr(asg), None, D.ptr(ws), 16, D.stream_ptr()) != 0   # workspace too small
    assert lib.tkb_kmeans_dev(None, 100, 8, 4, D.ptr(c), 3, 1.0, D.ptr(asg), None, D.ptr(ws), ws.numel(), D.stream_ptr()) != 0
    assert lib.tkb_kmeans_pq_dev(D.ptr(x), 100, 8, 3, D.ptr(c), 3, 1.0, D.ptr(ws), ws.numel(), D.stream_ptr()) != 0           # 8 % 3 != 0
    assert lib.tkb_kmeans_pq_dev(D.ptr(x), 100, 8, 16, D.ptr(c), 3, 1.0, D.ptr(ws), ws.numel(), D.stream_ptr()) != 0          # dims_per_block > 8


def test_push_exchange_with_chunk_minima_single_gpu(monkeypatch):
    """The push exchange with the minima region, driven rank by rank on one GPU over an index with long lists: every
    "rank" stores estimates AND chunk minima into the home buffers; the cm replay on the home rank == the unsharded path."""
    import torch
    from tinyknn_b200 import synth, ivf as ivf_mod, sharded as SH
    X = synth.clustered(300_000 + 48, 64, 40, seed=9)
    ivf = synth.build_ivf(X[:300_000], "euclidean", 12, seed=9)
    qs = X[300_000:].contiguous()
    G, k, n_probes = 2, 10, 6
    Qh = qs.shape[0] // G
    monkeypatch.setattr(ivf_mod, "CMIN_CHUNKS", 0)
    ref = ivf.query_batch(qs, k, n_probes=n_probes, order="device", return_distances=True, sub_batches=1)
    ref_heap = ivf._last["heap_idx"].cpu().numpy()
    monkeypatch.setattr(ivf_mod, "CMIN_CHUNKS", 1)
    shards = [SH.ShardedIVF(ivf, rank=r, world=G, drop_full_codes=False) for r in range(G)]
    homes = [sh._home(qs[r * Qh:(r + 1) * Qh].contiguous(), n_probes) for r, sh in enumerate(shards)]
    P = homes[0]["P"]
    tables = torch.cat([h["lut"]["tables"] for h in homes])
    probes = torch.cat([h["probes"] for h in homes])
    cap = shards[0].push_capacity(Qh, P)
    bufs = [SH.PeerBuffers(cap, rank=0, world=1, n_buf=1) for _ in range(G)]
    try:
        addr = np.array([b.local[0].address for b in bufs], dtype=np.int64)
        home_base = D.upload(addr)
        cm_table = D.upload(addr + bufs[0].nbytes - (addr >> 4))
        for sh in shards:
            sh._scan_push(tables, probes, Qh, P, home_base, cm_table)
        for b, sh in enumerate(shards):
            seg_r, gb = ivf._plan(sh.dev, homes[b]["probes"], Qh, P)
            total = int(gb.cpu().numpy()[1])
            ids, cnt, dst = sh._finish(homes[b], bufs[b].local[0], seg_r, k, (n_probes + 1) * k + 1, bufs[b].local_cmin[0])
            sl = slice(b * Qh, (b + 1) * Qh)
            assert np.array_equal(ivf._last["heap_idx"].cpu().numpy(), ref_heap[sl])
            assert np.array_equal(ids.cpu().numpy(), ref[0][sl]) and np.array_equal(dst.cpu().numpy(), ref[2][sl])
            assert total // 16 > 1024 * Qh                              # streams long enough for the cm rounds
    finally:
        for b in bufs:
            b.close()


def test_pull_exchange_single_gpu(monkeypatch):
    """The pull exchange driven rank by rank on one GPU: every "rank" scans the lists it owns into its own buffer (estimates +
    chunk minima), the home side computes where its segments live in the owners' buffers, copies the minima into its compact
    layout and replays with absolute addresses == the unsharded path, heaps included; with and without minima; and the
    capacity guard leaves out exactly the segments that do not fit."""
    import torch
    from tinyknn_b200 import synth, ivf as ivf_mod, sharded as SH
    X = synth.clustered(300_000 + 48, 64, 40, seed=9)
    ivf = synth.build_ivf(X[:300_000], "euclidean", 12, seed=9)
    qs = X[300_000:].contiguous()
    G, k, n_probes = 3, 10, 6
    Qh = qs.shape[0] // G
    monkeypatch.setattr(ivf_mod, "CMIN_CHUNKS", 0)
    ref = ivf.query_batch(qs, k, n_probes=n_probes, order="device", return_distances=True, sub_batches=1)
    ref_heap = ivf._last["heap_idx"].cpu().numpy()
    monkeypatch.setattr(ivf_mod, "CMIN_CHUNKS", 1)
    shards = [SH.ShardedIVF(ivf, rank=r, world=G, drop_full_codes=False) for r in range(G)]
    homes = [sh._home(qs[r * Qh:(r + 1) * Qh].contiguous(), n_probes) for r, sh in enumerate(shards)]
    P = homes[0]["P"]
    tables = torch.cat([h["lut"]["tables"] for h in homes])
    probes = torch.cat([h["probes"] for h in homes])
    cap = G * shards[0].push_capacity(Qh, P)                       # an owner serves the queries of all ranks
    bufs = [SH.PeerBuffers(cap, rank=0, world=1, n_buf=1) for _ in range(G)]
    try:
        addr = np.array([b.local[0].address for b in bufs], dtype=np.int64)
        owner_base = D.upload(addr)
        cm_table = D.upload(addr + bufs[0].nbytes - (addr >> 4))
        for use_cm in (True, False):
            groups = torch.stack([sh._scan_pull(tables, probes, Qh, P, bufs[r].local[0], bufs[r].local_cmin[0] if use_cm else None, cap)
                                  for r, sh in enumerate(shards)])
            assert int(groups[:, 0].max()) <= cap
            for b, sh in enumerate(shards):
                seg_addr, seg_local, cmin = sh._pull_home(probes, homes[b], owner_base, groups, cm_table if use_cm else None)
                assert (cmin is not None) == use_cm
                ids, cnt, dst = sh._finish(homes[b], None, seg_addr, k, (n_probes + 1) * k + 1, cmin, seg_local)
                sl = slice(b * Qh, (b + 1) * Qh)
                assert np.array_equal(ivf._last["heap_idx"].cpu().numpy(), ref_heap[sl])
                assert np.array_equal(ids.cpu().numpy(), ref[0][sl]) and np.array_equal(dst.cpu().numpy(), ref[2][sl])
        # the guard: with room for half of rank 0's segments the plan drops the rest (offset -1) and still reports the full total
        full = int(groups[0, 0])
        g2 = shards[0]._scan_pull(tables, probes, Qh, P, bufs[0].local[0], None, full // 2)
        seg = shards[0].ivf._last["scan_seg_off"].cpu().numpy()
        sizes = np.asarray(ivf.to_device()["host_sizes"])
        pr = probes.cpu().numpy()
        nbytes = 16 * ((sizes[pr] + 15) // 16)
        kept = seg >= 0
        assert int(g2[0]) == full and kept.any() and not kept[np.asarray(shards[0].owner)[pr] == 0].all()
        assert (seg[kept] + nbytes[kept]).max() <= full // 2
    finally:
        for b in bufs:
            b.close()


def test_pull_minima_copies_every_alignment():
    """tkb_ivf_pull_minima_dev against numpy: segments of every length class (empty, shorter than a word, many 1 KB blocks) at
    every byte alignment of source and destination, skipped probes, two owners; the bytes around every destination stay."""
    from tinyknn_b200._lib import lib, check, PROBE_SKIP
    rng = np.random.default_rng(4)
    n_lists, Q, P = 40, 24, 7
    sizes = rng.choice([0, 1, 15, 16, 17, 40, 63, 700, 4097, 9000, 33000], size=n_lists).astype(np.int32)
    owner = rng.integers(0, 2, size=n_lists).astype(np.int32)
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
    probes[3, 2] = PROBE_SKIP
    nch = (sizes.astype(np.int64) + 15) // 16
    src = [rng.integers(0, 256, size=200_000, dtype=np.uint8) for _ in range(2)]       # the two owners' minima regions
    seg_addr = np.full((Q, P), -1, dtype=np.int64)
    seg_local = np.full((Q, P), -1, dtype=np.int64)
    run_local, expect = 5, np.full(400_000, 0xEE, dtype=np.uint8)
    for q in range(Q):
        for s_ in range(P):
            l = probes[q, s_]
            if l == PROBE_SKIP:
                continue
            a = int(rng.integers(0, 200_000 - nch[l] - 1))                              # any byte alignment of the source
            seg_addr[q, s_] = 16 * a                                                    # "estimate address": minima index = address >> 4
            run_local += int(rng.integers(0, 4))                                        # any byte alignment of the destination
            seg_local[q, s_] = 16 * run_local
            expect[run_local:run_local + nch[l]] = src[owner[l]][a:a + nch[l]]
            run_local += int(nch[l])
    d_src = [D.upload(x) for x in src]
    cm_table = D.upload(np.array([D.ptr(x) for x in d_src], dtype=np.int64))
    dst = D.upload(np.full(400_000, 0xEE, dtype=np.uint8))
    d_probes, d_sizes, d_owner, d_addr, d_local = (D.upload(x) for x in (probes, sizes, owner, seg_addr, seg_local))   # named: they outlive the launch
    check(lib.tkb_ivf_pull_minima_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), D.ptr(d_owner), n_lists,
                                      D.ptr(d_addr), D.ptr(d_local), D.ptr(cm_table), D.ptr(dst), D.stream_ptr()))
    assert np.array_equal(dst.cpu().numpy(), expect)


def test_saved_index_answers_identically(tmp_path):
    """save_index / load_index (memory-mapped) -> to_device -> query_batch == the original index."""
    np.random.seed(6)
    X = np.random.randn(4000, 32).astype(np.float32)
    qs = np.random.randn(64, 32).astype(np.float32)
    ivf = tinyknn.IVF("euclidean", 16, tinyknn.FastPQ(2)).fit(X[:2000]).build(X, n_probes=1)
    got = tinyknn.load_index(tinyknn.save_index(ivf, str(tmp_path / "idx")))
    a = ivf.query_batch(qs, 10, n_probes=4, order="device", return_distances=True)
    b = got.query_batch(qs, 10, n_probes=4, order="device", return_distances=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_glove_shape_full_size_properties():
    """BASELINE.json configs[1] at full size (1 183 514 x 100, 1087 lists): size-independent properties of the path --
    idempotence, sub-batching and the fused kernel give the same answers, ids == oracle (reference selection order) on a
    sample, and every estimate of a whole-database brute-force scan == the C oracle (61 M lookups per query)."""
    from tinyknn_b200 import synth
    n, nq = 1_183_514, 2048
    X = synth.clustered(n + nq, 100, 2000, seed=10)
    ivf = synth.build_ivf(X[:n], "angular", 1087, seed=10)
    qs = X[n:].contiguous()
    kw = dict(k=10, n_probes=10, order="device", return_distances=True)
    a = ivf.query_batch(qs, **kw)
    for other in (ivf.query_batch(qs, **kw), ivf.query_batch(qs, sub_batches=1, **kw), ivf.query_batch(qs, sub_batches=1, fused=True, **kw)):
        assert all(np.array_equal(x, y) for x, y in zip(a, other))
    assert (a[1] == 10).all() and (np.diff(a[2], axis=1) >= 0).all()          # k results each, ascending distances
    S = O.IVFState.from_ivf(ivf)
    K = O.Kernels("port", "avx")
    qh = qs.cpu().numpy()
    ids, cnt = ivf.query_batch(qh[:128], 10, n_probes=10, order="numpy")
    for i in range(128):
        assert set(ids[i][:cnt[i]]) == set(O.ivf_query(S, qh[i], 10, n_probes=10, kernels=K))
    # whole database, one list after the other, through FastPQ's own brute-force entry point
    packed = np.concatenate([td.packed for td in ivf.pq_transformed_points if isinstance(td, tuple)])
    td_all = tinyknn.fast_pq.TransformedData(16 * len(packed), packed)
    for i in range(2):
        q = qh[i] / np.linalg.norm(qh[i])
        dt = ivf.pq.distance_table(q)
        est = dt.estimate_distances(td_all)
        exp = np.zeros(2 * len(packed), np.uint64)
        O.estimate_pq(packed, np.ascontiguousarray(dt.tables), exp, True, "avx")
        assert np.array_equal(est.view(np.uint8), exp.view(np.uint8)[:len(est)])


@pytest.mark.parametrize("order", ["avx", "sse"])
def test_sift_shape_full_size_both_orders(order, monkeypatch):
    """BASELINE.json configs[2] at full size (1 000 000 x 128 euclidean, 1024 lists; d = 128 is rotated to 64: M = 32) in both
    accumulation orders (`_fast_pq_256.pyx` = avx, `_fast_pq.pyx` = sse): ids == oracle (the reference's selection order) on 96
    queries, the throughput mode is idempotent and independent of sub-batching, and in the avx order the tensor-core scan and
    the CUDA-core scan give the same heaps and ids for the whole batch."""
    from tinyknn_b200 import synth, fast_pq, ivf as ivf_mod
    n, nq = 1_000_000, 2048
    X = synth.clustered(n + nq, 128, 2000, seed=12)
    fast_pq.set_order(order)
    try:
        ivf = synth.build_ivf(X[:n], "euclidean", 1024, seed=12)
        qs = X[n:].contiguous()
        kw = dict(k=10, n_probes=10, order="device", return_distances=True)
        a = ivf.query_batch(qs, **kw)
        heaps = ivf._last["heap_idx"].cpu().numpy().copy()
        for other in (ivf.query_batch(qs, **kw), ivf.query_batch(qs, sub_batches=1, **kw)):
            assert all(np.array_equal(x, y) for x, y in zip(a, other))
        assert (a[1] == 10).all() and (np.diff(a[2], axis=1) >= 0).all()
        if order == "avx" and ivf_mod._tc_possible(ivf.to_device()):
            monkeypatch.setattr(ivf_mod, "TC_SCAN", "0")                  # the same batch through the CUDA-core kernels only
            b = ivf.query_batch(qs, sub_batches=1, **kw)
            assert np.array_equal(ivf._last["heap_idx"].cpu().numpy(), heaps)
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
            monkeypatch.setattr(ivf_mod, "TC_SCAN", "auto")
        S = O.IVFState.from_ivf(ivf)
        K = O.Kernels("port", order)
        qh = qs.cpu().numpy()
        ids, cnt = ivf.query_batch(qh[:96], 10, n_probes=10, order="numpy")
        for i in range(96):
            assert set(ids[i][:cnt[i]]) == set(O.ivf_query(S, qh[i], 10, n_probes=10, kernels=K))
    finally:
        fast_pq.set_order("avx")


def test_graphed_batch_equals_eager(golden):
    """IVF.graphed: the captured batch replays to the same ids / counts / distances as the eager query_batch, for
    several different query sets through the same graph, with and without sub-batches."""
    from tinyknn_b200 import synth
    X = synth.clustered(120_000 + 3 * 4096, 64, 80, seed=2)
    ivf = synth.build_ivf(X[:120_000], "euclidean", 128, seed=2)
    for sub in (1, 2):
        g = ivf.graphed(4096, 10, n_probes=8, sub_batches=sub)
        assert g.launches_per_replay > 5
        for i in range(3):
            qs = X[120_000 + i * 4096:120_000 + (i + 1) * 4096].contiguous()
            ref = ivf.query_batch(qs, 10, n_probes=8, order="device", return_distances=True, sub_batches=1)
            got = g(qs, return_distances=True)
            assert all(np.array_equal(a, b) for a, b in zip(ref, got)), (sub, i)
            got_h = g(qs.cpu().numpy(), return_distances=True)           # host queries in
            assert all(np.array_equal(a, b) for a, b in zip(ref, got_h))


def test_async_results_equal_sync(golden):
    """query_batch(to_host="async"): two batches in flight, collected out of band, equal the synchronous results."""
    from tinyknn_b200 import synth
    X = synth.clustered(100_000 + 2 * 6000, 64, 60, seed=4)
    ivf = synth.build_ivf(X[:100_000], "euclidean", 100, seed=4)
    import torch
    qa = X[100_000:106_000].cpu().pin_memory()
    qb = X[106_000:112_000].cpu().pin_memory()
    ref_a = ivf.query_batch(qa.numpy(), 10, n_probes=6, order="device", return_distances=True)
    ref_b = ivf.query_batch(qb.numpy(), 10, n_probes=6, order="device", return_distances=True)
    pa = ivf.query_batch(qa.numpy(), 10, n_probes=6, order="device", return_distances=True, to_host="async")
    pb = ivf.query_batch(qb.numpy(), 10, n_probes=6, order="device", return_distances=True, to_host="async")
    assert all(np.array_equal(x, y) for x, y in zip(ref_b, pb.result()))
    assert all(np.array_equal(x, y) for x, y in zip(ref_a, pa.result()))


from tinyknn_b200._lib import lib, check, DTYPE_F32, ORDER_AVX, ORDER_SSE           # noqa: E402
from tinyknn_b200._transform import transform_data                                  # noqa: E402
from test_gpu_parity import _ivf_from_state                                         # noqa: E402


# ---- probe selection as one kernel (tkb_coarse_probes_dev) -------------------------------------------------------------------

def _coarse_staged(cc, nck, C, M, tables, Q, centers, d, qn, Rc, P, order):
    st = D.stream_ptr()
    est = D.empty((Q, 16 * nck), np.uint8)
    ws = D.scan_workspace(Q * nck)
    check(lib.tkb_estimate_native_dev(D.ptr(cc), nck, M, D.ptr(tables), Q, D.ptr(est), 16 * nck, order, 1, D.ptr(ws), ws.numel(), st))
    hi, hv = D.empty((Q, Rc), np.int64), D.empty((Q, Rc), np.int32)
    check(lib.tkb_replay_fresh_dev(D.ptr(est), 16 * nck, nck, C, D.ptr(hi), D.ptr(hv), Q, Rc, 1, st))
    probes, dc = D.empty((Q, P), np.int32), D.empty((Q, Rc), np.float32)
    if Rc <= P:
        check(lib.tkb_select_probes_dev(D.ptr(hi), None, DTYPE_F32, Q, Rc, P, D.ptr(probes), st))
    else:
        check(lib.tkb_gather_dists_dev(D.ptr(centers), DTYPE_F32, C, d, D.ptr(qn), D.ptr(hi), Q, Rc, D.ptr(dc), st))
        check(lib.tkb_select_probes_dev(D.ptr(hi), D.ptr(dc), DTYPE_F32, Q, Rc, P, D.ptr(probes), st))
    return [x.cpu().numpy() for x in (probes, hi, hv, dc)] + [est.cpu().numpy()]


@pytest.mark.parametrize("M,order", [(52, "avx"), (32, "avx"), (8, "avx"), (6, "sse"), (52, "sse")])
def test_coarse_probes_one_kernel_equals_staged_and_oracle(M, order):
    """tkb_coarse_probes_dev == estimate + fresh replay + gather + select_probes, bit for bit (probes, heap arrays, distances),
    and its heap == the oracle's replay of the same estimates: random LUT-like, saturating ("hot": -1 slots survive) and
    full-range tables; centroid counts that end inside a chunk, fewer centroids than candidates, R <= P."""
    rng = np.random.default_rng(M + len(order))
    o = ORDER_AVX if order == "avx" else ORDER_SSE
    for trial, (C, P) in enumerate([(1087, 10), (1024, 10), (17, 3), (9, 2), (300, 40), (40, 30), (1, 1), (129, 1)]):
        d = (100, 16, 7)[trial % 3]
        Rc = min(2 * P + 10, C)
        Q = 6
        kind = ("lut", "hot", "full", "narrow")[trial % 4]
        codes = rng.integers(0, 16, size=(-(-C // 16) * 16, M), dtype=np.uint8)
        codes[C:] = 0                                               # padding rows are encoded zero vectors, whatever their code
        packed = transform_data(codes)
        nck = len(packed)
        tabs = []
        for _ in range(Q):
            if kind == "full":
                t = rng.integers(0, 256, size=(M, 16)).astype(np.uint8)
            elif kind == "hot":
                t = (rng.integers(8, 30, size=(M, 16)) - 20).astype(np.int8).view(np.uint8)
            elif kind == "narrow":
                t = (rng.integers(0, 28, size=(M, 16)) - 4).astype(np.int8).view(np.uint8)
            else:
                t = np.round(rng.exponential(6.0, size=(M, 16)) - 4).clip(-4, 128 / M ** 0.5).astype(np.int8).view(np.uint8)
            tabs.append(t)
        tables = D.upload(np.stack(tabs))
        cc = D.to_native(D.upload(packed), nck, M)
        cen = rng.standard_normal((C, d)).astype(np.float32)
        if C > 3:
            cen[2] = cen[1]                                         # tied distances: the slot decides
        qs = rng.standard_normal((Q, d)).astype(np.float32)
        centers, qn = D.upload(cen), D.upload(qs)
        exp = _coarse_staged(cc, nck, C, M, tables, Q, centers, d, qn, Rc, P, o)
        probes, hi, hv, dc = D.empty((Q, P), np.int32), D.empty((Q, Rc), np.int64), D.empty((Q, Rc), np.int32), D.empty((Q, Rc), np.float32)
        check(lib.tkb_coarse_probes_dev(D.ptr(cc), nck, C, M, D.ptr(tables), Q, D.ptr(centers), d, D.ptr(qn), Rc, P, o,
                                        D.ptr(probes), D.ptr(hi), D.ptr(hv), D.ptr(dc), D.stream_ptr()))
        got = [x.cpu().numpy() for x in (probes, hi, hv, dc)]
        assert np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2]), (trial, C, P, kind)
        assert np.array_equal(got[0], exp[0]), (trial, C, P, kind)
        if Rc > P:
            assert np.array_equal(got[3], exp[3]), (trial, C, P, kind)
        for q in range(Q):                                          # and the oracle, from the staged path's estimates
            oi, ov = np.zeros(Rc, np.int64), np.zeros(Rc, np.int32)
            O.init_heap(oi, ov, True)
            O.replay(exp[4][q], C, oi, ov, True)
            assert np.array_equal(got[1][q], oi) and np.array_equal(got[2][q], ov), (trial, q)
        # outputs are optional
        p2 = D.empty((Q, P), np.int32)
        check(lib.tkb_coarse_probes_dev(D.ptr(cc), nck, C, M, D.ptr(tables), Q, D.ptr(centers), d, D.ptr(qn), Rc, P, o,
                                        D.ptr(p2), None, None, None, D.stream_ptr()))
        assert np.array_equal(p2.cpu().numpy(), exp[0])


def test_query_batch_with_one_kernel_probe_selection(golden, monkeypatch):
    """IVF.query_batch(order="device") with TKB_COARSE_FUSED: same ids, counts, distances, probe heaps and final heaps."""
    from tinyknn_b200 import ivf as ivf_mod
    z = golden["ivf"]
    for name in z["names"]:
        ivf = _ivf_from_state(O.ivf_state_from_arrays(z, name + "_"))
        ivf._keep_heaps = True
        qs = z[name + "_q"]
        for npr in (1, 3, 8):
            monkeypatch.setattr(ivf_mod, "COARSE_FUSED", False)
            a = ivf.query_batch(qs, 10, n_probes=npr, order="device", return_distances=True, sub_batches=1)
            ha = ivf._last["center_heap"].cpu().numpy(), ivf._last["heap_idx"].cpu().numpy()
            monkeypatch.setattr(ivf_mod, "COARSE_FUSED", True)
            b = ivf.query_batch(qs, 10, n_probes=npr, order="device", return_distances=True, sub_batches=1)
            hb = ivf._last["center_heap"].cpu().numpy(), ivf._last["heap_idx"].cpu().numpy()
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), (name, npr)
            assert np.array_equal(ha[0], hb[0]) and np.array_equal(ha[1], hb[1]), (name, npr)


def test_coarse_probes_rejects_bad_arguments():
    with pytest.raises(ValueError):
        check(lib.tkb_coarse_probes_dev(None, 4, 60, 6, None, 1, None, 8, None, 30, 10, ORDER_AVX, None, None, None, None, None))
    with pytest.raises(ValueError):
        check(lib.tkb_coarse_probes_dev(None, 4, 20, 8, None, 1, None, 8, None, 30, 10, ORDER_AVX, None, None, None, None, None))   # R > C
    check(lib.tkb_coarse_probes_dev(None, 4, 60, 8, None, 0, None, 8, None, 30, 10, ORDER_AVX, None, None, None, None, None))      # Q = 0
