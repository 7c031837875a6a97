"""GPU tests of code written AFTER round 1's GPU budget was spent: none of this has run on hardware yet, so the file is
skipped unless TKB_RUN_UNVALIDATED=1 (first thing to run next round: `TKB_RUN_UNVALIDATED=1 pytest tests/test_unvalidated_gpu.py`).
Covers tkb_assign_dev (IVF.build's coarse assignment) and the chunk minima inside the push exchange."""
import os

import numpy as np
import pytest

import tinyknn_b200 as tinyknn
from oracle import restate as O
from tinyknn_b200 import _device as D

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TKB_RUN_UNVALIDATED") != "1", reason="not yet validated on a GPU (TKB_RUN_UNVALIDATED=1 runs it)")]


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
    ids, cnt = ivf.query_batch(qh[:32], 10, n_probes=10, order="numpy")
    for i in range(32):
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
