"""The list-major tensor-core scan (csrc/tkb_scan_tc.cu, tcgen05.mma kind::i8) against the CUDA-core scan and the oracle:
same estimate bytes, same chunk minima, same query results (ref: tinyknn/_fast_pq_256.pyx:65-156, tinyknn/ivf.py:140-150)."""
import numpy as np
import pytest

import tinyknn_b200 as tinyknn
from tinyknn_b200 import _device as D
from tinyknn_b200 import ivf as ivf_mod
from tinyknn_b200._lib import PROBE_SKIP

pytestmark = pytest.mark.gpu


def _index(n, n_lists, seed=5, d=128):
    from tinyknn_b200 import synth
    X = synth.clustered(n + 512, d, max(8, n_lists // 2), seed)
    ivf = synth.build_ivf(X[:n], "euclidean", n_lists, seed=seed)
    return ivf, X[n:].contiguous()


def _scan(ivf, tables, probes, seg_off, est_bytes, tc):
    """One scan of the planned segments into zeroed buffers with / without the tensor-core kernel."""
    import torch
    dev = ivf.to_device()
    Q, P = probes.shape
    est = torch.zeros(est_bytes, dtype=torch.uint8, device=D.device())
    cmin = torch.zeros(est_bytes // 16 + 16, dtype=torch.uint8, device=D.device())
    old = ivf_mod.TC_SCAN
    ivf_mod.TC_SCAN = "1" if tc else "0"
    try:
        ivf._last = {}
        ivf._scan(dev, tables, probes, Q, P, est, seg_off, cmin=cmin)
        torch.cuda.synchronize()
    finally:
        ivf_mod.TC_SCAN = old
    refolded = None
    if tc:
        refolded = int(ivf._last["tc_ws"][:16].cpu().numpy().view(np.int32)[2])
    return est.cpu().numpy(), cmin.cpu().numpy(), refolded


def _tables(kind, Q, M, rng):
    if kind == "real":
        return None
    if kind == "hot":                               # large entries: many lane sums exceed the certificate's threshold
        t = rng.integers(-4, 24, size=(Q, M, 16))
    elif kind == "neg":                             # N_l > 128 for some queries: they must go to the CUDA-core kernel
        t = rng.integers(-4, 24, size=(Q, M, 16))
        t[::3] = rng.integers(-20, 8, size=t[::3].shape)
    else:                                           # full int8 range
        t = rng.integers(-128, 128, size=(Q, M, 16))
    return t.astype(np.int8).view(np.uint8)


@pytest.mark.parametrize("kind", ["real", "hot", "neg", "full"])
def test_tc_scan_bytes_equal_cuda_core_scan(kind):
    import torch
    rng = np.random.default_rng(11)
    ivf, qs = _index(60_000, 24)                    # ~2500 vectors per list: ~20 tiles; 300 queries x 6 probes over 24 lists
    dev = ivf.to_device()
    Q, P = 300, 6
    lut = ivf.pq.distance_tables(qs[:Q], signed=True)
    tables = lut["tables"]
    t = _tables(kind, Q, dev["M"], rng)
    if t is not None:
        tables = D.upload(t)
    ivf._last = {}
    probes = ivf._coarse(dev, lut, Q, P, min(2 * P + 10, dev["C"]), "device")
    pr = probes.cpu().numpy().copy()
    pr[5, 2] = PROBE_SKIP                           # a slot that does not exist
    pr[7, :] = pr[7, 0]                             # the same list in every slot of a query
    probes = D.upload(pr)
    seg_off, _ = ivf._plan(dev, probes, Q, P)
    so = seg_off.cpu().numpy().copy()
    so[9, 1] = -1                                   # a segment another rank would scan
    seg_off = D.upload(so)
    est_bytes = Q * P * 16 * max(dev["max_real_chunks"], 1)
    a_est, a_cm, _ = _scan(ivf, tables, probes, seg_off, est_bytes, tc=False)
    b_est, b_cm, refolded = _scan(ivf, tables, probes, seg_off, est_bytes, tc=True)
    assert a_est.any()
    assert np.array_equal(a_est, b_est), (kind, int((a_est != b_est).sum()), int(np.flatnonzero(a_est != b_est)[0]))
    assert np.array_equal(a_cm, b_cm), kind
    print("tc scan %s: %d bytes identical, %d pairs refolded" % (kind, int((a_est != 0).sum()), refolded))
    if kind == "hot":
        assert refolded > 0


def test_tc_scan_many_queries_per_list_and_long_lists():
    """More than 64 queries on one list (several groups) and lists longer than one work item (64 tiles = 8192 vectors)."""
    import torch
    ivf, qs = _index(100_000, 6, seed=9)            # ~16 000 vectors per list: ~130 tiles -> 3 items per (list, group)
    dev = ivf.to_device()
    Q, P = 400, 3                                   # ~200 queries per list: 4 groups
    lut = ivf.pq.distance_tables(qs[:Q], signed=True)
    ivf._last = {}
    probes = ivf._coarse(dev, lut, Q, P, min(2 * P + 10, dev["C"]), "device")
    seg_off, _ = ivf._plan(dev, probes, Q, P)
    est_bytes = Q * P * 16 * max(dev["max_real_chunks"], 1)
    a_est, a_cm, _ = _scan(ivf, lut["tables"], probes, seg_off, est_bytes, tc=False)
    b_est, b_cm, _ = _scan(ivf, lut["tables"], probes, seg_off, est_bytes, tc=True)
    assert np.array_equal(a_est, b_est) and np.array_equal(a_cm, b_cm)


def test_query_batch_with_tc_scan_equals_oracle():
    from oracle import restate as O
    ivf, qs = _index(40_000, 32, seed=4)
    qh = qs[:200].cpu().numpy()
    old = ivf_mod.TC_SCAN
    try:
        ivf_mod.TC_SCAN = "0"
        ref = ivf.query_batch(qh, 10, n_probes=8, order="device", return_distances=True)
        ivf_mod.TC_SCAN = "1"
        got = ivf.query_batch(qh, 10, n_probes=8, order="device", return_distances=True)
        ids, cnt = ivf.query_batch(qh, 10, n_probes=8, order="numpy")
    finally:
        ivf_mod.TC_SCAN = old
    assert all(np.array_equal(a, b) for a, b in zip(ref, got))
    S = O.IVFState.from_ivf(ivf)
    K = O.Kernels("port", "avx")
    for i in range(0, 200, 5):
        exp = O.ivf_query(S, qh[i], 10, n_probes=8, kernels=K)
        assert set(ids[i][:cnt[i]]) == set(exp), i
