#!/usr/bin/env python3
"""BASELINE.json configs[0] (ref: examples/example.py:24-27,60-67; README.md:71-86): FastPQ(2) on X = randn(16000, 128), 1000
queries; per query a distance table and the estimates of the whole database. Prints one JSON line: the reference's protocol through
the single-query API (one C-ABI call per step, host arrays in and out), the batched device path for the same work, and the
reference's own kernels on one host core.

    python tools/c1_example.py [--n 16000] [--queries 1000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16000)
    ap.add_argument("--queries", type=int, default=1000)
    args = ap.parse_args()
    import torch
    import tinyknn_b200 as tinyknn
    from tinyknn_b200 import _device as D
    from tinyknn_b200._lib import lib, check, ORDER_AVX
    np.random.seed(10)
    X = np.random.randn(args.n, 128).astype(np.float32)
    qs = np.random.randn(args.queries, 128).astype(np.float32)
    pq = tinyknn.FastPQ(2)
    t0 = time.perf_counter()
    td = pq.fit_transform(X)
    fit_s = time.perf_counter() - t0
    # --- the reference's loop (example.py:60-67): distance_table(q) then estimate_distances(data) per query
    for q in qs[:20]:
        pq.distance_table(q).estimate_distances(td)
    torch.cuda.synchronize()
    t_lut = t_scan = 0.0
    ests = []
    for q in qs:
        t0 = time.perf_counter()
        dt = pq.distance_table(q)
        t1 = time.perf_counter()
        est = dt.estimate_distances(td)
        t2 = time.perf_counter()
        t_lut += t1 - t0; t_scan += t2 - t1
        ests.append(est.copy())
    # --- the same work batched on the device
    Q = len(qs)
    nat = D.mirror_native(td.packed)
    n_chunks, M = td.packed.shape
    qd = D.upload(qs)
    est_d = D.empty((Q, 16 * n_chunks), np.uint8)
    ws = D.scan_workspace(0)
    times = []
    for rep in range(6):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        lut = pq.distance_tables(qd, signed=True)
        e1.record()
        check(lib.tkb_estimate_native_dev(D.ptr(nat), n_chunks, M, D.ptr(lut["tables"]), Q, D.ptr(est_d), 16 * n_chunks, ORDER_AVX, 1,
                                          D.ptr(ws), ws.numel(), D.stream_ptr()))
        e2.record()
        torch.cuda.synchronize()
        if rep:
            times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    lut_ms, scan_ms = (float(np.median([t[i] for t in times])) for i in range(2))
    same = all(np.array_equal(est_d[i, :args.n].cpu().numpy().view(np.int8), ests[i]) for i in range(0, Q, 37))
    # --- the reference's kernels on one host core (oracle/_ref), same loop
    ref = None
    try:
        from oracle import restate as O, ref_loader
        K = O.Kernels("ref" if ref_loader.have_ref_kernels() else "port", "avx")
        P = O.PQState.from_pq(pq)
        tdo = O.TransformedData(td.size, td.packed)
        a = b = 0.0
        bad = 0
        for i, q in enumerate(qs):
            t0 = time.perf_counter()
            dto = O.make_dtable(P, q, K)
            t1 = time.perf_counter()
            eo = dto.estimate_distances(tdo)
            t2 = time.perf_counter()
            a += t1 - t0; b += t2 - t1
            bad += int(not np.array_equal(eo, ests[i]))
        ref = dict(lut_us_per_query=1e6 * a / Q, scan_us_per_query=1e6 * b / Q, queries_per_s=Q / (a + b), cores=1,
                   kind="reference" if ref_loader.have_ref_kernels() else "port", estimates_differing_from_gpu=bad)
    except Exception as e:                                   # noqa: BLE001
        ref = dict(error=str(e))
    print(json.dumps(dict(
        config="examples/example.py FastPQ exact-PQ scan: n=%d, d=128, %d queries, dims_per_block=2" % (args.n, Q),
        fit_transform_s=fit_s,
        single_query_api=dict(lut_us_per_query=1e6 * t_lut / Q, scan_us_per_query=1e6 * t_scan / Q, queries_per_s=Q / (t_lut + t_scan),
                              note="the reference's loop through the drop-in API: two C-ABI calls per query, host arrays in and out, "
                                   "a device synchronisation per call"),
        batched_device=dict(lut_us_per_query=1e3 * lut_ms / Q, scan_us_per_query=1e3 * scan_ms / Q,
                            queries_per_s=Q / ((lut_ms + scan_ms) * 1e-3), estimates_identical_to_single_query_api=bool(same)),
        reference_cpu=ref, published="README.md:71-86: 7101 q/s (90.3 us LUT + 50.6 us scan per query), hardware unstated")))


if __name__ == "__main__":
    main()
