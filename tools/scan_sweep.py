#!/usr/bin/env python3
"""BASELINE.json configs[3]: FastPQ brute-force scan throughput sweep -- 10M x 128 codes (16 B/vector after the 64-dim
rotation, M = 32), query batch 1..4096, single B200. One JSON line per batch size.

    python tools/scan_sweep.py [--n 10000000] [--batches 1,4,16,64,256,1024,4096] [--cpu-sample 1000000]

For each batch: `estimate_distances` for Q LUTs over the same codes (tkb_estimate_native_dev; every estimate is written,
1 B per (query, vector)) and the brute-force `top` path (estimates + exact heap replay, R = 30). Codes are uniform random
nibbles, LUTs come from real distance tables of Gaussian queries so that saturation is realistic.
The CPU line is the reference's own estimate_pq_avx (oracle/_ref) on one core over a bounded sample of the same codes.
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
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--batches", default="1,4,16,64,256,1024,4096")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    args = ap.parse_args()
    import torch
    import __graft_entry__
    __graft_entry__.build()
    import tinyknn_b200 as tinyknn                       # noqa: F401
    from tinyknn_b200 import synth, _device as D
    from tinyknn_b200._lib import lib, check, ORDER_AVX
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(10)
    d = 128
    sample = torch.randn(100_000, d, device=dev, generator=g)
    pq = synth.fit_pq(sample, seed=10)
    M = pq.centers.shape[1] // pq.dims_per_block
    n_chunks = (args.n + 15) // 16
    # uniform random nibbles, straight in the reference's packed layout: every byte = two independent codes
    packed = torch.randint(-2 ** 63, 2 ** 63 - 1, (n_chunks, M), device=dev, dtype=torch.int64, generator=g)
    nat = D.to_native(packed, n_chunks, M)
    code_bytes = n_chunks * M * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    batches = [int(b) for b in args.batches.split(",")]
    qmax = max(batches)
    queries = torch.randn(qmax, d, device=dev, generator=g)
    lut = pq.distance_tables(queries, signed=True)
    st = D.stream_ptr()
    for Q in batches:
        est = D.empty((Q, 16 * n_chunks), np.uint8)
        ws = D.scan_workspace(min(Q * n_chunks, 1 << 24))
        hi, hv = D.empty((Q, 30), np.int64), D.empty((Q, 30), np.int32)
        t_scan, t_top = [], []
        for rep in range(args.reps + 1):
            flush.zero_()                                # codes larger than L2 anyway; this also evicts the estimates
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            check(lib.tkb_estimate_native_dev(D.ptr(nat), n_chunks, M, D.ptr(lut["tables"]), Q, D.ptr(est), 16 * n_chunks,
                                              ORDER_AVX, 1, D.ptr(ws), ws.numel(), st))
            e1.record()
            check(lib.tkb_replay_fresh_dev(D.ptr(est), 16 * n_chunks, n_chunks, args.n, D.ptr(hi), D.ptr(hv), Q, 30, 1, st))
            e2.record()
            torch.cuda.synchronize()
            if rep:
                t_scan.append(e0.elapsed_time(e1)); t_top.append(e0.elapsed_time(e2))
        ms, ms_top = float(np.median(t_scan)), float(np.median(t_top))
        pairs = Q * 16 * n_chunks
        alg = pairs * (M // 2 + 1)                       # M/2 code bytes read + 1 estimate byte written per (query, vector)
        flagged = int(ws[:8].cpu().numpy().view(np.int64)[0])
        print(json.dumps(dict(
            metric="PQ-scan codes/s", workload="FastPQ brute-force scan, %d x 128 synthetic codes (M=%d, %d B/vector), batch %d"
            % (args.n, M, M // 2, Q), batch=Q, value=pairs / (ms * 1e-3), unit="codes/s", ms=ms,
            queries_per_s=Q / (ms * 1e-3), top_ms=ms_top, top_queries_per_s=Q / (ms_top * 1e-3),
            roofline=dict(bound="hbm", kernel="estimate_fast", achieved=alg / (ms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                          frac=alg / (ms * 1e-3) / 1e9 / peak, algorithmic_bytes_per_launch=alg,
                          physical_lower_bound_bytes=code_bytes + pairs, flagged_chunks=flagged,
                          note="algorithmic bytes count the codes once per query; with Q > 1 the CTAs of a stripe share them in L2"))))
        # the same estimates from the list-major tensor-core kernel (tkb_ivf_scan_tc_dev): the database as ONE inverted list that
        # every query probes; bytes compared with the CUDA-core kernel's
        if Q >= 16 and M == 32 and lib.tkb_ivf_scan_tc_supported():
            import ctypes
            tiles8 = -(-n_chunks // 8) * 8
            off = torch.tensor([0, tiles8], dtype=torch.int64, device=dev)
            size = torch.tensor([args.n], dtype=torch.int32, device=dev)
            probes = torch.zeros((Q, 1), dtype=torch.int32, device=dev)
            seg = (torch.arange(Q, dtype=torch.int64, device=dev) * (16 * n_chunks)).reshape(Q, 1).contiguous()
            est2 = D.empty((Q, 16 * n_chunks), np.uint8)
            need = ctypes.c_int64(0)
            check(lib.tkb_ivf_scan_tc_workspace(Q, 1, 1, ctypes.byref(need)))
            tws = D.empty((need.value,), np.uint8)
            t_tc = []
            for rep in range(args.reps + 1):
                flush.zero_()
                e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                e0.record()
                check(lib.tkb_ivf_scan_tc_dev(D.ptr(nat), D.ptr(off), D.ptr(size), 1, M, D.ptr(lut["tables"]), D.ptr(probes), Q, 1,
                                              D.ptr(est2), D.ptr(seg), None, None, 0, n_chunks, D.ptr(tws), tws.numel(), st))
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    t_tc.append(e0.elapsed_time(e1))
            ms_tc = float(np.median(t_tc))
            same = bool(torch.equal(est[:, :args.n], est2[:, :args.n]))
            print(json.dumps(dict(
                metric="PQ-scan codes/s", workload="FastPQ brute-force scan, %d x 128 synthetic codes (M=%d, %d B/vector), batch %d"
                % (args.n, M, M // 2, Q), batch=Q, kernel="ivf_scan_tc (tcgen05.mma kind::i8, list-major)", value=pairs / (ms_tc * 1e-3),
                unit="codes/s", ms=ms_tc, queries_per_s=Q / (ms_tc * 1e-3), speedup_over_cuda_core_scan=ms / ms_tc,
                estimates_identical_to_cuda_core_scan=same, refolded_pairs=int(tws[:16].cpu().numpy().view(np.int32)[2]))))
            del est2
        del est
    # CPU: the reference's kernel, one core, bounded sample
    try:
        from oracle import ref_loader, restate as O
        K = O.Kernels("ref" if ref_loader.have_ref_kernels() else "port", "avx")
        ns = min(args.cpu_sample, args.n) // 16 * 16
        sub = packed[:ns // 16].cpu().numpy().view(np.uint64)
        tab = np.ascontiguousarray(lut["tables"][0].cpu().numpy().reshape(-1).view(np.uint64))
        out = np.zeros(2 * (ns // 16), np.uint64)
        K.estimate_pq(sub, tab, out, True)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            K.estimate_pq(sub, tab, out, True)
        dt = (time.perf_counter() - t0) / reps
        got = D.empty((1, ns), np.uint8)
        ws = D.scan_workspace(ns // 16)
        check(lib.tkb_estimate_native_dev(D.ptr(D.to_native(packed[:ns // 16].contiguous(), ns // 16, M)), ns // 16, M,
                                          D.ptr(lut["tables"]), 1, D.ptr(got), ns, ORDER_AVX, 1, D.ptr(ws), ws.numel(), st))
        same = bool(np.array_equal(got.cpu().numpy().reshape(-1), out.view(np.uint8)))
        print(json.dumps(dict(impl="reference", metric="PQ-scan codes/s", value=ns / dt, unit="codes/s", cores=1,
                              kind="reference" if ref_loader.have_ref_kernels() else "port",
                              sample="%d codes x 1 query, %d repetitions, estimate_pq_avx" % (ns, reps),
                              gpu_estimates_bit_exact_on_sample=same)))
    except Exception as e:                               # noqa: BLE001
        print(json.dumps(dict(impl="reference", error=str(e))))


if __name__ == "__main__":
    main()
