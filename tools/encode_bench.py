#!/usr/bin/env python3
"""Throughput of the PQ encoder (tkb_encode_dev) on synthetic rows resident in HBM: rotated 128-d (f64 rotation, 16.4 kflop
per vector) and unrotated 100-d. Prints one JSON line per case. Usage: python tools/encode_bench.py [n_vectors]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import tinyknn_b200 as tinyknn  # noqa: E402
from tinyknn_b200 import _lib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    rng = np.random.default_rng(0)
    for name, d in (("rot128", 128), ("plain100", 100)):
        fit = rng.standard_normal((20_000, d)).astype(np.float32)
        pq = tinyknn.FastPQ(2, use_kmeans=False).fit(fit) if d != 100 else tinyknn.FastPQ(2).fit(fit[:4000])
        X = torch.randn(n, d, device="cuda", dtype=torch.float32)
        for _ in range(2):
            out = pq.encode_device(X)
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            out = pq.encode_device(X)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        M = out.shape[1]
        Dp = pq.centers.shape[1]
        flops = (2.0 * Dp * X.shape[1] if pq.R is not None else 0.0) + M * 16 * 8.0
        print(json.dumps(dict(case=name, n=n, d=d, M=M, ms=ms, vectors_per_s=n / (ms * 1e-3),
                              read_GBps=n * d * 4 / (ms * 1e-3) / 1e9, tflops=(n * flops) / (ms * 1e-3) / 1e12,
                              rotation="f64" if pq.R is not None else None, launches=(_lib.launch_count() - l0) // reps)))
        del X, out


if __name__ == "__main__":
    main()
