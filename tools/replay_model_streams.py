#!/usr/bin/env python3
"""Step 1 of the replay model (CPU only): build a small GloVe-like index on the host (same list length and probe count as
BASELINE.json configs[1]: ~1087 vectors per list, n_probes=10) and save, for 192 queries, the estimate stream the heap
replay consumes (oracle kernels). Output: /tmp/rq/streams.npy. Step 2: tools/replay_model.py."""
import os
os.makedirs("/tmp/rq", exist_ok=True)
import sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tinyknn_b200 as tinyknn
from oracle import restate as O
np.random.seed(10)
n, d, nl, nq = 200_000, 100, 184, 192
means = np.random.randn(2000, d) * 2
X = (means[np.random.randint(2000, size=n + nq)] + np.random.randn(n + nq, d)).astype(np.float32)
t0 = time.time()
ivf = tinyknn.IVF("angular", nl, tinyknn.FastPQ(2))
ivf.fit(X[:40000])
ivf.build(X[:n], n_probes=1, device=False)
print("built", time.time() - t0)
S = O.IVFState.from_ivf(ivf); K = O.Kernels("port", "avx")
P, k = 10, 10; R = (P + 1) * k + 1
streams = []
for q in X[n:]:
    q = q / np.linalg.norm(q)
    dt = O.make_dtable(S.pq, q, K)
    probes = dt.top(S.pq_transformed_centers, S.active_centers, k=P)
    parts = []
    for l in probes:
        nn, packed = S.pq_transformed_points[l]
        est = np.zeros(2 * len(packed), np.uint64)
        K.estimate_pq(packed, np.ascontiguousarray(dt.tables), est, True)
        e = est.view(np.int8).astype(np.int32)
        e[nn:] = 127          # padding never admitted
        parts.append(e)
    streams.append(np.concatenate(parts))
np.save("/tmp/rq/streams.npy", np.array(streams, dtype=object), allow_pickle=True)
print("streams", len(streams), np.mean([len(s) for s in streams]))
