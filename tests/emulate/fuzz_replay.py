"""One-off fuzz on the emulator: IVF plan + replay (plain and chunk-minimum) against the oracle, adversarial shapes."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
import numpy as np
import emu_torch; emu_torch.install()
from tinyknn_b200 import _device as D
from tinyknn_b200._lib import lib, check, PROBE_SKIP, PLAN_SEND
from oracle import restate as O

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 300
t_end = time.time() + budget
trial = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 * 100000 + trial)
    signd = bool(rng.integers(0, 2))
    n_lists = int(rng.integers(1, 30))
    big = rng.random() < 0.3
    sizes = rng.integers(0, 6000 if big else 300, size=n_lists).astype(np.int32)
    if rng.random() < 0.5: sizes[rng.integers(0, n_lists)] = 0
    if rng.random() < 0.5: sizes[rng.integers(0, n_lists)] = int(rng.choice([1, 15, 16, 17, 31, 32, 33]))
    nc8 = (-(-sizes.astype(np.int64) // 128)) * 8
    off = np.zeros(n_lists + 1, np.int64); off[1:] = np.cumsum(nc8)
    tot = max(int(off[-1]), 1)
    ids = rng.permutation(16 * tot + 5).astype(np.int64)[:16 * tot] + 10 ** 11
    Q = int(rng.integers(1, 40)); P = int(rng.integers(1, min(n_lists, 16) + 1)); R = int(rng.choice([1, 2, 3, 7, 21, 64, 111, 255, 256, 300]))
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
    if rng.random() < 0.4: probes[rng.integers(0, Q), rng.integers(0, P)] = PROBE_SKIP
    d_off, d_sizes, d_ids, d_probes = (D.upload(x) for x in (off, sizes, ids, probes))
    d_seg, d_gb, d_ws = D.empty((Q, P), np.int64), D.empty((3,), np.int64), D.empty((Q,), np.int64)
    check(lib.tkb_ivf_plan_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), None, n_lists, PLAN_SEND, 0, 1, 0,
                               D.ptr(d_seg), D.ptr(d_gb), D.ptr(d_ws), 8 * Q, D.stream_ptr()))
    seg = d_seg.cpu().numpy(); total = int(d_gb.cpu().numpy()[1])
    kind = int(rng.integers(0, 4))
    pe = rng.integers(0, 256, size=max(total, 16), dtype=np.uint8)
    if kind == 1: pe = (pe // 16 + (60 if not signd else 0)).astype(np.uint8)          # ties
    if kind == 2: pe = np.sort(pe)[::-1].copy() if signd else np.sort(pe)[::-1].copy()  # descending: every vector a candidate
    if kind == 3: pe[:] = 5                                                                # constant
    # chunk minima as the scan would write them
    v = pe[:total // 16 * 16].reshape(-1, 16)
    cm = np.zeros(total // 16 + 16 + 16, np.uint8)
    if len(v): cm[:len(v)] = (v.view(np.int8).min(1).view(np.uint8) if signd else v.min(1))
    d_pe, d_cm = D.upload(pe), D.upload(cm)
    outs = []
    for use_cm in (False, True):
        hi, hv, fb = D.empty((Q, R), np.int64), D.empty((Q, R), np.int32), D.empty((Q,), np.int32)
        if use_cm:
            check(lib.tkb_ivf_replay_fresh_cm_dev(D.ptr(d_pe), D.ptr(d_seg), D.ptr(d_cm), D.ptr(d_off), D.ptr(d_sizes), n_lists, D.ptr(d_ids),
                                                  D.ptr(d_probes), Q, P, D.ptr(hi), D.ptr(hv), R, int(signd), 1, D.ptr(fb), D.stream_ptr()))
        else:
            check(lib.tkb_ivf_replay_fresh_dev(D.ptr(d_pe), 0, D.ptr(d_seg), D.ptr(d_off), D.ptr(d_sizes), n_lists, D.ptr(d_ids),
                                               D.ptr(d_probes), Q, P, D.ptr(hi), D.ptr(hv), R, int(signd), 1, D.ptr(fb), D.stream_ptr()))
        outs.append((hi.cpu().numpy(), hv.cpu().numpy()))
    for q in range(Q):
        oi, ov = np.zeros(R, np.int64), np.zeros(R, np.int32)
        O.init_heap(oi, ov, signd)
        for s in range(P):
            l = int(probes[q, s])
            if l == PROBE_SKIP or sizes[l] == 0: continue
            ncr = -(-int(sizes[l]) // 16)
            O.replay(pe[seg[q, s]:seg[q, s] + 16 * ncr], int(sizes[l]), oi, ov, signd, np.ascontiguousarray(ids[16 * off[l]:16 * off[l] + 16 * ncr]))
        for name, (a, b) in zip(("plain", "cm"), outs):
            if not (np.array_equal(a[q], oi) and np.array_equal(b[q], ov)):
                print("MISMATCH", name, "seed", seed0, "trial", trial, dict(signd=signd, n_lists=n_lists, Q=Q, P=P, R=R, kind=kind, q=q)); sys.exit(1)
    trial += 1
print("ok", trial, "trials, seed", seed0)
