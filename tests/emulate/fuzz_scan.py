"""One-off fuzz on the emulator: IVF scan (plain, chunk minima) on the native layout against the oracle's estimates."""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
import numpy as np
import emu_torch; emu_torch.install()
from tinyknn_b200 import _device as D
from tinyknn_b200._lib import lib, check, PROBE_SKIP, PLAN_SEND, ORDER_AVX, ORDER_SSE
from tinyknn_b200._transform import transform_data
from oracle import restate as O
seed0 = int(sys.argv[1]); budget = float(sys.argv[2]); t_end = time.time() + budget; trial = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 * 100000 + trial)
    order = ("avx", "sse")[int(rng.integers(0, 2))]
    M = int(rng.choice([4, 8, 32, 52, 12, 64])) if order == "avx" else int(rng.choice([2, 6, 32, 52, 10]))
    signd = 1
    n_lists = int(rng.integers(1, 12))
    sizes = rng.integers(0, 900, size=n_lists).astype(np.int32)
    sizes[rng.integers(0, n_lists)] = int(rng.choice([0, 1, 16, 17, 128, 129]))
    nc = -(-sizes.astype(np.int64) // 16); nc8 = -(-nc // 8) * 8
    off = np.zeros(n_lists + 1, np.int64); off[1:] = np.cumsum(nc8)
    tot = max(int(off[-1]), 8)
    codes = rng.integers(0, 16, size=(16 * tot, M), dtype=np.uint8)
    packed = transform_data(codes)
    Q = int(rng.integers(1, 12)); P = int(rng.integers(1, n_lists + 1))
    kind = int(rng.integers(0, 4))
    tabs = []
    for _ in range(Q):
        if kind == 0: t = np.round(rng.exponential(6.0, size=(M, 16)) - 4).clip(-4, 128 / M ** 0.5).astype(np.int8).view(np.uint8)
        elif kind == 1: t = (rng.integers(8, 30, size=(M, 16)) - 20).astype(np.int8).view(np.uint8)
        elif kind == 2: t = rng.integers(0, 256, size=(M, 16)).astype(np.uint8)
        else: t = (rng.integers(0, 28, size=(M, 16)) - 4).astype(np.int8).view(np.uint8)
        tabs.append(t)
    tabs = np.stack(tabs)
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(Q)]).astype(np.int32)
    if rng.random() < 0.3: probes[rng.integers(0, Q), rng.integers(0, P)] = PROBE_SKIP
    nat = D.to_native(D.upload(packed), tot, M)
    d_off, d_sizes, d_probes, d_tabs = (D.upload(x) for x in (off, sizes, probes, tabs))
    d_seg, d_gb, d_ws = D.empty((Q, P), np.int64), D.empty((3,), np.int64), D.empty((Q,), np.int64)
    check(lib.tkb_ivf_plan_dev(D.ptr(d_probes), Q, P, D.ptr(d_sizes), None, n_lists, PLAN_SEND, 0, 1, 0, D.ptr(d_seg), D.ptr(d_gb), D.ptr(d_ws), 8 * Q, D.stream_ptr()))
    seg = d_seg.cpu().numpy(); total = int(d_gb.cpu().numpy()[1])
    o = ORDER_AVX if order == "avx" else ORDER_SSE
    mq = int(P * max(nc.max(), 1))
    for use_cm in (False, True):
        est = D.empty((max(total, 16),), np.uint8); est.fill_(0xEE)
        cm = D.empty((max(total, 16) // 16 + 16,), np.uint8); cm.fill_(0xEE)
        ws = D.empty((64,), np.uint8)
        if use_cm:
            check(lib.tkb_ivf_scan_native_cm_dev(D.ptr(nat), D.ptr(d_off), D.ptr(d_sizes), n_lists, M, D.ptr(d_tabs), D.ptr(d_probes), Q, P, D.ptr(est), D.ptr(d_seg), D.ptr(cm), mq, o, signd, D.ptr(ws), 64, D.stream_ptr()))
        else:
            check(lib.tkb_ivf_scan_native_dev(D.ptr(nat), D.ptr(d_off), D.ptr(d_sizes), n_lists, M, D.ptr(d_tabs), D.ptr(d_probes), Q, P, D.ptr(est), 0, D.ptr(d_seg), mq, o, signd, D.ptr(ws), 64, D.stream_ptr()))
        e = est.cpu().numpy(); c = cm.cpu().numpy()
        for q in range(Q):
            for s in range(P):
                l = int(probes[q, s])
                if l == PROBE_SKIP or sizes[l] == 0: continue
                ncr = int(nc[l]); pk = np.ascontiguousarray(packed[off[l]:off[l] + ncr])
                exp = np.zeros(2 * ncr, np.uint64); O.estimate_pq(pk, O.transform_tables(tabs[q]), exp, True, order)
                got = e[seg[q, s]:seg[q, s] + 16 * ncr]
                if not np.array_equal(got, exp.view(np.uint8)):
                    print("MISMATCH est", seed0, trial, order, M, kind, q, s, use_cm); sys.exit(1)
                if use_cm:
                    em = exp.view(np.int8).reshape(-1, 16).min(1).view(np.uint8)
                    if not np.array_equal(c[seg[q, s] // 16:seg[q, s] // 16 + ncr], em):
                        print("MISMATCH cmin", seed0, trial, order, M, kind, q, s); sys.exit(1)
    trial += 1
print("ok", trial, "trials, seed", seed0)
