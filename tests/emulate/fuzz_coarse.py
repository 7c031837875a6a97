"""One-off fuzz on the emulator: tkb_coarse_probes_dev against the four staged kernels (probes, heap arrays, distances).
    python tests/emulate/fuzz_coarse.py <seed> <seconds>"""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, HERE)
import numpy as np
import emu_torch; emu_torch.install()
from tinyknn_b200 import _device as D
from tinyknn_b200._lib import lib, check, DTYPE_F32, ORDER_AVX, ORDER_SSE
from tinyknn_b200._transform import transform_data
seed0 = int(sys.argv[1]); budget = float(sys.argv[2]); t_end = time.time() + budget; trial = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 * 100000 + trial)
    order = ("avx", "sse")[int(rng.integers(0, 2))]
    M = int(rng.choice([4, 8, 32, 52, 12])) if order == "avx" else int(rng.choice([2, 6, 32, 52]))
    o = ORDER_AVX if order == "avx" else ORDER_SSE
    C = int(rng.choice([1, 2, 15, 16, 17, 40, 100, 257, 1087]))
    P = int(rng.integers(1, 50)); Rc = min(2 * P + 10, C); Q = int(rng.integers(1, 8)); d = int(rng.choice([4, 7, 16, 100]))
    codes = rng.integers(0, 16, size=(-(-C // 16) * 16, M), dtype=np.uint8); codes[C:] = 0
    packed = transform_data(codes); nck = len(packed)
    kind = int(rng.integers(0, 4)); tabs = []
    for _ in range(Q):
        if kind == 0: t = np.round(rng.exponential(6.0, size=(M, 16)) - 4).clip(-4, 128 / M ** 0.5).astype(np.int8).view(np.uint8)
        elif kind == 1: t = (rng.integers(8, 30, size=(M, 16)) - 20).astype(np.int8).view(np.uint8)
        elif kind == 2: t = rng.integers(0, 256, size=(M, 16)).astype(np.uint8)
        else: t = (rng.integers(0, 28, size=(M, 16)) - 4).astype(np.int8).view(np.uint8)
        tabs.append(t)
    tables = D.upload(np.stack(tabs)); cc = D.to_native(D.upload(packed), nck, M)
    cen = rng.standard_normal((C, d)).astype(np.float32)
    if C > 3 and rng.random() < 0.5: cen[2] = cen[1]
    centers, qn = D.upload(cen), D.upload(rng.standard_normal((Q, d)).astype(np.float32))
    st = D.stream_ptr()
    est, ws = D.empty((Q, 16 * nck), np.uint8), D.empty((64,), np.uint8)
    check(lib.tkb_estimate_native_dev(D.ptr(cc), nck, M, D.ptr(tables), Q, D.ptr(est), 16 * nck, o, 1, D.ptr(ws), 64, st))
    hi, hv = D.empty((Q, Rc), np.int64), D.empty((Q, Rc), np.int32)
    check(lib.tkb_replay_fresh_dev(D.ptr(est), 16 * nck, nck, C, D.ptr(hi), D.ptr(hv), Q, Rc, 1, st))
    pr, dc = D.empty((Q, P), np.int32), D.empty((Q, Rc), np.float32)
    if Rc <= P:
        check(lib.tkb_select_probes_dev(D.ptr(hi), None, DTYPE_F32, Q, Rc, P, D.ptr(pr), st))
    else:
        check(lib.tkb_gather_dists_dev(D.ptr(centers), DTYPE_F32, C, d, D.ptr(qn), D.ptr(hi), Q, Rc, D.ptr(dc), st))
        check(lib.tkb_select_probes_dev(D.ptr(hi), D.ptr(dc), DTYPE_F32, Q, Rc, P, D.ptr(pr), st))
    p2, h2, v2, d2 = D.empty((Q, P), np.int32), D.empty((Q, Rc), np.int64), D.empty((Q, Rc), np.int32), D.empty((Q, Rc), np.float32)
    check(lib.tkb_coarse_probes_dev(D.ptr(cc), nck, C, M, D.ptr(tables), Q, D.ptr(centers), d, D.ptr(qn), Rc, P, o, D.ptr(p2), D.ptr(h2), D.ptr(v2), D.ptr(d2), st))
    same = np.array_equal(pr.numpy(), p2.numpy()) and np.array_equal(hi.numpy(), h2.numpy()) and np.array_equal(hv.numpy(), v2.numpy())
    if Rc > P: same = same and np.array_equal(dc.numpy().view(np.uint32), d2.numpy().view(np.uint32))
    if not same:
        print("MISMATCH", seed0, trial, dict(order=order, M=M, C=C, P=P, Q=Q, d=d, kind=kind)); sys.exit(1)
    trial += 1
print("ok", trial, "trials, seed", seed0)
