#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the reference package from /root/reference with its Cython kernels compiled into
oracle/_ref (oracle/build_ref.py), runs it on seeded inputs and stores inputs + outputs. The
fixtures pin, beyond what the reference's own tests pin: LUT bytes, heap arrays after query_pq,
probe lists and final ids of IVF.query. Every array in the files was produced by reference code.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_loader, restate as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")


def scan_cases(t):
    """estimate_pq / query_pq of both kernel modules on random codes and tables."""
    sse, avx = t._fast_pq, sys.modules["tinyknn._fast_pq_avx"]
    rng = np.random.default_rng(1234)
    out = {}
    cases = []
    for ci, (order, signd, M, n, R, tabkind, with_labels) in enumerate([
            ("sse", True, 2, 16, 3, "full", False), ("sse", False, 6, 45, 7, "full", True),
            ("sse", True, 14, 200, 30, "narrow", True), ("sse", False, 32, 333, 111, "narrow", False),
            ("avx", True, 4, 16, 1, "full", False), ("avx", False, 8, 100, 10, "full", True),
            ("avx", True, 32, 1000, 30, "narrow", True), ("avx", False, 32, 777, 64, "narrow", True),
            ("avx", True, 52, 640, 111, "lut", True), ("avx", True, 32, 2048, 200, "lut", False)]):
        n16 = -(-n // 16) * 16
        codes = rng.integers(0, 16, size=(n16, M), dtype=np.uint8)
        if tabkind == "full":
            tab = rng.integers(0, 256, size=(M, 16)).astype(np.uint8)
        elif tabkind == "narrow":
            tab = rng.integers(0, 28, size=(M, 16)).astype(np.int16)
            tab = (tab - 4).astype(np.int8).view(np.uint8) if signd else tab.astype(np.uint8)
        else:   # a LUT shaped like the real thing: small negatives, max ~ 128/sqrt(M)
            tab = np.round(rng.exponential(6.0, size=(M, 16)) - 4).clip(-4, 128 / M ** 0.5).astype(np.int8).view(np.uint8)
        packed = t._transform.transform_data(codes)
        T = t._transform.transform_tables(tab)
        mod = avx if order == "avx" else sse
        est = np.zeros(2 * len(packed), dtype=np.uint64)
        getattr(mod, "estimate_pq_" + order)(packed, T, est, signd)
        # 64-bit labels only through the avx module (the sse module truncates them under Cython 3)
        labels = None
        if with_labels:
            labels = rng.permutation(10 ** 6)[:n16].astype(np.int64) + (10 ** 12 if order == "avx" else 0)
            labels[n16 // 2:n16 // 2 + 8] = labels[:8]           # duplicates exercise the dedupe
        hi, hv = np.zeros(R, np.int64), np.zeros(R, np.int32)
        mod.init_heap(hi, hv, signd)
        snaps = []
        for rep in range(2):                                     # heap state persists across calls
            getattr(mod, "query_pq_" + order)(packed, n, T, hi, hv, signd, labels)
            snaps.append((hi.copy(), hv.copy()))
        p = "c%d_" % ci
        out.update({p + "codes": codes, p + "tab": tab, p + "packed": packed, p + "tables": T, p + "est": est,
                    p + "labels": np.zeros(0, np.int64) if labels is None else labels,
                    p + "heap_idx": np.stack([s[0] for s in snaps]), p + "heap_val": np.stack([s[1] for s in snaps])})
        cases.append((order, signd, M, n, R, with_labels))
    out["cases"] = np.array([(o == "avx", s, M, n, R, l) for o, s, M, n, R, l in cases], dtype=np.int64)
    return out


def lut_cases(t):
    """FastPQ.distance_table / udistance_table on fitted quantizers (rotated f64 path, f32 path)."""
    out = {}
    np.random.seed(10)
    for name, (n, d, dpb) in {"d128": (2000, 128, 2), "d100": (1500, 100, 2), "d10": (300, 10, 1), "d24": (400, 24, 4)}.items():
        X = (np.random.randn(n, d) * (1 + np.arange(d) % 3)).astype(np.float32)
        pq = t.FastPQ(dpb)
        pq.fit(X)
        qs = (np.random.randn(64, d) * 1.5).astype(np.float32)
        st, ut, qrot, shift, scale = [], [], [], [], []
        for q in qs:
            dt = pq.distance_table(q)
            st.append(dt.tables.copy()); qrot.append(np.asarray(dt.q, dtype=np.float64)); shift.append(float(dt.mean)); scale.append(float(dt.scale))
            ut.append(pq.udistance_table(q).tables.copy())
        out.update({name + "_centers": pq.centers, name + "_R": np.zeros((0, 0)) if pq.R is None else pq.R,
                    name + "_sqrt": np.array(pq.sqrt_n_blocks), name + "_dpb": np.array(dpb), name + "_q": qs,
                    name + "_tables": np.stack(st), name + "_utables": np.stack(ut), name + "_qrot": np.stack(qrot),
                    name + "_shift": np.array(shift), name + "_scale": np.array(scale)})
    out["names"] = np.array(["d128", "d100", "d10", "d24"])
    return out


def ivf_cases(t):
    """IVF.query end to end on small clustered indexes, both metrics, a probe sweep."""
    out = {}
    names = []
    np.random.seed(10)
    for name, (metric, n, d, ncl, bp) in {"euc": ("euclidean", 3000, 32, 40, 1), "ang": ("angular", 2500, 20, 36, 2),
                                           "euc128": ("euclidean", 1200, 128, 24, 1)}.items():
        means = np.random.randn(25, d) * 2
        X = (means[np.random.randint(25, size=n)] + np.random.randn(n, d)).astype(np.float32)
        qs = (means[np.random.randint(25, size=48)] + np.random.randn(48, d)).astype(np.float32)
        ivf = t.IVF(metric, ncl, t.FastPQ(2))
        ivf.fit(X).build(X, n_probes=bp)
        S = O.IVFState.from_ivf(ivf)
        out.update(O.ivf_state_to_arrays(S, prefix=name + "_"))
        out[name + "_q"] = qs
        for npr in (1, 3, 8):
            res = np.full((len(qs), 10), -1, dtype=np.int64)
            for i, q in enumerate(qs):
                r = ivf.query(q.copy(), 10, n_probes=npr)
                res[i, :len(r)] = r
            out["%s_res_p%d" % (name, npr)] = res
        names.append(name)
    out["names"] = np.array(names)
    return out


def encode_cases(t):
    """FastPQ.transform of the reference (fast_pq.py:147-184): inputs, fitted quantizer, packed codes."""
    out, names = {}, []
    np.random.seed(10)
    for name, (n, d, dtype, kmeans) in {"rot128": (530, 128, np.float32, True), "plain100": (500, 100, np.float32, True),
                                        "rot20": (333, 20, np.float32, True), "rot128_f64": (200, 128, np.float64, True),
                                        "plain100_f64": (150, 100, np.float64, True), "rot200": (180, 200, np.float32, True)}.items():
        means = np.random.randn(12, d) * 2
        X = (means[np.random.randint(12, size=n)] + np.random.randn(n, d)).astype(dtype)
        pq = t.FastPQ(2, use_kmeans=kmeans).fit(X)
        td = pq.transform(X)
        assert td.size == n
        out[name + "_X"], out[name + "_centers"], out[name + "_packed"] = X, pq.centers, td.packed
        out[name + "_R"] = np.zeros((0, 0)) if pq.R is None else pq.R
        names.append(name)
    out["names"] = np.array(names)
    return out


def main():
    assert build_ref.build(), "needs /root/reference"
    t = ref_loader.load_ref_package()
    only = sys.argv[1:]
    for fname, fn in (("scan.npz", scan_cases), ("lut.npz", lut_cases), ("ivf.npz", ivf_cases), ("encode.npz", encode_cases)):
        if only and fname not in only:
            continue
        arrays = fn(t)
        np.savez_compressed(os.path.join(HERE, fname), **arrays)
        print(fname, os.path.getsize(os.path.join(HERE, fname)) // 1024, "KiB")


if __name__ == "__main__":
    main()
