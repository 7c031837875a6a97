"""The scan kernels replace the reference's saturating fold by a plain sum wherever a per-query CERTIFICATE proves that no prefix
of the fold can leave [-128, 127] (csrc/tkb_scan_core.cuh prepare_lut, csrc/tkb_scan_tc.cu tc_query_meta_kernel).
This file restates the round-2 thresholds in numpy and attacks them on the CPU: for random, realistic and adversarial tables, and
for random AND worst-case code vectors (the ones that maximise some prefix), every (vector, lane) the certificate accepts must fold
to its plain sum under the reference's recurrence (ref: tinyknn/_fast_pq_256.pyx:126-156: two lanes, rows (j >> 1) & 1, saturating
int8 add after every row). No GPU, no library: pure arithmetic."""
import numpy as np


def lane_rows(M, lane):
    return [j for j in range(M) if ((j >> 1) & 1) == lane]


def thresholds(T):
    """(eligible, k_0, k_1) for one signed table T (M, 16) int: the rule of prepare_lut / tc_query_meta_kernel."""
    M = T.shape[0]
    ks, ok = [], True
    for lane in (0, 1):
        rows = lane_rows(M, lane)
        neg = np.maximum(-T[rows].min(axis=1), 0)
        pmax = np.cumsum(np.maximum(T[rows].max(axis=1), 0))
        ok &= neg.sum() <= 128                                   # no prefix can go below -128
        over = np.nonzero(pmax > 127)[0]
        ks.append(127 - int(neg[over[0] + 1:].sum()) if len(over) else 127)
    return ok, ks[0], ks[1]


def fold(vals):
    """The reference's recurrence on one lane: vals (n, rows) int -> saturating int8 fold, and the plain sum."""
    a = np.zeros(len(vals), dtype=np.int64)
    for j in range(vals.shape[1]):
        a = np.clip(a + vals[:, j], -128, 127)
    return a, vals.sum(axis=1)


def adversarial_codes(T, rows):
    """Code vectors that maximise the prefix ending at every row k and then fall as far as possible (the vectors the bound is
    about), plus the all-max and all-min vectors."""
    Tr = T[rows]
    hi, lo = Tr.argmax(axis=1), Tr.argmin(axis=1)
    out = [hi.copy(), lo.copy()]
    for k in range(len(rows)):
        c = hi.copy()
        c[k + 1:] = lo[k + 1:]
        out.append(c)
        c2 = lo.copy()                                           # the opposite shape probes the lower side
        c2[k + 1:] = hi[k + 1:]
        out.append(c2)
    return np.array(out)


def tables(rng, kind, M):
    if kind == "realistic":                                      # what distance_table produces: row range ~25, minimum a few below 0
        base = rng.integers(-6, 2, size=(M, 1))
        return np.clip(base + rng.integers(0, 26, size=(M, 16)), -128, 127)
    if kind == "hot":                                            # large positive entries: sums saturate early
        return rng.integers(-4, 60, size=(M, 16))
    if kind == "negative":                                       # deep minima: N_l near or above 128
        return rng.integers(-20, 12, size=(M, 16))
    if kind == "spiky":                                          # one huge row among small ones
        T = rng.integers(-3, 6, size=(M, 16))
        T[rng.integers(M)] = rng.integers(-128, 128, size=16)
        return T
    return rng.integers(-128, 128, size=(M, 16))                 # full range


def test_certified_pairs_fold_to_their_plain_sum():
    rng = np.random.default_rng(2026)
    certified = rejected = 0
    for kind in ("realistic", "hot", "negative", "spiky", "full"):
        for M in (32, 52, 8, 20):
            for _ in range(40):
                T = tables(rng, kind, M)
                ok, k0, k1 = thresholds(T)
                if not ok:                                       # the kernels fold such queries step by step
                    continue
                for lane, k in ((0, k0), (1, k1)):
                    rows = lane_rows(M, lane)
                    codes = np.concatenate([rng.integers(0, 16, size=(300, len(rows))), adversarial_codes(T, rows)])
                    vals = T[rows][np.arange(len(rows))[None, :], codes]
                    f, s = fold(vals)
                    acc = s <= k
                    assert np.array_equal(f[acc], s[acc]), (kind, M, lane, k)
                    certified += int(acc.sum())
                    rejected += int((~acc).sum())
    assert certified > 50_000 and rejected > 1_000              # both sides of the test were exercised


def test_new_thresholds_never_below_the_round_one_bound_and_usually_far_above():
    """127 - N_l (all negatives, round 1) is the weakest form of the same argument: the round-2 threshold may only be larger."""
    rng = np.random.default_rng(7)
    gains = []
    for _ in range(300):
        T = tables(rng, "realistic", 32)
        ok, k0, k1 = thresholds(T)
        for lane, k in ((0, k0), (1, k1)):
            n = int(np.maximum(-T[lane_rows(32, lane)].min(axis=1), 0).sum())
            assert k >= 127 - n
            gains.append(k - (127 - n))
    assert np.mean(gains) > 5      # (these synthetic tables: ~12; the tables of the benchmark indexes: ~80 -> ~120)
