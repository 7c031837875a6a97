"""FastPQ / _FastDistanceTable with the reference's API (ref: tinyknn/fast_pq.py), query side on B200.

What runs where:
  * fit / transform (build time, out of the hot-path scope): host numpy + sklearn/scipy, same
    algorithm and RNG call order as the reference so a seeded fit gives the same quantizer.
  * distance_table / udistance_table (ref: fast_pq.py:186-252): `tkb_lut_build_dev` -- the LUT is built
    on the GPU and stays there; the host copy in `.tables` is what the reference returns.
  * estimate_distances (ref: fast_pq.py:270-282): `tkb_estimate_dev` over a device mirror of `packed`.
  * top (ref: fast_pq.py:284-312): scan + exact heap replay on the device, exact rescoring distances with
    `tkb_gather_dists_dev`; only the reference's own `bottom_k` (np.argpartition over <= rescore floats)
    runs on the host, because its output ORDER is numpy-defined ("selection order").
"""
import os
import warnings
from collections import namedtuple

import numpy as np

from . import _device as D
from . import _kernels
from ._lib import lib, check, ORDER_AVX, ORDER_SSE, DTYPE_F32, DTYPE_F64
from ._transform import transform_data
from .utils import pad1, pad2, bottom_k

# ref: fast_pq.py:21-27 -- `avx` picks the accumulation order of the scan and the dimension padding.
avx = True
dpad = 4
from ._fast_pq_avx import query_pq_avx as query_pq, estimate_pq_avx as estimate_pq  # noqa: E402
from ._fast_pq_avx import init_heap  # noqa: E402,F401


def set_order(order):
    """Switch between the reference's two builds: "avx" (default, `avx = True`) and "sse"."""
    global avx, dpad, query_pq, estimate_pq
    from . import _fast_pq, _fast_pq_avx
    if order == "avx":
        avx, dpad = True, 4
        query_pq, estimate_pq = _fast_pq_avx.query_pq_avx, _fast_pq_avx.estimate_pq_avx
    elif order == "sse":
        avx, dpad = False, 2
        query_pq, estimate_pq = _fast_pq.query_pq_sse, _fast_pq.estimate_pq_sse
    else:
        raise ValueError("order must be 'avx' or 'sse'")


def _order():
    return ORDER_AVX if avx else ORDER_SSE


# "fast": PRMT/register-LUT kernel on the device-native layout (default).
# "generic": the step-by-step kernel on the reference layout (kept for A/B parity checks).
SCAN_IMPL = "fast"


TransformedData = namedtuple("TransformedData", "size packed")

_GAUSS_CODE = np.array(
    [(0.0, 0.0)]
    + [(r * np.cos(t), r * np.sin(t))
       for r, m in ((1, 6), (2, 9))
       for t in np.linspace(0, 2 * np.pi, m, endpoint=False)])


# `fit` on the GPU by default? (TKB_FIT_DEVICE=1; off: the reference's sklearn path, so that the reference's own tests see the
# reference's own fit)
FIT_DEVICE = os.environ.get("TKB_FIT_DEVICE", "0") != "0"


def kmeans_pp_init(X, k, rng):
    """Greedy k-means++ seeding (Arthur & Vassilvitskii; 2 + log k candidates per step, the one that lowers the potential most
    is kept -- the variant sklearn uses) of k centres from the rows of X, numpy Generator `rng`. Rows may repeat when X has fewer
    than k distinct ones."""
    n = X.shape[0]
    X64 = X.astype(np.float64)
    xn = np.einsum("ij,ij->i", X64, X64)
    out = np.empty((k, X.shape[1]), dtype=X.dtype)
    first = int(rng.integers(n))
    out[0] = X[first]
    d2 = np.maximum(xn + xn[first] - 2.0 * X64 @ X64[first], 0.0)
    trials = 2 + int(np.log(k))
    for c in range(1, k):
        tot = d2.sum()
        if not tot > 0:
            cand = rng.integers(n, size=1)
        else:
            cand = np.minimum(np.searchsorted(np.cumsum(d2), rng.random(trials) * tot), n - 1)
        dc = np.maximum(xn[None, :] + xn[cand][:, None] - 2.0 * X64[cand] @ X64.T, 0.0)      # (trials, n)
        pot = np.minimum(d2[None, :], dc)
        best = int(np.argmin(pot.sum(axis=1)))
        out[c] = X[cand[best]]
        d2 = pot[best]
    return out


class FastPQ:
    def __init__(self, dims_per_block, use_kmeans=True, rotate_dim=64):
        self.dims_per_block = dims_per_block
        self.centers = None          # f32 (16, padded_or_rotated_dim): centre c of block m at [c, m*dpb:(m+1)*dpb]
        self.sqrt_n_blocks = None
        self.use_kmeans = use_kmeans
        self.rotate_dim = rotate_dim
        self.R = None                # f64 (min(rotate_dim, d_pad), d_pad) random orthonormal rows, or None

    # ------------------------------------------------------------------ build time (host) ----
    def fit(self, data, verbose=False, device=None, seed=0):
        """ref: fast_pq.py:50-104. device=True: the per-block k-means runs on the GPU (`tkb_kmeans_pq_dev`, all blocks in one
        pass over the rows per iteration; k-means++ seeding on a host subsample with numpy's Generator(seed)); None: the module
        default FIT_DEVICE (TKB_FIT_DEVICE=1), else sklearn on the host like the reference. Parity unpinned either way: the
        reference's own fit is random."""
        assert data.size > 0, "Can't fit no data"
        true_d = data.shape[1]
        dpb = self.dims_per_block
        data = pad2(data, 16, dpad * dpb)
        d = data.shape[1]
        if self.rotate_dim is not None and true_d != 100:        # ref: fast_pq.py:77-82 (GloVe-100 exception)
            from scipy.stats import ortho_group
            self.R = ortho_group.rvs(dim=d)
            if d > self.rotate_dim:
                d = self.rotate_dim
                self.R = self.R[:d]
            data = data @ self.R.T
        if (FIT_DEVICE if device is None else device) and self.use_kmeans:
            books = self._fit_code_device(data, seed)
        else:
            books = self._fit_code(data, verbose=verbose)        # list of M arrays (16, dpb)
        self.centers = np.array(books, dtype=np.float32).transpose(1, 0, 2).reshape(16, d)
        self.sqrt_n_blocks = np.sqrt(d // dpb)
        return self

    def fit_transform(self, data, verbose=False):
        return self.fit(data, verbose).transform(data, verbose)

    def _fit_code_device(self, data, seed=0, iters=25):
        """The M codebooks of 16 centres by Lloyd's algorithm on the GPU (ref: fast_pq.py:106-145 fits one sklearn KMeans per
        block). Returns the list of M arrays (16, dpb) `_fit_code` returns."""
        X = np.ascontiguousarray(data, dtype=np.float32)
        n, d = X.shape
        dpb = self.dims_per_block
        rng = np.random.default_rng(seed)
        sub = X[rng.choice(n, min(n, 4096), replace=False)]
        centers = np.empty((16, d), dtype=np.float32)
        for m in range(d // dpb):
            centers[:, m * dpb:(m + 1) * dpb] = kmeans_pp_init(sub[:, m * dpb:(m + 1) * dpb], 16, rng)
        Xd, Cd = D.upload(X), D.upload(centers)
        ws = D.empty((16384 + 160 * d,), np.uint8)
        check(lib.tkb_kmeans_pq_dev(D.ptr(Xd), n, d, dpb, D.ptr(Cd), iters, float(np.abs(X).max()), D.ptr(ws), ws.numel(),
                                    D.stream_ptr()))
        centers = Cd.cpu().numpy()
        return [centers[:, m * dpb:(m + 1) * dpb].copy() for m in range(d // dpb)]

    def _fit_code(self, data, verbose=False):
        n, d = data.shape
        dpb = self.dims_per_block
        blocks = data.reshape(n, d // dpb, dpb).transpose(1, 0, 2)
        if not self.use_kmeans:
            assert dpb == 2, "Fixed code only defined for dpb = 2"
            return [_GAUSS_CODE @ np.linalg.cholesky(np.cov(b.T, bias=True)).T + b.mean(axis=0) for b in blocks]
        import sklearn.cluster
        from sklearn.exceptions import ConvergenceWarning
        km = sklearn.cluster.KMeans(16, n_init=2)
        it = blocks
        if verbose:
            import tqdm
            it = tqdm.tqdm(blocks)
        books = []
        for b in it:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", category=ConvergenceWarning)
                km.fit(b)
            books.append(km.cluster_centers_.copy())
        return books

    def transform(self, data, verbose=False, device=None):
        """ref: fast_pq.py:147-184. Build-time, not on the query path. device=None: encode on the GPU
        (`tkb_encode_dev`) when one is present, else with numpy on the host; True / False force one of them."""
        assert self.centers is not None, "PQ has not been fitted"
        if data.size == 0:
            return data
        if device is None:
            device = D.torch().cuda.is_available()
        if device:
            true_n = data.shape[0]
            packed = self.encode_device(D.upload(np.ascontiguousarray(data)) if isinstance(data, np.ndarray) else data)
            return TransformedData(true_n, packed.cpu().numpy().view(np.uint64))
        return self._transform_host(data)

    def _transform_host(self, data):
        true_n = data.shape[0]
        dpb = self.dims_per_block
        data = pad2(data, 16, dpad * dpb)
        if self.R is not None:
            data = data @ self.R.T
        n, d = data.shape
        M = d // dpb
        books = self.centers.reshape(16, M, dpb).transpose(1, 0, 2)          # (M, 16, dpb)
        codes = np.empty((n, M), dtype=np.uint8)
        # nearest-of-16 per block with the reference's distance expansion |x|^2 + |c|^2 - 2 x.c
        for lo in range(0, n, 1 << 16):
            x = data[lo:lo + (1 << 16)].reshape(-1, M, dpb).transpose(1, 0, 2)   # (M, rows, dpb)
            for m in range(M):
                xm, cm = x[m], books[m]
                part = np.einsum("ij,ij->i", xm, xm)[:, None] + np.einsum("ij,ij->i", cm, cm)[None] - 2 * xm @ cm.T
                codes[lo:lo + xm.shape[0], m] = np.argmin(part, axis=1)
        return TransformedData(true_n, transform_data(codes))

    def encode_device(self, rows, row_index=None, n_out=None):
        """Batched encoder on the GPU (new, additive API; `tkb_encode_dev`): rows f32/f64 (n, d) device tensor ->
        packed codes, int64 bit patterns (n_out/16, M) in the reference layout. Output position i encodes
        rows[row_index[i]] (an index outside [0, n) = the zero vector = the reference's padding rows); without
        row_index position i is row i and n_out defaults to n rounded up to 16."""
        t = D.require_cuda()
        assert self.centers is not None, "PQ has not been fitted"
        n, d = rows.shape
        rows = rows.contiguous()
        assert rows.dtype in (t.float32, t.float64)
        Dpad, Dp, M = self._lut_dims(d)
        if n_out is None:
            n_out = (len(row_index) if row_index is not None else n)
            n_out = -(-n_out // 16) * 16
        assert n_out % 16 == 0 and (row_index is None or len(row_index) == n_out)
        cen, R = self._dev_state()
        dpb = self.dims_per_block
        books = np.ascontiguousarray(self.centers, dtype=np.float32).reshape(16, M, dpb).transpose(1, 0, 2)
        cnorm = np.stack([np.einsum("ij,ij->i", b, b) for b in books]).astype(np.float32)   # utils.py:78 (Ynorm2, f32)
        out = D.empty((n_out // 16, M), np.int64)
        cnorm_dev = D.upload(cnorm)                                   # named: the tensor must outlive the launch
        check(lib.tkb_encode_dev(D.ptr(rows), DTYPE_F64 if rows.dtype == t.float64 else DTYPE_F32, n, d,
                                 D.ptr(row_index), n_out, D.ptr(cen), D.ptr(cnorm_dev), Dp, dpb, D.ptr(R), Dpad,
                                 D.ptr(out), D.stream_ptr()))
        return out

    # ------------------------------------------------------------------ query time (device) --
    def _dev_state(self):
        """Device copies of the quantizer (centers, R), rebuilt if the host arrays were replaced."""
        key = (id(self.centers), id(self.R))
        st = self.__dict__.get("_dev")
        if st is None or st[0] != key:
            st = (key, D.upload(np.ascontiguousarray(self.centers, dtype=np.float32)),
                  None if self.R is None else D.upload(np.ascontiguousarray(self.R, dtype=np.float64)))
            self.__dict__["_dev"] = st
        return st[1], st[2]

    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if k != "_dev"}

    def _lut_dims(self, d):
        dpb = self.dims_per_block
        Dpad = -(-d // (dpad * dpb)) * (dpad * dpb)
        Dp = self.centers.shape[1]
        if self.R is not None:
            assert self.R.shape == (Dp, Dpad), "query dimension does not match the fitted rotation"
        else:
            assert Dp == Dpad, "query dimension does not match the fitted quantizer"
        return Dpad, Dp, Dp // dpb

    def distance_tables(self, queries, signed=True, normalize=False, buf=None):
        """Batched LUT build (new, additive API). queries: f32 (Q, d) host array or device tensor.
        Returns a dict of device tensors: tables u8 (Q, M, 16), q f32 (Q, d) (normalised when
        `normalize`), q_rot f64 (Q, Dp), shift f64 (Q,), scale f64 (Q,). buf: optional allocator
        `buf(name, shape, np_dtype)` (IVF.query_batch reuses its per-stream workspace; default: fresh tensors)."""
        t = D.require_cuda()
        assert self.centers is not None, "PQ has not been fitted"
        if isinstance(queries, np.ndarray):
            queries = D.upload(np.ascontiguousarray(queries, dtype=np.float32))
        Q, d = queries.shape
        Dpad, Dp, M = self._lut_dims(d)
        cen, R = self._dev_state()
        if buf is None:
            buf = lambda name, shape, dt: D.empty(shape, dt)                 # noqa: E731
        out = dict(tables=buf("lut_tables", (Q, M, 16), np.uint8), q=buf("lut_q", (Q, d), np.float32),
                   q_rot=buf("lut_q_rot", (Q, Dp), np.float64), shift=buf("lut_shift", (Q,), np.float64),
                   scale=buf("lut_scale", (Q,), np.float64))
        check(lib.tkb_lut_build_dev(
            D.ptr(queries), Q, d, int(bool(normalize)), D.ptr(out["q"]), D.ptr(cen), Dp, self.dims_per_block,
            D.ptr(R), Dpad, float(self.sqrt_n_blocks), float(np.log(M)), int(bool(signed)),
            D.ptr(out["tables"]), D.ptr(out["q_rot"]), D.ptr(out["shift"]), D.ptr(out["scale"]), D.stream_ptr()))
        return out

    def _single_table(self, q, signed):
        q = np.asarray(q)
        (d,) = q.shape
        lut = self.distance_tables(np.ascontiguousarray(q, dtype=np.float32)[None], signed=signed)
        tables = lut["tables"][0].cpu().numpy().reshape(-1).view(np.uint64)
        q_rot = lut["q_rot"][0].cpu().numpy()
        shift = lut["shift"][0].item()
        if self.R is None:                      # unrotated path keeps f32 (ref: fast_pq.py:206-215)
            q_rot = q_rot.astype(np.float32)
            shift = np.float32(shift)
        else:
            shift = np.float64(shift)
        dt = _FastDistanceTable(q_rot, q, tables, shift, np.float64(lut["scale"][0].item()), signed=signed)
        dt._dev_tables = lut["tables"][0]
        return dt

    def distance_table(self, q):
        """ref: fast_pq.py:186-222"""
        return self._single_table(q, True)

    def udistance_table(self, q):
        """ref: fast_pq.py:224-252 (experimental in the reference)"""
        return self._single_table(q, False)


class _FastDistanceTable:
    def __init__(self, q, raw_q, transformed_tables, mean, scale, signed):
        self.q = q
        self.raw_q = raw_q
        self.tables = transformed_tables
        self.mean = mean
        self.scale = scale
        self.signed = signed
        self._dev_tables = None

    def __repr__(self):
        return (f"FastDistanceTable(q={self.q}, tables={self.tables}, mean={self.mean}, "
                f"scale={self.scale}, signed={self.signed})")

    def _tables_dev(self, M):
        if self._dev_tables is None:
            tab = np.ascontiguousarray(self.tables, dtype=np.uint64)
            assert tab.shape[0] >= 2 * M
            self._dev_tables = D.upload(tab[:2 * M].view(np.uint8))
        return self._dev_tables

    def _scan_dev(self, packed):
        """Device estimates (uint8 tensor of 16*n_chunks) of a host `packed` array."""
        if not (isinstance(packed, np.ndarray) and packed.dtype == np.uint64 and packed.ndim == 2
                and packed.flags.c_contiguous):
            raise ValueError("transformed data must be a C-contiguous 2-D uint64 array")
        n_chunks, M = packed.shape
        est = D.empty((16 * n_chunks,), np.uint8)
        if n_chunks and SCAN_IMPL == "fast":
            ws = D.scan_workspace(n_chunks)
            check(lib.tkb_estimate_native_dev(D.ptr(D.mirror_native(packed)), n_chunks, M, D.ptr(self._tables_dev(M)), 1,
                                              D.ptr(est), 16 * n_chunks, _order(), int(bool(self.signed)),
                                              D.ptr(ws), ws.numel(), D.stream_ptr()))
        elif n_chunks:
            check(lib.tkb_estimate_dev(D.ptr(D.mirror(packed)), n_chunks, M, D.ptr(self._tables_dev(M)), 1,
                                       D.ptr(est), 16 * n_chunks, _order(), int(bool(self.signed)), D.stream_ptr()))
        return est

    def estimate_distances(self, transformed_data, out=None, rescale=False):
        """ref: fast_pq.py:270-282"""
        true_n, packed = transformed_data
        n_chunks = len(packed)
        if out is None:
            out = np.zeros(2 * n_chunks, dtype=np.uint64)
        else:
            _kernels._buf(out, np.uint64, 1, "out", writable=True)
            if out.shape[0] < 2 * n_chunks:
                raise ValueError("out: need %d uint64" % (2 * n_chunks))
        est = self._scan_dev(packed)
        if n_chunks:
            D.torch().from_numpy(out.view(np.uint8)[:16 * n_chunks]).copy_(est)
        res = out.view(np.int8 if self.signed else np.uint8)[:true_n]
        if not rescale:
            return res
        as_float = np.ascontiguousarray(res, dtype=np.float32)
        return self.q @ self.q + (as_float / self.scale + self.mean)

    def _heap_dev(self, transformed_data, rescore):
        """Scan + exact heap replay on the device; returns device (indices i64[R], values i32[R])."""
        true_n, packed = transformed_data
        est = self._scan_dev(packed)
        hidx, hval = D.empty((rescore,), np.int64), D.empty((rescore,), np.int32)
        check(lib.tkb_replay_fresh_dev(D.ptr(est), 16 * len(packed), len(packed), true_n, D.ptr(hidx), D.ptr(hval),
                                       1, rescore, int(bool(self.signed)), D.stream_ptr()))
        return hidx, hval

    def top(self, transformed_data, data, k=1, rescore=None):
        """ref: fast_pq.py:284-312"""
        true_n, packed = transformed_data
        assert len(data) == true_n
        k = min(k, true_n)
        if not rescore:
            rescore = min(2 * k + 10, true_n)
        assert true_n >= rescore >= k
        hidx, _ = self._heap_dev(transformed_data, rescore)
        if rescore <= k:
            return hidx.cpu().numpy()
        dists = exact_dists_dev(data, self.raw_q, hidx)
        indices = hidx.cpu().numpy()
        return indices[bottom_k(dists.cpu().numpy(), k)]


def _rows_dev(data):
    """Device mirror of the raw data matrix in the dtype numpy's `Y - x` would produce (x is f32)."""
    if not isinstance(data, np.ndarray):
        data = np.asarray(data)
    if data.dtype == np.float32:
        return D.mirror(data) if data.flags.c_contiguous else D.upload(data), DTYPE_F32
    if data.dtype == np.float64 and data.flags.c_contiguous:
        return D.mirror(data), DTYPE_F64
    return D.upload(np.ascontiguousarray(data, dtype=np.float64)), DTYPE_F64


def exact_dists_dev(data, q, idx_dev):
    """|data[idx] - q|^2 for a device vector of row indices (the arithmetic of knn_brute1,
    ref: utils.py:89-91). Returns a device tensor (f32 for f32 data, else f64)."""
    rows, dt = _rows_dev(data)
    qd = D.upload(np.ascontiguousarray(q, dtype=np.float32).reshape(1, -1))
    assert qd.shape[1] == rows.shape[1], "query and data dimensions differ"
    R = idx_dev.shape[0]
    out = D.empty((R,), np.float32 if dt == DTYPE_F32 else np.float64)
    check(lib.tkb_gather_dists_dev(D.ptr(rows), dt, rows.shape[0], rows.shape[1], D.ptr(qd), D.ptr(idx_dev),
                                   1, R, D.ptr(out), D.stream_ptr()))
    return out
