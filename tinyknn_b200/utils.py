"""Host-side helpers with the reference's names and behaviour (ref: tinyknn/utils.py).

Only `pad1`, `bottom_k` and `knn_brute1` sit on the query path, and there the arithmetic
(`knn_brute1`'s distances) runs on the GPU -- see fast_pq._FastDistanceTable.top and ivf.IVF.query.
The rest (ground truth, build-time grouping) is kept so `tinyknn_b200` is importable wherever
`tinyknn` was.
"""
import time
from contextlib import contextmanager

import numpy as np


def _padded(arr, shape):
    out = np.zeros(shape, dtype=arr.dtype)
    out[tuple(slice(0, s) for s in arr.shape)] = arr
    return out


def pad1(arr, m):
    """Zero-pad a vector to a multiple of m (ref: utils.py:6-11)."""
    (s,) = arr.shape
    return _padded(arr, (-(-s // m) * m,))


def pad2(arr, m1, m2):
    """Zero-pad a matrix to multiples of (m1, m2) (ref: utils.py:14-19)."""
    s1, s2 = arr.shape
    return _padded(arr, (-(-s1 // m1) * m1, -(-s2 // m2) * m2))


def bottom_k(arr, k):
    """Indices of the k smallest entries, unordered (ref: utils.py:22-25)."""
    n = len(arr)
    return np.arange(n) if k >= n else np.argpartition(arr, k)[:k]


def bottom_k_2d(arr, k):
    """Row-wise bottom_k (ref: utils.py:28-31)."""
    rows, cols = arr.shape
    if k >= cols:
        return np.resize(np.arange(cols), arr.shape)
    return np.argpartition(arr, k, axis=1)[:, :k]


@contextmanager
def timer(verbose, text):
    """ref: utils.py:34-41"""
    t0 = time.time()
    if verbose:
        print(text)
    yield
    if verbose:
        print(f"Took {time.time() - t0:.1f}s")


def _sqnorms(A):
    return np.einsum("ij,ij->i", A, A)


def cdist(X, Y, chunk=100):
    """Squared Euclidean distances R[i, j] = |X_i - Y_j|^2 (ref: utils.py:44-63)."""
    ny = _sqnorms(Y)
    out = np.zeros((X.shape[0], Y.shape[0]))
    for lo in range(0, X.shape[0], chunk):
        blk = X[lo:lo + chunk]
        out[lo:lo + chunk] = _sqnorms(blk)[:, None] + ny
        out[lo:lo + chunk] -= 2 * blk @ Y.T
    return out


def knn_brute(X, Y, k, metric="euclidean", chunk=100):
    """k nearest rows of Y for every row of X, unordered (ref: utils.py:66-86)."""
    assert k <= Y.shape[0], f"Can't find knn with {k=} and {Y.shape[0]} targets."
    if metric == "angular":
        X = X / np.linalg.norm(X, axis=1, keepdims=True)
        Y = Y / np.linalg.norm(Y, axis=1, keepdims=True)
    elif metric != "euclidean":
        raise ValueError(f"Metric not supported: {metric}")
    ny = _sqnorms(Y)
    res = np.zeros((X.shape[0], k), dtype=int)
    for lo in range(0, X.shape[0], chunk):
        blk = X[lo:lo + chunk]
        part = _sqnorms(blk)[:, None] + ny[None] - 2 * blk @ Y.T
        res[lo:lo + chunk] = bottom_k_2d(part, k)
    return res


def knn_brute1(x, Y, k):
    """Exact nearest k rows of Y to x (ref: utils.py:89-92). Host numpy version, kept for API parity;
    the query path computes these distances with tkb_gather_dists_dev instead."""
    diff = Y - x
    return bottom_k(np.einsum("ij,ij->i", diff, diff), k)


def group_data_by_indices(X, indices, k):
    """Split rows of X into k groups: row i goes to every group listed in indices[i]
    (ref: utils.py:95-162). Returns (parts, ids); within a group rows are ordered column by column
    of `indices`, and inside a column by argsort of that column (the reference's order)."""
    assert 0 <= np.min(indices) and np.max(indices) < k
    parts = [[] for _ in range(k)]
    ids = [[] for _ in range(k)]
    for col in indices.T:
        order = np.argsort(col)
        groups, counts = np.unique(col[order], return_counts=True)
        bounds = np.concatenate(([0], np.cumsum(counts)))
        rows = X[order]
        for g, lo, hi in zip(groups, bounds[:-1], bounds[1:]):
            parts[g].append(rows[lo:hi])
            ids[g].append(order[lo:hi])
    for part, id_list in zip(parts, ids):
        if not part:
            part.append(np.empty((0, X.shape[1])))
            id_list.append(np.empty(0))
    return [np.vstack(p) for p in parts], [np.hstack(i) for i in ids]


def knn_brute_device(X, Y, k, metric="euclidean"):
    """knn_brute (ref: utils.py:66-86) for k in {1, 2} with the rows x centroids contraction on the GPU (`tkb_assign_dev`).
    The norms are numpy's own (np.einsum, as in the reference), the dot products are the FMA chain the reference's BLAS
    computes, so for k = 1 the result equals knn_brute's wherever the minimum is unique; for k = 2 each row holds the
    same two indices in ascending-distance order (np.argpartition's order is unspecified). Returns int (n, k)."""
    from . import _device as D
    from ._lib import lib, check, DTYPE_F32, DTYPE_F64
    assert 1 <= k <= 2 and k <= Y.shape[0], f"Can't find knn with {k=} and {Y.shape[0]} targets."
    if metric == "angular":
        X = X / np.linalg.norm(X, axis=1, keepdims=True)
        Y = Y / np.linalg.norm(Y, axis=1, keepdims=True)
    elif metric != "euclidean":
        raise ValueError(f"Metric not supported: {metric}")
    T = np.result_type(X.dtype, Y.dtype)
    T = np.float32 if T == np.float32 else np.float64
    xn = _sqnorms(X).astype(T)                       # Xnorm2 / Ynorm2 in their own dtypes, promoted like numpy would
    yn = _sqnorms(Y).astype(T)
    Xd, Yd = D.upload(np.ascontiguousarray(X, dtype=T)), D.upload(np.ascontiguousarray(Y, dtype=T))
    xnd, ynd = D.upload(xn), D.upload(yn)            # named: the tensors must outlive the launch
    out = D.empty((X.shape[0], k), np.int32)
    check(lib.tkb_assign_dev(D.ptr(Xd), DTYPE_F32 if T == np.float32 else DTYPE_F64, X.shape[0], X.shape[1], D.ptr(Yd),
                             Y.shape[0], D.ptr(xnd), D.ptr(ynd), k, D.ptr(out), None, 0, D.stream_ptr()))
    return out.cpu().numpy().astype(int)
