"""numpy-buffer front end of the host-buffer C ABI (shared by _fast_pq and _fast_pq_avx).

The argument checks mirror what Cython's typed memoryviews enforce for the reference kernels
(`uint64_t[:, ::1] data`, `uint64_t[::1] tables`, `int64_t[::1] indices`, `int[::1] vals`, ...):
a wrong dtype, rank or a non C-contiguous buffer raises ValueError before anything is launched.
"""
import numpy as np

from . import _lib
from ._lib import lib, check, ORDER_SSE, ORDER_AVX


def _buf(a, dtype, ndim, name, writable=False):
    if not isinstance(a, np.ndarray):
        raise TypeError("%s: expected a numpy array, got %s" % (name, type(a).__name__))
    if a.dtype != dtype:
        raise ValueError("Buffer dtype mismatch for %s: expected %s but got %s" % (name, np.dtype(dtype), a.dtype))
    if a.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions for %s (expected %d, got %d)" % (name, ndim, a.ndim))
    if a.size and not a.flags.c_contiguous:
        raise ValueError("%s: ndarray is not C-contiguous" % name)
    if writable and not a.flags.writeable:
        raise ValueError("%s: buffer source array is read-only" % name)
    return a.ctypes.data


def estimate_pq(data, tables, out, signd, order):
    """ref: estimate_pq_sse (_fast_pq.pyx:101-111) / estimate_pq_avx (_fast_pq_256.pyx:52-62)"""
    dp = _buf(data, np.uint64, 2, "data")
    tp = _buf(tables, np.uint64, 1, "tables")
    op = _buf(out, np.uint64, 1, "out", writable=True)
    n_chunks, M = data.shape
    if n_chunks == 0:
        return
    if tables.shape[0] < 2 * M:
        raise ValueError("tables: need %d uint64 for %d sub-quantizers, got %d" % (2 * M, M, tables.shape[0]))
    if out.shape[0] < 2 * n_chunks:
        raise ValueError("out: need %d uint64, got %d" % (2 * n_chunks, out.shape[0]))
    check(lib.tkb_estimate_pq_host(dp, n_chunks, M, tp, op, order, int(bool(signd))))


def query_pq(data, n, tables, indices, vals, signd, labels, order):
    """ref: query_pq_sse (_fast_pq.pyx:114-206) / query_pq_avx (_fast_pq_256.pyx:65-123)"""
    dp = _buf(data, np.uint64, 2, "data")
    tp = _buf(tables, np.uint64, 1, "tables")
    ip = _buf(indices, np.int64, 1, "indices", writable=True)
    vp = _buf(vals, np.int32, 1, "vals", writable=True)
    n_chunks, M = data.shape
    n = int(n)
    R = indices.shape[0]
    if vals.shape[0] < R:
        raise ValueError("vals: shorter than indices")
    if n_chunks == 0 or R == 0:
        return
    if tables.shape[0] < 2 * M:
        raise ValueError("tables: need %d uint64 for %d sub-quantizers, got %d" % (2 * M, M, tables.shape[0]))
    lp = None
    if labels is not None:
        lp = _buf(labels, np.int64, 1, "labels")
        if labels.shape[0] < min(n, 16 * n_chunks):
            raise ValueError("labels: need at least n entries")
    check(lib.tkb_query_pq_host(dp, n_chunks, M, n, tp, ip, vp, R, order, int(bool(signd)), lp))


def init_heap(indices, vals, signd):
    """ref: init_heap (_fast_pq.pyx:240-252)"""
    ip = _buf(indices, np.int64, 1, "indices", writable=True)
    vp = _buf(vals, np.int32, 1, "vals", writable=True)
    check(lib.tkb_init_heap(ip, vp, indices.shape[0], int(bool(signd))))


def insert(indices, vals, i, v):
    """ref: insert (_fast_pq.pyx:274-307)"""
    ip = _buf(indices, np.int64, 1, "indices", writable=True)
    vp = _buf(vals, np.int32, 1, "vals", writable=True)
    check(lib.tkb_insert(ip, vp, indices.shape[0], int(i), int(v)))


def insert_is(indices, vals, i, v):
    """ref: insert_is (_fast_pq.pyx:256-271)"""
    ip = _buf(indices, np.int64, 1, "indices", writable=True)
    vp = _buf(vals, np.int32, 1, "vals", writable=True)
    check(lib.tkb_insert_is(ip, vp, indices.shape[0], int(i), int(v)))
