"""Drop-in for the reference module `tinyknn._fast_pq` (ref: tinyknn/_fast_pq.pyx): the SSE
accumulation order, executed by the sm_100a kernels behind the host-buffer C ABI."""
from . import _kernels
from ._kernels import init_heap, insert, insert_is  # noqa: F401  (same names as the reference module)
from ._lib import ORDER_SSE


def estimate_pq_sse(data, tables, out, signd):
    _kernels.estimate_pq(data, tables, out, signd, ORDER_SSE)


def query_pq_sse(data, n, tables, indices, vals, signd, labels=None):
    _kernels.query_pq(data, n, tables, indices, vals, signd, labels, ORDER_SSE)
