"""Drop-in for the reference module `tinyknn._fast_pq_avx` (ref: tinyknn/_fast_pq_256.pyx, built
under that name by setup.py:39-41): the default two-lane AVX accumulation order."""
from . import _kernels
from ._kernels import init_heap, insert, insert_is  # noqa: F401
from ._lib import ORDER_AVX


def estimate_pq_avx(data, tables, out, signd):
    _kernels.estimate_pq(data, tables, out, signd, ORDER_AVX)


def query_pq_avx(data, n, tables, indices, vals, signd, labels=None):
    _kernels.query_pq(data, n, tables, indices, vals, signd, labels, ORDER_AVX)
