"""tinyknn_b200 -- B200-native drop-in for the query hot path of thomasahle/tinyknn.

Same export list as the reference package (ref: tinyknn/__init__.py:1-6):
    import tinyknn_b200 as tinyknn
"""
from . import _transform
from . import _fast_pq
from . import _fast_pq_avx
from .fast_pq import FastPQ, avx
from .ivf import IVF
from . import utils
from .utils import bottom_k, bottom_k_2d, cdist, knn_brute, group_data_by_indices
from .io import save_index, load_index          # additive: stable on-disk index format (not part of the reference's exports)

__all__ = ["FastPQ", "IVF", "avx", "utils", "bottom_k", "bottom_k_2d", "cdist", "knn_brute",
           "group_data_by_indices", "_transform", "_fast_pq", "_fast_pq_avx"]
