"""Bit layouts of the kernel boundary (ref: tinyknn/_transform.py). Host-side, build-time.

The contract (pinned by the reference's tests/test_transform.py:80-101): with codes of shape
(n, M), chunk c = vectors 16c..16c+15, `packed[c, 2p + h]` is a uint64 whose byte b belongs to
vector 16c + 8h + b, low nibble = code of sub-quantizer 2p, high nibble = sub-quantizer 2p + 1.
Seen as bytes, each chunk row is M/2 groups of 16 bytes: one byte per vector of the chunk.
"""
import numpy as np


def transform_data(data0):
    data0 = np.asarray(data0)
    n, d = data0.shape
    assert n % 16 == 0, "Number of rows must be divisible by 16"
    assert np.all(data0 < 16) and np.all(0 <= data0), "Input must be 4 bit values"
    nib = data0.astype(np.uint8, copy=False).reshape(n // 16, 16, d // 2, 2)
    pairs = nib[..., 0] | (nib[..., 1] << 4)                # [chunk, vector, pair]
    by_pair = np.ascontiguousarray(pairs.transpose(0, 2, 1))  # [chunk, pair, vector]
    return by_pair.reshape(n // 16, d * 8).view(np.uint64)


def unpack(transformed_data):
    chunks, d = transformed_data.shape
    by_pair = np.ascontiguousarray(transformed_data).view(np.uint8).reshape(chunks, d // 2, 16)
    pairs = by_pair.transpose(0, 2, 1)                       # [chunk, vector, pair]
    out = np.empty((chunks, 16, d // 2, 2), dtype=np.uint64)
    out[..., 0] = pairs & 15
    out[..., 1] = pairs >> 4
    return out.reshape(chunks * 16, d)


def transform_tables(tables0):
    d, b = tables0.shape
    assert b == 16
    assert tables0.dtype == np.uint8
    return np.ascontiguousarray(tables0).reshape(2 * d, 8).view(np.uint64)[:, 0]
