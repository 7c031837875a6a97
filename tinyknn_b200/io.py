"""Stable on-disk format for a built index.

The reference has no format of its own: its benchmark pickles the `(pq, ivf)` pair (ref: bench.py:88-103), which ties a
file to the Python classes and copies every per-list array on load. `IVF` / `FastPQ` objects of this package pickle the
same way (tests/test_host_api.py); this module adds a layout that can be memory-mapped and uploaded without touching the
bytes: a directory of plain `.npy` files plus `meta.json`.

    meta.json            format name + version, metric, n_clusters, the FastPQ parameters, `dpad` of the build (4 = avx
                         order, 2 = sse: it fixes the padded dimension), shapes
    pq_centers.npy       f32 (16, Dp)            FastPQ.centers
    pq_R.npy             f64 (Dp, Dpad)          FastPQ.R (absent when the quantizer does not rotate)
    all_centers.npy      (n_clusters, d)         IVF.all_centers
    active_centers.npy   f32 (C, d)              IVF.active_centers
    center_codes.npy     u64 (ceil(C/16), M)     IVF.pq_transformed_centers.packed (reference chunk layout)
    list_sizes.npy       i64 (n_clusters,)       vectors per list; -1 = the reference's `None` slot
    codes.npy            u64 (sum ceil(n_l/16), M)  the lists' packed codes back to back, each padded to whole chunks
    ids.npy              i64 (sum n_l,)          the lists' ids back to back
    data.npy             (N, d)                  IVF.data (optional: include_data=False leaves it out)

`load_index(path, mmap=True)` returns an `IVF` whose per-list attributes are views into the mapped files -- the types
and layouts of the reference (ref: ivf.py:77, 91-102), so the oracle and the compiled reference can query it too.
"""
import json
import os

import numpy as np

from .fast_pq import FastPQ, TransformedData
from .ivf import IVF

FORMAT, VERSION = "tinyknn-b200-index", 1


def _is_td(x):
    return isinstance(x, tuple) and len(x) == 2 and isinstance(x[1], np.ndarray)


def save_index(ivf, path, include_data=True):
    """Write a built IVF (after `fit` + `build`, or one produced by synth.build_ivf) to directory `path`."""
    from . import fast_pq as _fp
    os.makedirs(path, exist_ok=True)
    pq = ivf.pq
    assert pq.centers is not None and _is_td(ivf.pq_transformed_centers), "index has not been built"
    M = int(ivf.pq_transformed_centers.packed.shape[1])
    n_lists = len(ivf.pq_transformed_points)
    sizes = np.full(n_lists, -1, dtype=np.int64)
    code_parts, id_parts = [], []
    for l, td in enumerate(ivf.pq_transformed_points):
        if td is None:
            continue
        if not _is_td(td):                                   # FastPQ.transform returns an empty input unchanged
            sizes[l] = 0
            continue
        n_l, packed = int(td[0]), td[1]
        assert packed.dtype == np.uint64 and packed.shape == ((n_l + 15) // 16, M)
        sizes[l] = n_l
        code_parts.append(packed)
        ids_l = np.asarray(ivf.ids[l], dtype=np.int64)
        assert ids_l.shape == (n_l,)
        id_parts.append(ids_l)
    codes = np.concatenate(code_parts) if code_parts else np.zeros((0, M), dtype=np.uint64)
    ids = np.concatenate(id_parts) if id_parts else np.zeros(0, dtype=np.int64)
    arrays = dict(pq_centers=np.ascontiguousarray(pq.centers, dtype=np.float32), all_centers=np.asarray(ivf.all_centers),
                  active_centers=np.ascontiguousarray(ivf.active_centers, dtype=np.float32),
                  center_codes=np.ascontiguousarray(ivf.pq_transformed_centers.packed), list_sizes=sizes, codes=codes, ids=ids)
    if pq.R is not None:
        arrays["pq_R"] = np.ascontiguousarray(pq.R, dtype=np.float64)
    has_data = bool(include_data and isinstance(getattr(ivf, "data", None), np.ndarray))
    if has_data:
        arrays["data"] = ivf.data
    for name, a in arrays.items():
        np.save(os.path.join(path, name + ".npy"), a, allow_pickle=False)
    meta = dict(format=FORMAT, version=VERSION, metric=ivf.metric, n_clusters=int(ivf.n_clusters), M=M,
                center_size=int(ivf.pq_transformed_centers.size), dpad=int(_fp.dpad), has_data=has_data,
                has_R=pq.R is not None,
                pq=dict(dims_per_block=int(pq.dims_per_block), use_kmeans=bool(pq.use_kmeans),
                        rotate_dim=None if pq.rotate_dim is None else int(pq.rotate_dim),
                        sqrt_n_blocks=float(pq.sqrt_n_blocks)))
    tmp = os.path.join(path, "meta.json.tmp")
    with open(tmp, "w") as f:
        json.dump(meta, f, indent=1)
    os.replace(tmp, os.path.join(path, "meta.json"))         # meta.json last: its presence marks a complete index
    return path


def load_index(path, mmap=True, data=None):
    """Read an index written by `save_index`. mmap=True maps the arrays read-only (per-list attributes are views, nothing
    is copied until `to_device()` uploads). `data`: the raw vectors when the index was saved without them."""
    from . import fast_pq as _fp
    with open(os.path.join(path, "meta.json")) as f:
        meta = json.load(f)
    if meta.get("format") != FORMAT or int(meta.get("version", -1)) > VERSION:
        raise ValueError("not a %s directory (or written by a newer version): %s" % (FORMAT, path))
    if int(meta["dpad"]) != int(_fp.dpad):
        raise ValueError("index was built with dpad=%d (%s order) but the active order pads to %d; call fast_pq.set_order first"
                         % (meta["dpad"], "avx" if meta["dpad"] == 4 else "sse", _fp.dpad))
    ld = lambda name: np.load(os.path.join(path, name + ".npy"), mmap_mode="r" if mmap else None, allow_pickle=False)
    p = meta["pq"]
    pq = FastPQ(p["dims_per_block"], use_kmeans=p["use_kmeans"], rotate_dim=p["rotate_dim"])
    ivf = IVF(meta["metric"], meta["n_clusters"], pq)
    pq.centers = np.ascontiguousarray(ld("pq_centers"))
    pq.R = np.ascontiguousarray(ld("pq_R")) if meta["has_R"] else None
    pq.sqrt_n_blocks = np.float64(p["sqrt_n_blocks"])
    ivf.all_centers = np.ascontiguousarray(ld("all_centers"))
    ivf.active_centers = np.ascontiguousarray(ld("active_centers"))
    ivf.pq_transformed_centers = TransformedData(int(meta["center_size"]), np.ascontiguousarray(ld("center_codes")))
    sizes, codes, ids = np.asarray(ld("list_sizes")), ld("codes"), ld("ids")
    M, d = int(meta["M"]), ivf.active_centers.shape[1]
    assert codes.dtype == np.uint64 and codes.ndim == 2 and codes.shape[1] == M
    nc = np.where(sizes > 0, (sizes + 15) // 16, 0)
    c_off = np.concatenate(([0], np.cumsum(nc)))
    i_off = np.concatenate(([0], np.cumsum(np.maximum(sizes, 0))))
    if c_off[-1] != codes.shape[0] or i_off[-1] != ids.shape[0]:
        raise ValueError("corrupt index: list sizes do not add up to the stored codes / ids")
    for l, n_l in enumerate(sizes):
        if n_l < 0:
            continue
        if n_l == 0:
            ivf.pq_transformed_points[l] = np.empty((0, d))
            ivf.ids[l] = np.empty(0)
        else:
            ivf.pq_transformed_points[l] = TransformedData(int(n_l), codes[c_off[l]:c_off[l + 1]])
            ivf.ids[l] = ids[i_off[l]:i_off[l + 1]]
    if meta["has_data"]:
        ivf.data = ld("data")
    elif data is not None:
        ivf.data = data
    return ivf
