"""Synthetic index construction for benchmarks and large parity tests (BUILD-TIME TOOLING).

The reference builds indexes with sklearn KMeans and a Python loop of tiny GEMMs
(ref: tinyknn/ivf.py:19-104, tinyknn/fast_pq.py:50-184) -- hours at 10M-100M vectors. The hot path
this repository rebuilds only *consumes* an index, so to make inputs of the BASELINE.json shapes
this module builds one with plain torch ops on the GPU (Lloyd iterations, nearest-of-16 encoding,
list grouping) and hands back an ordinary `IVF` object whose attributes have exactly the
reference's types and layouts (ref: ivf.py:77, 91-102), so the oracle / the compiled reference can
query the very same index on the CPU. Nothing here is on the timed path.
"""
import numpy as np

from . import _device as D
from . import fast_pq as _fp
from .fast_pq import FastPQ, TransformedData
from .ivf import IVF


def clustered(n, d, n_components, seed, normalize=False, sigma=1.0, device=None, dtype=None):
    """Gaussian mixture: means ~ N(0, 4I), points = mean + N(0, sigma^2 I)."""
    t = D.require_cuda()
    device = device or D.device()
    g = t.Generator(device=device).manual_seed(seed)
    means = t.randn(n_components, d, generator=g, device=device) * 2
    X = t.empty(n, d, device=device, dtype=t.float32)
    step = 1 << 22                                                 # blocks: 100M x 128 must not need a second 51 GB array
    for lo in range(0, n, step):
        m = min(step, n - lo)
        comp = t.randint(n_components, (m,), generator=g, device=device)
        blk = t.randn(m, d, generator=g, device=device)
        if sigma != 1.0:
            blk *= sigma
        blk += means[comp]
        if normalize:
            blk /= blk.norm(dim=1, keepdim=True)
        X[lo:lo + m] = blk
    return X


class DeviceRows:
    """Row access to a device matrix with numpy semantics, for indexes whose raw vectors are too large to mirror on the
    host (100M x 128 f32 = 51 GB): `rows[ids]` fetches just those rows. Enough for the oracle's rescoring."""

    def __init__(self, X):
        self.X = X
        self.shape = tuple(X.shape)
        self.dtype = np.dtype(np.float32)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        t = D.torch()
        idx = t.as_tensor(np.asarray(idx, dtype=np.int64), device=self.X.device)
        idx = t.where(idx < 0, idx + self.shape[0], idx)
        return self.X[idx].cpu().numpy()


def _nearest(X, C, chunk=None):
    t = D.torch()
    if chunk is None:                                              # the (chunk x centres) distance block stays around 1 GB
        chunk = max(1024, min(1 << 18, (1 << 28) // max(1, C.shape[0])))
    out = t.empty(X.shape[0], dtype=t.int64, device=X.device)
    cn = (C * C).sum(1)
    for lo in range(0, X.shape[0], chunk):
        x = X[lo:lo + chunk]
        out[lo:lo + chunk] = (cn[None] - 2 * x @ C.T).argmin(1)
    return out


def kmeans(X, k, iters, seed, deterministic=False):
    """A few Lloyd iterations (empty clusters are re-seeded from random points). deterministic=True sums the members of a
    cluster in a fixed order (index_add_ uses atomics otherwise): every rank of a multi-GPU job that builds the index from
    the same seed must end up with the very same centroids, hence the same lists."""
    t = D.torch()
    g = t.Generator(device=X.device).manual_seed(seed)
    C = X[t.randperm(X.shape[0], generator=g, device=X.device)[:k]].clone()
    for _ in range(iters):
        a = _nearest(X, C)
        cnt = t.bincount(a, minlength=k)
        if deterministic:
            prev = t.are_deterministic_algorithms_enabled()
            t.use_deterministic_algorithms(True)
            try:
                S = t.zeros_like(C).index_add_(0, a, X)
            finally:
                t.use_deterministic_algorithms(prev)
        else:
            S = t.zeros_like(C).index_add_(0, a, X)
        C = t.where(cnt[:, None] > 0, S / cnt.clamp(min=1)[:, None], C)
        empty = (cnt == 0).nonzero().flatten()
        if len(empty):
            C[empty] = X[t.randint(X.shape[0], (len(empty),), generator=g, device=X.device)]
    return C


def fit_pq(X_sample, dims_per_block=2, rotate_dim=64, iters=8, seed=0, dpad=None):
    """FastPQ with torch-fitted codebooks: attributes as after FastPQ.fit (ref: fast_pq.py:50-104)."""
    t = D.torch()
    n, true_d = X_sample.shape
    dpb = dims_per_block
    dpad = _fp.dpad if dpad is None else dpad                      # ref: fast_pq.py:21-27 (4 for the avx build, 2 for sse)
    Dpad = -(-true_d // (dpad * dpb)) * (dpad * dpb)
    Xp = t.zeros(n, Dpad, device=X_sample.device, dtype=t.float32)
    Xp[:, :true_d] = X_sample
    pq = FastPQ(dpb, use_kmeans=True, rotate_dim=rotate_dim)
    d = Dpad
    if rotate_dim is not None and true_d != 100:                  # ref: fast_pq.py:77-82
        g = t.Generator(device="cpu").manual_seed(seed)
        Qm, _ = t.linalg.qr(t.randn(Dpad, Dpad, generator=g, dtype=t.float64))
        R = Qm.T.contiguous()
        if Dpad > rotate_dim:
            d = rotate_dim
            R = R[:d].contiguous()
        pq.R = R.numpy()
        Xp = (Xp.double() @ R.to(Xp.device).T).float()
    M = d // dpb
    blocks = Xp.reshape(n, M, dpb).permute(1, 0, 2).contiguous()   # (M, n, dpb)
    g = t.Generator(device=Xp.device).manual_seed(seed + 1)
    C = blocks[:, t.randperm(n, generator=g, device=Xp.device)[:16]].clone()     # (M, 16, dpb)
    for _ in range(iters):
        a = t.cdist(blocks, C).argmin(2)                           # (M, n)
        oh = t.nn.functional.one_hot(a, 16).to(blocks.dtype)       # (M, n, 16)
        cnt = oh.sum(1)                                            # (M, 16)
        S = oh.transpose(1, 2) @ blocks                            # (M, 16, dpb)
        C = t.where(cnt[..., None] > 0, S / cnt.clamp(min=1)[..., None], C)
    pq.centers = np.ascontiguousarray(C.permute(1, 0, 2).reshape(16, d).cpu().numpy(), dtype=np.float32)
    pq.sqrt_n_blocks = np.sqrt(d // dpb)
    return pq


def encode(pq, X, dpad=None, chunk=1 << 17):
    """Nearest-of-16 codes per block, uint8 (n, M), on the device (ref: fast_pq.py:147-184)."""
    t = D.torch()
    dpb = pq.dims_per_block
    n, true_d = X.shape
    dpad = _fp.dpad if dpad is None else dpad
    Dpad = -(-true_d // (dpad * dpb)) * (dpad * dpb)
    d = pq.centers.shape[1]
    M = d // dpb
    C = t.from_numpy(pq.centers).to(X.device).reshape(16, M, dpb).permute(1, 0, 2).contiguous()   # (M,16,dpb)
    R = None if pq.R is None else t.from_numpy(pq.R).to(X.device).float()
    codes = t.empty(n, M, dtype=t.uint8, device=X.device)
    for lo in range(0, n, chunk):
        x = X[lo:lo + chunk].float()
        if Dpad != true_d:
            x = t.nn.functional.pad(x, (0, Dpad - true_d))
        if R is not None:
            x = x @ R.T
        xb = x.reshape(-1, M, dpb).permute(1, 0, 2)                # (M, rows, dpb)
        codes[lo:lo + chunk] = t.cdist(xb, C).argmin(2).T.to(t.uint8)
    return codes


def pack_codes(codes):
    """torch version of transform_data (ref: _transform.py:4-77): uint8 (n16, M) -> int64 bit patterns (n16/16, M)."""
    t = D.torch()
    n, M = codes.shape
    assert n % 16 == 0 and M % 2 == 0
    c = codes.reshape(n // 16, 16, M // 2, 2)
    byte = c[..., 0] | (c[..., 1] << 4)                            # (chunks, 16, M/2)
    return byte.permute(0, 2, 1).contiguous().reshape(n // 16, M * 8).view(t.int64)


def build_ivf(X, metric, n_clusters, pq=None, kmeans_iters=6, fit_sample=200_000, seed=0, keep_device=True, host_data=True,
              deterministic=False):
    """IVF index over device data X (f32, (n, d)); one list per point (build_probes = 1).
    Returns an `IVF` whose host attributes mirror the reference's and whose device copy is in place."""
    t = D.require_cuda()
    n, d = X.shape
    if metric == "angular":
        X = X / X.norm(dim=1, keepdim=True)
    g = t.Generator(device=X.device).manual_seed(seed)
    sample = X[t.randperm(n, generator=g, device=X.device)[:min(n, fit_sample)]]
    centers = kmeans(sample, n_clusters, kmeans_iters, seed, deterministic=deterministic)
    if metric == "angular":
        centers = centers / centers.norm(dim=1, keepdim=True)
    if pq is None:
        pq = fit_pq(sample[:100_000], seed=seed)
    assign = _nearest(X, centers)
    used = t.unique(assign)                                        # active centres (ref: ivf.py:91-93)
    remap = t.full((n_clusters,), -1, dtype=t.int64, device=X.device)
    remap[used] = t.arange(len(used), device=X.device)
    assign = remap[assign]
    C = len(used)
    active = centers[used].contiguous()
    order = t.argsort(assign, stable=True)
    sizes = t.bincount(assign, minlength=C)
    chunks = ((sizes + 127) // 128) * 8                            # whole tiles of 8 chunks (native layout)
    chunk_off = t.zeros(C + 1, dtype=t.int64, device=X.device)
    chunk_off[1:] = chunks.cumsum(0)
    start = t.zeros(C + 1, dtype=t.int64, device=X.device)
    start[1:] = sizes.cumsum(0)
    total_slots = int(chunk_off[-1].item()) * 16
    sorted_list = assign[order]
    slot = 16 * chunk_off[sorted_list] + (t.arange(n, device=X.device) - start[sorted_list])
    codes = encode(pq, X)
    zero_code = encode(pq, t.zeros(1, d, device=X.device))[0]      # padding rows are zero vectors (fast_pq.py:165)
    all_codes = zero_code[None].repeat(total_slots, 1)
    all_codes[slot] = codes[order]
    packed = pack_codes(all_codes)
    ids_padded = t.full((total_slots,), -1, dtype=t.int64, device=X.device)
    ids_padded[slot] = order
    center_td = pq_transform_small(pq, active)

    ivf = IVF(metric, n_clusters, FastPQ(pq.dims_per_block, rotate_dim=pq.rotate_dim))
    ivf.pq.centers, ivf.pq.R, ivf.pq.sqrt_n_blocks = pq.centers, pq.R, pq.sqrt_n_blocks
    ivf.all_centers = centers.cpu().numpy()
    ivf.active_centers = np.ascontiguousarray(active.cpu().numpy(), dtype=np.float32)
    ivf.pq_transformed_centers = center_td
    packed_h = packed.cpu().numpy().view(np.uint64)
    order_h, sizes_h, off_h, start_h = order.cpu().numpy(), sizes.cpu().numpy(), chunk_off.cpu().numpy(), start.cpu().numpy()
    # host views hold exactly ceil(n/16) chunks per list, like the reference (tile padding is device-only)
    ivf.pq_transformed_points = [TransformedData(int(sizes_h[l]), packed_h[off_h[l]:off_h[l] + (int(sizes_h[l]) + 15) // 16])
                                 for l in range(C)] + [None] * (n_clusters - C)
    ivf.ids = [order_h[start_h[l]:start_h[l + 1]] for l in range(C)] + [None] * (n_clusters - C)
    ivf.data = X.cpu().numpy() if host_data else DeviceRows(X)       # host_data=False: rows are fetched on demand
    if keep_device:
        from ._lib import DTYPE_F32
        M = packed.shape[1]
        off_full = t.cat([chunk_off, chunk_off[-1:].repeat(n_clusters - C)])
        sizes_full = t.cat([sizes, t.zeros(n_clusters - C, dtype=sizes.dtype, device=X.device)]).to(t.int32)
        ivf.__dict__["_dev"] = dict(
            C=C, M=M, n_lists=n_clusters, max_chunks=int(chunks.max().item()),
            max_real_chunks=int((int(sizes.max().item()) + 15) // 16),
            codes=D.to_native(packed, packed.shape[0], M), n_chunks_total=int(packed.shape[0]), list_chunk_off=off_full,
            list_size=sizes_full, ids=ids_padded,
            center_codes=D.to_native(D.upload(center_td.packed), center_td.packed.shape[0], M),
            center_chunks=int(center_td.packed.shape[0]),
            centers=active.float().contiguous(), data=X.contiguous(), data_dtype=DTYPE_F32, d=d,
            host_sizes=sizes_full.cpu().numpy(), host_chunks=off_full.cpu().numpy(), unique_ids=True)
    return ivf


def pq_transform_small(pq, rows):
    """TransformedData of a small device matrix (used for the centroid codes)."""
    t = D.torch()
    n = rows.shape[0]
    n16 = -(-n // 16) * 16
    padded = t.zeros(n16, rows.shape[1], device=rows.device, dtype=rows.dtype)
    padded[:n] = rows
    return TransformedData(n, pack_codes(encode(pq, padded)).cpu().numpy().view(np.uint64))


def index_fingerprint(ivf):
    """Integer checksums (device int64[6], wrap-around sums of the bit patterns) of everything a rank of a list-sharded job
    must agree on with the others: list sizes, codes, ids, centroids, centroid codes, the PQ codebook."""
    t = D.torch()
    dev = ivf.to_device()
    bits = lambda x: x.contiguous().view(t.int32).to(t.int64).sum() if x.dtype == t.float32 else x.contiguous().view(-1).to(t.int64).sum()
    sizes = dev["list_size"].to(t.int64)
    codes = dev["codes"]
    n8 = codes.numel() // 8 * 8
    pqc = D.upload(np.ascontiguousarray(ivf.pq.centers, dtype=np.float32))
    return t.stack([(sizes * t.arange(1, len(sizes) + 1, device=sizes.device)).sum(), codes[:n8].view(t.int64).sum(),
                    dev["ids"].sum(), bits(dev["centers"]), dev["center_codes"][:dev["center_codes"].numel() // 8 * 8].view(t.int64).sum(),
                    bits(pqc)])


def index_consistent(ivf, dist, group=None):
    """True when every rank holds the same index (collective)."""
    fp = index_fingerprint(ivf)
    lo, hi = fp.clone(), fp.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool((lo == hi).all().item())


_REPL_SCALARS = ("C", "M", "n_lists", "max_chunks", "max_real_chunks", "n_chunks_total", "center_chunks", "d", "data_dtype")
_REPL_ARRAYS = ("codes", "list_chunk_off", "list_size", "ids", "center_codes", "centers", "data")


def _bcast_tensor(x, dist, src, group, like=None):
    """Broadcast one device tensor from `src` (the receivers pass x=None): header (dtype code, ndim, shape), then the payload
    in slices of at most 4 GiB (a 51 GB raw-vector matrix travels over NVLink in a fraction of a second per slice)."""
    t = D.torch()
    codes = [t.uint8, t.int8, t.int32, t.int64, t.float32, t.float64]
    hdr = t.zeros(8, dtype=t.int64, device=D.device())
    if x is not None:
        x = x.contiguous()
        hdr[0], hdr[1] = codes.index(x.dtype), x.dim()
        for i, n in enumerate(x.shape):
            hdr[2 + i] = n
    dist.broadcast(hdr, src, group=group)
    h = [int(v) for v in hdr.cpu().tolist()]
    shape = tuple(h[2:2 + h[1]])
    if x is None:
        x = t.empty(shape, dtype=codes[h[0]], device=D.device())
    flat = x.view(-1)
    step = (4 << 30) // max(1, flat.element_size())
    for lo in range(0, flat.numel(), step):
        dist.broadcast(flat[lo:lo + step], src, group=group)
    return x


def replicate_index(ivf, dist, group=None, src=0):
    """ONE index for a multi-GPU job (collective): rank `src` passes its built `IVF`, every other rank passes None and gets
    an `IVF` whose DEVICE copy (codes, CSR offsets, sizes, ids, centroid codes, centroids, raw vectors) and quantizer are
    rank `src`'s, bit for bit, received over NCCL -- nothing is rebuilt per rank, so the ranks of a list-sharded job cannot
    disagree about an offset. The receivers hold no host-side list views (pq_transformed_points / ids stay empty): the
    oracle checks run on the source rank, the query path only reads the device copy."""
    t = D.torch()
    rank = dist.get_rank(group)
    if rank == src:
        dev = ivf.to_device()
        meta = dict(metric=ivf.metric, n_clusters=ivf.n_clusters, dpb=ivf.pq.dims_per_block, rotate_dim=ivf.pq.rotate_dim,
                    has_R=ivf.pq.R is not None, unique_ids=bool(dev.get("unique_ids", False)),
                    scalars=[int(dev[k]) for k in _REPL_SCALARS])
        box = [meta]
    else:
        box = [None]
    dist.broadcast_object_list(box, src, group=group)
    meta = box[0]
    if rank != src:
        ivf = IVF(meta["metric"], meta["n_clusters"], FastPQ(meta["dpb"], rotate_dim=meta["rotate_dim"]))
        dev = dict(zip(_REPL_SCALARS, meta["scalars"]), unique_ids=meta["unique_ids"])
        ivf.__dict__["_dev"] = dev
    pqc = _bcast_tensor(D.upload(np.ascontiguousarray(ivf.pq.centers, dtype=np.float32)) if rank == src else None, dist, src, group)
    pqR = None
    if meta["has_R"]:
        pqR = _bcast_tensor(D.upload(np.ascontiguousarray(ivf.pq.R, dtype=np.float64)) if rank == src else None, dist, src, group)
    for name in _REPL_ARRAYS:
        dev[name] = _bcast_tensor(dev[name] if rank == src else None, dist, src, group)
    if rank != src:
        ivf.pq.centers = pqc.cpu().numpy()
        ivf.pq.R = None if pqR is None else pqR.cpu().numpy()
        ivf.pq.sqrt_n_blocks = np.sqrt(ivf.pq.centers.shape[1] // ivf.pq.dims_per_block)
        dev["host_sizes"] = dev["list_size"].cpu().numpy()
        dev["host_chunks"] = dev["list_chunk_off"].cpu().numpy()
        ivf.all_centers = ivf.active_centers = None
        ivf.data = DeviceRows(dev["data"])
    return ivf


def sync_index_from_rank0(ivf, dist, group=None):
    """Make every rank's DEVICE copy of the index equal to rank 0's (collective; NCCL broadcasts). Used when the ranks'
    independently built indexes differ: a list-sharded job needs one index. Host-side attributes (the per-list views the
    oracle reads) are left alone: only rank 0 checks against the oracle, and rank 0 is the source."""
    t = D.torch()
    dev = ivf.to_device()
    rank = dist.get_rank(group)
    scalars = ("C", "M", "n_lists", "max_chunks", "max_real_chunks", "n_chunks_total", "center_chunks")
    names = ("codes", "list_chunk_off", "list_size", "ids", "center_codes", "centers")
    hdr = t.tensor([int(dev[k]) for k in scalars] + [int(dev[n].numel()) for n in names], dtype=t.int64, device=D.device())
    dist.broadcast(hdr, 0, group=group)
    h = [int(v) for v in hdr.cpu().tolist()]
    for key, v in zip(scalars, h):
        dev[key] = v
    for n, numel in zip(names, h[len(scalars):]):
        x = dev[n].contiguous().view(-1)
        if x.numel() != numel:
            x = t.empty(numel, dtype=x.dtype, device=x.device)
        dist.broadcast(x, 0, group=group)
        dev[n] = x.view(-1, dev["d"]) if n == "centers" else x
    dev["host_sizes"] = dev["list_size"].cpu().numpy()
    dev["host_chunks"] = dev["list_chunk_off"].cpu().numpy()
    pqc = D.upload(np.ascontiguousarray(ivf.pq.centers, dtype=np.float32))
    dist.broadcast(pqc, 0, group=group)
    ivf.pq.centers = pqc.cpu().numpy()
    ivf.pq.__dict__.pop("_dev", None)
