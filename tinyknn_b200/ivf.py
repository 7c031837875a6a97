"""IVF index with the reference's API (ref: tinyknn/ivf.py); the query path runs on B200.

`fit` / `build` are build-time host code (numpy + sklearn, same algorithm and RNG call order as the
reference). `query` (ref: ivf.py:106-163) is a batch of one through `query_batch`, which keeps the
whole index resident in HBM and runs, per batch of queries:

  LUT build -> scan of the PQ-encoded centroids -> exact heap replay (2*n_probes+10 candidates) ->
  exact centroid distances -> probe lists -> scan of the probed inverted lists -> ordered exact heap
  replay ((n_probes+1)*k+1 candidates) -> exact rescoring distances -> k nearest.

Selection order (`order=`): the reference picks probe lists and the final k with np.argpartition,
whose output ORDER is numpy-build/CPU specific while the visiting order of the lists changes the
heap contents.
  order="numpy"  : the two tiny selections (<= 2*n_probes+10 and <= pass_1 floats per query) are done on the
                   host exactly like the reference: numpy's own arithmetic for the exact distances of those few
                   candidates and numpy's own argpartition -- the reference's ids on the machine that runs it.
                   `IVF.query` uses this.
  order="device" : everything stays on the GPU; ties and order are resolved as "ascending distance,
                   then heap slot". One host sync per batch. This is the throughput mode.
"""
import os
from contextlib import contextmanager, nullcontext

import numpy as np

from . import _device as D
from . import fast_pq as _fp
from ._lib import lib, check, DTYPE_F32, DTYPE_F64, PROBE_SKIP, PLAN_SEND, PLAN_RECV  # noqa: F401
from .fast_pq import FastPQ, TransformedData, query_pq  # noqa: F401  (ivf.py:5 re-export)
from .utils import timer, knn_brute, knn_brute_device, group_data_by_indices, bottom_k

_WORKSPACE_BYTES = 2 << 30          # cap of the per-batch estimate buffer (queries are sub-batched); see _workspace_cap
_N_STREAMS = int(os.environ.get("TKB_STREAMS", "2"))
_SUB_QUERIES = int(os.environ.get("TKB_SUB_QUERIES", "5000"))      # target queries per sub-batch
# One kernel after probe selection (tkb_ivf_query_fused_dev). Off by default: measured slower than the stage-by-stage
# kernels at large batches (a CTA holds its scan registers while it sits in the latency-bound replay), DESIGN.md 4.6.
FUSED = os.environ.get("TKB_FUSED", "0") != "0"
# Chunk minima (tkb_ivf_scan_native_cm_dev / tkb_ivf_replay_fresh_cm_dev) when a query may scan at least this many chunks:
# the replay of long probe lists then reads 1 byte per chunk instead of 16. 0 disables.
CMIN_CHUNKS = int(os.environ.get("TKB_CMIN_CHUNKS", "8192"))
# Reuse the temporaries of a block shape across batches (per stream) instead of allocating 23 tensors per block.
WORKSPACE_REUSE = os.environ.get("TKB_WORKSPACE_REUSE", "1") != "0"
# Probe selection as one kernel (tkb_coarse_probes_dev) instead of scan / replay / gather / select. Opt-in until it has been
# timed on hardware; results are identical (tests/test_gpu_build_and_batch.py, run on the emulator).
COARSE_FUSED = os.environ.get("TKB_COARSE_FUSED", "0") != "0"
# the centroid scan of probe selection through the tensor-core kernel (one list probed by every query) when that kernel applies
COARSE_TC = os.environ.get("TKB_COARSE_TC", "1") != "0"
# IVF.build: coarse assignment on the GPU (tkb_assign_dev). Opt-in until it has been validated on hardware.
ASSIGN_DEVICE = os.environ.get("TKB_ASSIGN_DEVICE", "0") != "0"
# List-major scan on the tensor cores (tkb_ivf_scan_tc_dev; csrc/tkb_scan_tc.cu): "0" never, "1" whenever the kernel applies
# (M = 32, avx order), "auto" = when the batch puts at least TC_MIN_SHARE queries on an average list.
TC_SCAN = os.environ.get("TKB_TC_SCAN", "auto")
TC_MIN_SHARE = float(os.environ.get("TKB_TC_MIN_SHARE", "6"))
_streams = {}
_ws_cap = {}


def _side_streams(n):
    """n side streams of the current device (created once)."""
    t = D.torch()
    pool = _streams.setdefault(t.cuda.current_device(), [])
    while len(pool) < n:
        pool.append(t.cuda.Stream())
    return pool[:n]


def _workspace_cap():
    """Bytes the estimate buffer of one block of queries may reserve: it is sized for the largest list in every probe slot
    (only the planned part is touched), so a skewed 100M-vector index needs room or its batches fall apart into blocks too
    small to fill the GPU. A quarter of the free memory, at least 2 GB, at most 32 GB."""
    dev_ = D.torch().cuda.current_device()
    if dev_ not in _ws_cap:                                         # asked once per device: cudaMemGetInfo synchronises
        free, _ = D.torch().cuda.mem_get_info()
        _ws_cap[dev_] = int(min(32 << 30, max(_WORKSPACE_BYTES, free // 4)))
    return _ws_cap[dev_]


_NO_STAGE = nullcontext()


class _Workspace:
    """The temporaries of one block shape on one stream, allocated once and reused by the following batches (reuse on one
    stream is stream-ordered, hence safe): a block of `query_batch` needs 23 device buffers, and asking torch for them costs
    the host more than launching the 13 kernels -- time a synchronous caller pays in front of every batch."""
    __slots__ = ("bufs",)

    def __init__(self):
        self.bufs = {}

    def __call__(self, name, shape, np_dtype):
        x = self.bufs.get(name)
        if x is None or tuple(x.shape) != tuple(shape):
            x = self.bufs[name] = D.empty(shape, np_dtype)
        return x


def _fresh(name, shape, np_dtype):
    return D.empty(shape, np_dtype)


def _tc_possible(dev):
    """The tensor-core scan exists for this index on this device (M = 32, avx order, fast scan, sm_100)."""
    return not (TC_SCAN == "0" or dev["M"] != 32 or _fp._order() != 1 or _fp.SCAN_IMPL != "fast" or not lib.tkb_ivf_scan_tc_supported())


def _tc_applies(dev, Q, P):
    """The tensor-core scan is built for M = 32 in the avx accumulation order and pays off when several queries of the
    batch probe the same list."""
    if TC_SCAN == "0" or dev["M"] != 32 or _fp._order() != 1 or _fp.SCAN_IMPL != "fast" or not lib.tkb_ivf_scan_tc_supported():
        return False
    return TC_SCAN == "1" or Q * P >= TC_MIN_SHARE * max(1, dev["C"])


def _sub_batches(Q):
    if _N_STREAMS <= 1 or Q < 2 * _SUB_QUERIES:
        return 1
    return max(2, Q // _SUB_QUERIES)


class IVF:
    def __init__(self, metric, n_clusters, pq=None):
        assert metric in ["euclidean", "angular"]
        self.metric = metric
        self.pq = FastPQ(dims_per_block=2) if pq is None else pq
        assert self.pq.centers is None, "PQ should not be pre-fitted"
        self.pq_transformed_points = [None] * n_clusters
        self.pq_transformed_centers = [None] * n_clusters
        self.n_clusters = n_clusters
        self.ids = [None] * n_clusters

    # ------------------------------------------------------------------ build time (host) ----
    def fit(self, X, verbose=False, device=None, seed=0, max_iters=25):
        """Coarse centroids by k-means on the raw vectors, then the PQ codebooks (ref: ivf.py:19-51). device=True: both on the
        GPU (`tkb_kmeans_dev`: exact-chain assignment + fixed-point centroid sums, deterministic for a seed; `FastPQ.fit(device=
        True)`); None: the module default (TKB_FIT_DEVICE=1), else sklearn on the host like the reference."""
        assert X.shape[0] >= 1
        device = _fp.FIT_DEVICE if device is None else device
        with timer(verbose, "Fitting IVF cluster centers..."):
            if self.metric == "angular":
                X = X / np.linalg.norm(X, axis=1, keepdims=True)
            if device:
                self.all_centers = self._fit_centers_device(X, seed, max_iters)
            else:
                import sklearn.cluster
                km = sklearn.cluster.KMeans(n_clusters=self.n_clusters, n_init=1, verbose=verbose)
                self.all_centers = km.fit(X).cluster_centers_
            if self.metric == "angular":
                self.all_centers /= np.linalg.norm(self.all_centers, axis=1, keepdims=True)
        with timer(verbose, "Fitting PQ to data..."):
            self.pq.fit(X, verbose=verbose, device=device, seed=seed)
        return self

    def _fit_centers_device(self, X, seed=0, max_iters=25):
        """Lloyd's k-means of the rows of X on the GPU. Seeding: k-means++ on a host subsample for up to 1024 clusters, distinct
        random rows above. Returns f32 (n_clusters, d); `self.fit_iters` = iterations run."""
        import ctypes
        Xf = np.ascontiguousarray(X, dtype=np.float32)
        n, d = Xf.shape
        k = self.n_clusters
        assert n >= k, f"n_samples={n} should be >= n_clusters={k}."
        rng = np.random.default_rng(seed)
        if k <= 1024:
            init = _fp.kmeans_pp_init(Xf[rng.choice(n, min(n, max(4096, 8 * k)), replace=False)], k, rng)
        else:
            init = Xf[rng.choice(n, k, replace=False)]
        Xd, Cd = D.upload(Xf), D.upload(np.ascontiguousarray(init, dtype=np.float32))
        need = ctypes.c_int64(0)
        check(lib.tkb_kmeans_workspace(n, d, k, ctypes.byref(need)))
        ws = D.empty((need.value,), np.uint8)
        assign = D.empty((n,), np.int32)
        done = ctypes.c_int(0)
        check(lib.tkb_kmeans_dev(D.ptr(Xd), n, d, k, D.ptr(Cd), int(max_iters), float(np.abs(Xf).max()), D.ptr(assign),
                                 ctypes.byref(done), D.ptr(ws), ws.numel(), D.stream_ptr()))
        self.fit_iters = int(done.value)
        self._fit_assign = assign.cpu().numpy()
        return Cd.cpu().numpy()

    def build(self, X, n_probes=2, verbose=False, device=None, assign_device=None):
        """Put every point into the lists of its n_probes nearest centroids (ref: ivf.py:53-104).
        device=None: the PQ encoding of all lists runs as ONE `tkb_encode_dev` launch when a GPU is present (rows gathered
        list by list, every list padded to 16 with zero vectors like FastPQ.transform does), else list by list on the host."""
        assert n_probes <= self.n_clusters, \
            f"Can't assign points to {n_probes} clusters, as index only has {self.n_clusters}"
        self.data = data = X.copy()
        if self.metric == "angular":
            data /= np.linalg.norm(data, axis=1, keepdims=True)
        if device is None:
            device = D.torch().cuda.is_available()
        with timer(verbose, "Computing nearest clusters..."):
            if assign_device is None:
                assign_device = ASSIGN_DEVICE
            if device and assign_device and n_probes <= 2 and data.dtype in (np.float32, np.float64):
                nearest = knn_brute_device(data, self.all_centers, k=n_probes, metric=self.metric)      # tkb_assign_dev
            else:
                nearest = knn_brute(data, self.all_centers, k=n_probes, metric=self.metric)
        with timer(verbose, "PQ Transforming active centers..."):
            self.active_centers = np.ascontiguousarray(self.all_centers[np.unique(nearest)], dtype=np.float32)
            self.pq_transformed_centers = self.pq.transform(self.active_centers)
        with timer(verbose, "Transforming points..."):
            n_active = self.active_centers.shape[0]
            groups, self.ids = group_data_by_indices(data, nearest, n_active)
            if device and data.dtype in (np.float32, np.float64):
                self._encode_lists_device(data, groups)
            else:
                for i, g in enumerate(groups):
                    self.pq_transformed_points[i] = self.pq.transform(g, device=False)
        self.__dict__.pop("_dev", None)
        return self

    def _encode_lists_device(self, data, groups):
        """pq.transform of every list in one launch: position i of the output encodes data[row_index[i]], -1 = padding."""
        sizes = np.array([len(i) for i in self.ids], dtype=np.int64)
        n16 = -(-sizes // 16) * 16
        off = np.concatenate(([0], np.cumsum(n16)))
        row_index = np.full(int(off[-1]), -1, dtype=np.int64)
        for l, ids_l in enumerate(self.ids):
            row_index[off[l]:off[l] + sizes[l]] = np.asarray(ids_l, dtype=np.int64)
        if off[-1]:
            packed = self.pq.encode_device(D.upload(data), row_index=D.upload(row_index)).cpu().numpy().view(np.uint64)
        for l in range(len(self.ids)):
            if sizes[l] == 0:
                self.pq_transformed_points[l] = groups[l]            # FastPQ.transform returns empty input as is (fast_pq.py:160)
            else:
                self.pq_transformed_points[l] = TransformedData(int(sizes[l]), packed[off[l] // 16:off[l + 1] // 16])

    # ------------------------------------------------------------------ device index ---------
    def __getstate__(self):
        return {k: v for k, v in self.__dict__.items() if k not in ("_dev", "_last", "_prof", "_keep_heaps", "_scan_log", "_ws")}

    def invalidate(self):
        """Forget the device copy (call after replacing index arrays by hand)."""
        self.__dict__.pop("_dev", None)

    def to_device(self):
        """Upload the index once: all inverted lists in ONE codes array (each list padded to whole
        16-vector chunks), CSR chunk offsets, true sizes, padded ids, centroid codes, centroids, raw data."""
        dev = self.__dict__.get("_dev")
        if dev is not None:
            return dev
        dev = self.__dict__["_dev"] = self._build_device()
        return dev

    def _build_device(self, owned=None):
        """The device copy of the index. owned: optional bool mask over the lists -- only the codes of those lists are
        uploaded (`local_codes` / `local_chunk_off`, `codes` stays None): what a rank of a list-sharded job holds
        (sharded.ShardedIVF), so that an index whose codes do not fit one GPU can still be sharded."""
        D.require_cuda()
        ctd = self.pq_transformed_centers
        C = int(ctd.size)
        M = ctd.packed.shape[1]
        n_lists = len(self.pq_transformed_points)
        sizes = np.zeros(n_lists, dtype=np.int32)
        chunks = np.zeros(n_lists + 1, dtype=np.int64)
        local = np.zeros(n_lists + 1, dtype=np.int64)
        parts, id_parts = [], []
        for l, td in enumerate(self.pq_transformed_points):
            n_l, packed = (0, None) if td is None or not isinstance(td, tuple) else td
            nc = 0 if packed is None else packed.shape[0]
            nc8 = -(-nc // 8) * 8                       # every list starts on a tile (8-chunk) boundary
            mine = owned is None or bool(owned[l])
            if nc:
                assert packed.shape[1] == M and packed.dtype == np.uint64
                if mine:
                    parts.append(packed)
                    if nc8 > nc:
                        parts.append(np.zeros((nc8 - nc, M), dtype=np.uint64))
                ids_l = np.full(16 * nc8, -1, dtype=np.int64)
                ids_l[:n_l] = np.asarray(self.ids[l], dtype=np.int64)[:n_l]
                id_parts.append(ids_l)
            sizes[l] = n_l
            chunks[l + 1] = chunks[l] + nc8
            local[l + 1] = local[l] + (nc8 if mine else 0)
        codes = np.concatenate(parts) if parts else np.zeros((1, M), dtype=np.uint64)
        ids = np.concatenate(id_parts) if id_parts else np.zeros(16, dtype=np.int64)
        real = ids[ids >= 0]
        unique_ids = bool(len(np.unique(real)) == len(real))     # one list per point: the heap's label dedupe is a no-op
        data = self.data
        if not isinstance(data, np.ndarray) or data.dtype not in (np.float32, np.float64):
            data = np.ascontiguousarray(data, dtype=np.float64)
        native = D.to_native(D.upload(codes), codes.shape[0], M)
        dev = dict(
            C=C, M=M, n_lists=n_lists, max_chunks=int(np.max(np.diff(chunks))) if n_lists else 0,
            max_real_chunks=int((int(sizes.max()) + 15) // 16) if n_lists else 0,
            codes=native, n_chunks_total=int(codes.shape[0]),
            list_chunk_off=D.upload(chunks), list_size=D.upload(sizes), ids=D.upload(ids),
            center_codes=D.to_native(D.upload(ctd.packed), ctd.packed.shape[0], M), center_chunks=int(ctd.packed.shape[0]),
            centers=D.upload(np.ascontiguousarray(self.active_centers, dtype=np.float32)),
            data=D.upload(data), data_dtype=DTYPE_F32 if data.dtype == np.float32 else DTYPE_F64,
            d=int(data.shape[1]), host_sizes=sizes, host_chunks=chunks, unique_ids=unique_ids)
        if owned is not None:
            dev.update(codes=None, local_codes=native, local_chunk_off=D.upload(local), n_chunks_total=int(local[-1]))
        return dev

    @staticmethod
    def _ref_codes(dev, key, n_chunks):
        """Reference-layout copy of a native code array (generic scan A/B mode only; exercises the round trip)."""
        ck = key + "_ref"
        if ck not in dev:
            dev[ck] = D.from_native(dev[key], n_chunks, dev["M"])
        return dev[ck]

    # ------------------------------------------------------------------ profiling hooks -----
    def profile(self, enabled=True):
        """Record a CUDA-event pair around every stage of the following query_batch calls
        (on the launching stream). Read them back with `stage_times()`."""
        self.__dict__["_prof"] = {} if enabled else None
        self.__dict__["_scan_log"] = []

    def _stage(self, name):
        """Context of one stage: nothing unless profile(True) asked for a CUDA-event pair around every stage."""
        return _NO_STAGE if self.__dict__.get("_prof") is None else self._timed_stage(name)

    @contextmanager
    def _timed_stage(self, name):
        prof = self.__dict__["_prof"]
        t = D.torch()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        prof.setdefault(name, []).append((e0, e1))

    def stage_times(self):
        """{stage: [ms, ...]} for the events recorded since profile(True); synchronises."""
        D.torch().cuda.synchronize()
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in (self.__dict__.get("_prof") or {}).items()}

    # ------------------------------------------------------------------ query time -----------
    def query(self, q, k, n_probes=1, pass_1=None):
        """ref: ivf.py:106-163. Returns an unordered int64 array of at most k ids."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        assert q.ndim == 1
        ids, counts = self.query_batch(q[None], k, n_probes=n_probes, pass_1=pass_1, order="numpy")
        return ids[0][:counts[0]]

    def query_batch(self, queries, k, n_probes=1, pass_1=None, order="device", return_distances=False,
                    to_host=True, sub_batches=None, fused=None):
        """Batched IVF.query (new, additive API). queries: f32 (Q, d), host array or device tensor.
        Returns (ids, counts[, dists]): ids int64 (Q, k) padded with -1, counts int32 (Q,).
        With to_host=False (order="device" only) the results stay on the GPU as torch tensors; to_host="async" returns a
        `PendingBatch` at once (device-to-host copies into pinned memory are queued behind the kernels; `.result()` waits),
        so that a caller can submit the next batch before it collects the previous one (bench.py's `e2e_pipelined`).
        sub_batches (order="device"): split the batch over side streams (None: automatic, 1: one stream).
        fused (order="device"): run everything after probe selection as one kernel (None: module default FUSED);
        False keeps the stage-by-stage kernels (scan / replay / gather / select), same results."""
        assert order in ("device", "numpy")
        assert to_host or order == "device"
        dev = self.to_device()
        t = D.torch()
        if isinstance(queries, np.ndarray):
            queries = np.ascontiguousarray(queries, dtype=np.float32)
        Q, d = queries.shape
        assert dev["d"] == d                                            # ref: ivf.py:161
        if Q == 0:
            ddt0 = np.float32 if dev["data_dtype"] == DTYPE_F32 else np.float64
            out0 = (np.zeros((0, k), np.int64), np.zeros((0,), np.int32), np.zeros((0, k), ddt0))
            return out0 if return_distances else out0[:2]
        C = dev["C"]
        P = min(n_probes, C)                                            # ref: fast_pq.py:291
        Rc = min(2 * P + 10, C)                                         # ref: fast_pq.py:293-295
        if pass_1 is None:
            pass_1 = (n_probes + 1) * k + 1                             # ref: ivf.py:135-136
        qb = int(max(1, min(Q, _workspace_cap() // max(1, P * 16 * max(dev["max_real_chunks"], 1)))))
        # Sub-batches on alternating side streams: the latency-bound stages of one sub-batch (heap replay, row
        # gathers) overlap the issue-bound scan of the next. Only in throughput mode; order="numpy" syncs per stage.
        n_sub = 1 if order != "device" else (_sub_batches(Q) if sub_batches is None else max(1, int(sub_batches)))
        if sub_batches is None and _tc_applies(dev, Q, P):              # the list-major scan shares a list's tile among ALL the
            n_sub = 1                                                   # queries of a launch: splitting the batch doubles its work
        if n_sub > 1:
            qb = min(qb, -(-Q // n_sub))
        fused = (FUSED if fused is None else bool(fused)) and order == "device" and bool(dev.get("unique_ids", False))
        outs = []
        res = None
        if order == "device":                                           # every block writes its rows of one result
            ddt = np.float32 if dev["data_dtype"] == DTYPE_F32 else np.float64
            res = (D.empty((Q, k), np.int64), D.empty((Q,), np.int32), D.empty((Q, k), ddt))
        blk = lambda lo, hi: None if res is None else tuple(r[lo:hi] for r in res)
        if n_sub > 1:
            if isinstance(queries, np.ndarray):
                queries = t.from_numpy(queries)
                if not queries.is_pinned():
                    queries = D.upload(queries.numpy())
            cur = t.cuda.current_stream()
            pool = _side_streams(min(n_sub, _N_STREAMS))
            for s_ in pool:
                s_.wait_stream(cur)
            for i, lo in enumerate(range(0, Q, qb)):
                with t.cuda.stream(pool[i % len(pool)]):
                    qs = queries[lo:min(Q, lo + qb)]
                    if not qs.is_cuda:
                        qs = qs.to(D.device(), non_blocking=True)
                    self._query_block(dev, qs, k, P, Rc, pass_1, order, blk(lo, min(Q, lo + qb)), fused)
            for s_ in pool:
                cur.wait_stream(s_)
        else:
            for lo in range(0, Q, qb):
                qs = queries[lo:min(Q, lo + qb)]
                if isinstance(qs, np.ndarray):
                    qs = D.upload(qs)
                outs.append(self._query_block(dev, qs, k, P, Rc, pass_1, order, blk(lo, min(Q, lo + qb)), fused))
        if order == "device":
            ids, cnt, dst = res
            if isinstance(to_host, str):                                # "async": copies queued, nothing waited for
                assert to_host == "async"
                return PendingBatch(ids, cnt, dst, return_distances)
            if to_host:
                ids, cnt = ids.cpu().numpy(), cnt.cpu().numpy()
                if return_distances:                                     # not copied unless asked for
                    dst = dst.cpu().numpy()
        else:
            ids, cnt, dst = (np.concatenate([o[i] for o in outs]) for i in range(3))
        if return_distances:
            return ids, cnt, dst
        return ids, cnt

    def graphed(self, n_queries, k, n_probes=1, pass_1=None, sub_batches=None):
        """CUDA-graph version of `query_batch(order="device")` for a fixed batch shape (new, additive API): the whole
        sequence of launches of one batch is captured once and replayed per call, which removes the host-side launch
        cost that a synchronous caller pays in front of every batch (about 26 launches). Returns a `GraphedBatch`;
        `graphed(...)(queries)` gives the same results as `query_batch(queries, k, n_probes, order="device")`.
        Validated on a B200 (tests/test_gpu_build_and_batch.py); `bench.py --graph` times it."""
        return GraphedBatch(self, int(n_queries), int(k), int(n_probes), pass_1, sub_batches)

    # -- stages of one block of queries (shared with the list-sharded index, sharded.py) ----------------
    def _coarse(self, dev, lut, Q, P, Rc, order, buf=_fresh):
        """Probe selection (ref: ivf.py:131 -> fast_pq.py:284-312): scan of the PQ-encoded centroids, exact heap
        replay of 2*n_probes+10 candidates, exact centroid distances, n_probes nearest. Returns int32 (Q, P)."""
        st = D.stream_ptr()
        C, M = dev["C"], dev["M"]
        sg = 1                                                           # IVF.query hard-codes signed=True (ivf.py:138,148)
        tables, qn = lut["tables"], lut["q"]
        cc, nck = dev["center_codes"], dev["center_chunks"]
        if COARSE_FUSED and order == "device" and _fp.SCAN_IMPL == "fast" and Rc <= 1024 and nck <= 4096:
            hci, hcv = buf("hci", (Q, Rc), np.int64), buf("hcv", (Q, Rc), np.int32)
            probes = buf("probes", (Q, P), np.int32)
            with self._stage("coarse_fused"):
                check(lib.tkb_coarse_probes_dev(D.ptr(cc), nck, C, M, D.ptr(tables), Q, D.ptr(dev["centers"]), dev["d"], D.ptr(qn),
                                                Rc, P, _fp._order(), D.ptr(probes), D.ptr(hci), D.ptr(hcv), None, st))
            self._last = dict(center_heap=hci, tables=tables)
            return probes
        est_c = buf("est_c", (Q, 16 * nck), np.uint8)
        with self._stage("coarse_scan"):
            if COARSE_TC and Q >= 256 and nck >= 64 and _tc_possible(dev):
                # every query scans ALL encoded centroids: one "list" that the whole batch probes -- the list-major tensor-core
                # scan at its best (one expanded tile per 64 queries). Same bytes as tkb_estimate_native_dev.
                import ctypes
                key = ("coarse_tc", Q)
                aux = dev.get(key)
                if aux is None:
                    t = D.torch()
                    aux = dev[key] = dict(off=D.upload(np.array([0, -(-nck // 8) * 8], dtype=np.int64)),
                                          size=D.upload(np.array([C], dtype=np.int32)),
                                          probes=t.zeros((Q, 1), dtype=t.int32, device=D.device()),
                                          seg=D.upload(np.arange(Q, dtype=np.int64) * (16 * nck)))
                need = ctypes.c_int64(0)
                check(lib.tkb_ivf_scan_tc_workspace(Q, 1, 1, ctypes.byref(need)))
                tws = buf("coarse_tc_ws", (need.value,), np.uint8)
                check(lib.tkb_ivf_scan_tc_dev(D.ptr(cc), D.ptr(aux["off"]), D.ptr(aux["size"]), 1, M, D.ptr(tables),
                                              D.ptr(aux["probes"]), Q, 1, D.ptr(est_c), D.ptr(aux["seg"]), None, None, 0, nck,
                                              D.ptr(tws), tws.numel(), st))
            elif _fp.SCAN_IMPL == "fast":
                ws = buf("coarse_ws", (64,), np.uint8)
                check(lib.tkb_estimate_native_dev(D.ptr(cc), nck, M, D.ptr(tables), Q, D.ptr(est_c), 16 * nck,
                                                  _fp._order(), sg, D.ptr(ws), ws.numel(), st))
            else:
                check(lib.tkb_estimate_dev(D.ptr(self._ref_codes(dev, "center_codes", nck)), nck, M, D.ptr(tables), Q,
                                           D.ptr(est_c), 16 * nck, _fp._order(), sg, st))
        hci, hcv = buf("hci", (Q, Rc), np.int64), buf("hcv", (Q, Rc), np.int32)
        with self._stage("coarse_replay"):
            check(lib.tkb_replay_fresh_dev(D.ptr(est_c), 16 * nck, nck, C, D.ptr(hci), D.ptr(hcv), Q, Rc, sg, st))
        probes = buf("probes", (Q, P), np.int32)
        with self._stage("coarse_select"):
            if Rc <= P:
                check(lib.tkb_select_probes_dev(D.ptr(hci), None, DTYPE_F32, Q, Rc, P, D.ptr(probes), st))
            else:
                dc = buf("dc", (Q, Rc), np.float32)
                check(lib.tkb_gather_dists_dev(D.ptr(dev["centers"]), DTYPE_F32, C, dev["d"], D.ptr(qn), D.ptr(hci),
                                               Q, Rc, D.ptr(dc), st))
                if order == "device":
                    check(lib.tkb_select_probes_dev(D.ptr(hci), D.ptr(dc), DTYPE_F32, Q, Rc, P, D.ptr(probes), st))
                else:
                    # parity mode: the <= 2 n_probes + 10 exact distances per query with numpy's own arithmetic on the host
                    # (knn_brute1, utils.py:89-92), not the GPU's differently ordered sums: a near-tie must break as it does in
                    # the reference, because the visiting order of the lists decides the heap
                    hci_h, q_h = hci.cpu().numpy(), qn.cpu().numpy()
                    cen = np.asarray(self.active_centers)
                    top = np.empty((Q, P), dtype=np.int32)
                    for i in range(Q):
                        diff = cen[hci_h[i]] - q_h[i]                    # a -1 slot indexes the last centroid, like the reference
                        top[i] = hci_h[i][bottom_k(np.einsum("ij,ij->i", diff, diff), P)]
                    probes = D.upload(top)
        self._last = dict(center_heap=hci, tables=tables)
        return probes

    def _scan(self, dev, tables, probes, Q, P, est, seg_off, codes_key="codes", off_key="list_chunk_off", cmin=None,
              push_cm=None, buf=_fresh):
        """Estimates of every (query, probed list) segment present in `seg_off` (ref: the scan half of
        query_pq_*, ivf.py:142-150), written compactly into `est`."""
        st = D.stream_ptr()
        M, n_lists = dev["M"], dev["n_lists"]
        max_q_chunks = P * max(dev["max_real_chunks"], 1)
        self._last.update(scan_probes=probes, scan_seg_off=seg_off)
        if self.__dict__.get("_prof") is not None:                      # every block of a profiled batch (bench.py sums them)
            self.__dict__.setdefault("_scan_log", []).append((probes, seg_off))
        with self._stage("scan"):
            if _fp.SCAN_IMPL == "fast":
                ws = buf("scan_ws", (64,), np.uint8)                     # its first 8 bytes count the recomputed chunks
                use_tc = seg_off is not None and _tc_applies(dev, Q, P)
                if use_tc:
                    import ctypes
                    need = ctypes.c_int64(0)
                    check(lib.tkb_ivf_scan_tc_workspace(Q, P, n_lists, ctypes.byref(need)))
                    tws = buf("tc_ws", (need.value,), np.uint8)
                    check(lib.tkb_ivf_scan_tc_dev(D.ptr(dev[codes_key]), D.ptr(dev[off_key]), D.ptr(dev["list_size"]), n_lists, M,
                                                  D.ptr(tables), D.ptr(probes), Q, P, D.ptr(est), D.ptr(seg_off), D.ptr(cmin),
                                                  None if push_cm is None else D.ptr(push_cm[0]), 0 if push_cm is None else push_cm[1],
                                                  max_q_chunks, D.ptr(tws), tws.numel(), st))
                    self._last["tc_ws"] = tws
                elif push_cm is not None:                                # (per-home minima table, queries per rank)
                    check(lib.tkb_ivf_scan_native_push_cm_dev(D.ptr(dev[codes_key]), D.ptr(dev[off_key]), D.ptr(dev["list_size"]), n_lists, M,
                                                              D.ptr(tables), D.ptr(probes), Q, P, D.ptr(seg_off), D.ptr(push_cm[0]), push_cm[1],
                                                              max_q_chunks, _fp._order(), 1, D.ptr(ws), ws.numel(), st))
                elif cmin is not None:
                    check(lib.tkb_ivf_scan_native_cm_dev(D.ptr(dev[codes_key]), D.ptr(dev[off_key]), D.ptr(dev["list_size"]), n_lists, M,
                                                         D.ptr(tables), D.ptr(probes), Q, P, D.ptr(est), D.ptr(seg_off), D.ptr(cmin),
                                                         max_q_chunks, _fp._order(), 1, D.ptr(ws), ws.numel(), st))
                else:
                    check(lib.tkb_ivf_scan_native_dev(D.ptr(dev[codes_key]), D.ptr(dev[off_key]), D.ptr(dev["list_size"]), n_lists, M,
                                                      D.ptr(tables), D.ptr(probes), Q, P, D.ptr(est), 0, D.ptr(seg_off),
                                                      max_q_chunks, _fp._order(), 1, D.ptr(ws), ws.numel(), st))
                self._last["patch_ws"] = ws
            else:
                check(lib.tkb_ivf_scan_dev(D.ptr(self._ref_codes(dev, codes_key, dev["n_chunks_total"])), D.ptr(dev[off_key]),
                                           D.ptr(dev["list_size"]), n_lists, M, D.ptr(tables), D.ptr(probes), Q, P,
                                           D.ptr(est), 0, D.ptr(seg_off), max(dev["max_real_chunks"], 1), _fp._order(), 1, st))

    def _replay_rescore(self, dev, qn, probes, Q, P, k, pass_1, est, seg_off, order, out=None, cmin=None, buf=_fresh,
                        cm_seg=None, absolute=False):
        """Ordered exact heap replay over the probed lists, then exact rescoring and the k nearest
        (ref: ivf.py:137-163). `probes`/`seg_off` are the rows of these Q queries. absolute: `seg_off` holds absolute
        addresses (segments in other GPUs' peer-mapped buffers, sharded.py's pull exchange), `cm_seg` is the compact layout
        the chunk minima `cmin` are addressed by."""
        st = D.stream_ptr()
        n_lists = dev["n_lists"]
        hi_, hv_ = buf("heap_idx", (Q, pass_1), np.int64), buf("heap_val", (Q, pass_1), np.int32)
        fb = buf("fallback", (Q,), np.int32)
        with self._stage("replay"):
            if absolute:
                check(lib.tkb_ivf_replay_fresh_pull_dev(D.ptr(seg_off), D.ptr(cm_seg) if cmin is not None else None, D.ptr(cmin),
                                                        D.ptr(dev["list_chunk_off"]), D.ptr(dev["list_size"]), n_lists,
                                                        D.ptr(dev["ids"]), D.ptr(probes), Q, P, D.ptr(hi_), D.ptr(hv_), pass_1, 1,
                                                        int(dev.get("unique_ids", False)), D.ptr(fb), st))
            elif cmin is not None:
                check(lib.tkb_ivf_replay_fresh_cm_dev(D.ptr(est), D.ptr(seg_off), D.ptr(cmin), D.ptr(dev["list_chunk_off"]),
                                                      D.ptr(dev["list_size"]), n_lists, D.ptr(dev["ids"]), D.ptr(probes), Q, P,
                                                      D.ptr(hi_), D.ptr(hv_), pass_1, 1, int(dev.get("unique_ids", False)),
                                                      D.ptr(fb), st))
            else:
                check(lib.tkb_ivf_replay_fresh_dev(D.ptr(est), 0, D.ptr(seg_off), D.ptr(dev["list_chunk_off"]),
                                                   D.ptr(dev["list_size"]), n_lists, D.ptr(dev["ids"]), D.ptr(probes), Q, P,
                                                   D.ptr(hi_), D.ptr(hv_), pass_1, 1, int(dev.get("unique_ids", False)),
                                                   D.ptr(fb), st))
        ddt = np.float32 if dev["data_dtype"] == DTYPE_F32 else np.float64
        dd = buf("dd", (Q, pass_1), ddt)
        with self._stage("rescore"):
            check(lib.tkb_gather_dists_dev(D.ptr(dev["data"]), dev["data_dtype"], dev["data"].shape[0], dev["d"],
                                           D.ptr(qn), D.ptr(hi_), Q, pass_1, D.ptr(dd), st))
        self._last.update(probes=probes, heap_idx=hi_, heap_val=hv_)
        if order == "device":
            oi, oc, od = out if out is not None else (D.empty((Q, k), np.int64), D.empty((Q,), np.int32), D.empty((Q, k), ddt))
            with self._stage("select"):
                check(lib.tkb_select_topk_dev(D.ptr(hi_), D.ptr(dd), dev["data_dtype"], Q, pass_1, k,
                                              D.ptr(oi), D.ptr(od), D.ptr(oc), st))
            return oi, oc, od
        # parity mode: the exact rescoring distances with numpy's own arithmetic on the host (ref: ivf.py:161-163,
        # utils.py:89-92), so that near-ties among the <= pass_1 candidates break exactly as in the reference
        hi_h, dd_h, q_h = hi_.cpu().numpy(), dd.cpu().numpy(), qn.cpu().numpy()
        ids = np.full((Q, k), -1, dtype=np.int64)
        dst = np.full((Q, k), np.inf, dtype=ddt)
        cnt = np.zeros(Q, dtype=np.int32)
        for i in range(Q):
            keep = hi_h[i] != -1                                         # ref: ivf.py:154-155
            cand, cd = hi_h[i][keep], dd_h[i][keep]
            if len(cand) > k:                                            # ref: ivf.py:158-163
                diff = self.data[cand] - q_h[i]
                cd = np.einsum("ij,ij->i", diff, diff)
                best = bottom_k(cd, k)
                cand, cd = cand[best], cd[best]
            cnt[i] = len(cand)
            ids[i, :len(cand)], dst[i, :len(cand)] = cand, cd
        return ids, cnt, dst

    def _plan(self, dev, probes, Q, P, mode=PLAN_SEND, rank=0, n_ranks=1, q_per_rank=0, rows=None, buf=_fresh):
        """Segment offsets of the compact estimate buffer (tkb_ivf_plan_dev). Returns (seg_off, group_bytes).
        n_ranks == 1 plans every segment (list ownership only matters to the sharded modes)."""
        rows = Q if rows is None else rows
        seg_off = buf("seg_off", (rows, P), np.int64)
        gb = buf("plan_gb", (2 * n_ranks + 1,), np.int64)
        ws = buf("plan_ws", (max(rows, 1) * n_ranks,), np.int64)
        with self._stage("plan"):
            check(lib.tkb_ivf_plan_dev(D.ptr(probes), Q, P, D.ptr(dev["list_size"]),
                                       D.ptr(dev.get("list_owner")) if n_ranks > 1 else None, dev["n_lists"],
                                       mode, rank, n_ranks, q_per_rank, D.ptr(seg_off), D.ptr(gb), D.ptr(ws), 8 * ws.numel(),
                                       D.stream_ptr()))
        return seg_off, gb

    def _fused_tail(self, dev, lut, probes, Q, P, k, pass_1, out=None, want_heap=False):
        """Everything after probe selection in one kernel (ref: ivf.py:135-163): scan of the probed lists, exact heap
        replay in probe order, exact rescoring, k nearest. Returns (ids, counts, dists) device tensors."""
        import ctypes
        st = D.stream_ptr()
        M, mlc = dev["M"], max(dev["max_real_chunks"], 1)
        need = ctypes.c_int64(0)
        check(lib.tkb_ivf_query_fused_workspace(Q, P, pass_1, M, _fp._order(), dev["data_dtype"], mlc, ctypes.byref(need)))
        ws = D.empty((need.value,), np.uint8)
        ddt = np.float32 if dev["data_dtype"] == DTYPE_F32 else np.float64
        oi, oc, od = out if out is not None else (D.empty((Q, k), np.int64), D.empty((Q,), np.int32), D.empty((Q, k), ddt))
        hi_ = hv_ = None
        if want_heap:
            hi_, hv_ = D.empty((Q, pass_1), np.int64), D.empty((Q, pass_1), np.int32)
        with self._stage("fused"):
            check(lib.tkb_ivf_query_fused_dev(
                D.ptr(dev["codes"]), D.ptr(dev["list_chunk_off"]), D.ptr(dev["list_size"]), dev["n_lists"], M,
                D.ptr(lut["tables"]), D.ptr(probes), Q, P, D.ptr(dev["ids"]), D.ptr(dev["data"]), dev["data_dtype"],
                dev["data"].shape[0], dev["d"], D.ptr(lut["q"]), pass_1, k, _fp._order(), mlc,
                D.ptr(oi), D.ptr(od), D.ptr(oc), D.ptr(hi_), D.ptr(hv_), D.ptr(ws), ws.numel(), st))
        self._last.update(probes=probes, scan_probes=probes, scan_seg_off=None, fused_ws=ws, heap_idx=hi_, heap_val=hv_)
        return oi, oc, od

    def _query_block(self, dev, qs, k, P, Rc, pass_1, order, out=None, fused=False):
        Q = qs.shape[0]
        # Temporaries: reused per (stream, block shape) in throughput mode. Not while profiling (bench.py reads the probe
        # lists of every block of a step afterwards) and not in the parity mode, whose callers look at `_last`.
        buf = _fresh
        if order == "device" and WORKSPACE_REUSE and self.__dict__.get("_prof") is None:
            cache = self.__dict__.setdefault("_ws", {})
            key = (D.stream_ptr(), Q, P, Rc, k, pass_1)
            buf = cache.get(key)
            if buf is None:
                if len(cache) >= 16:
                    cache.clear()
                buf = cache[key] = _Workspace()
        # 1. LUTs (ref: ivf.py:125-128)
        with self._stage("lut"):
            lut = self.pq.distance_tables(qs, signed=True, normalize=(self.metric == "angular"), buf=buf)
        # 2. probe selection
        probes = self._coarse(dev, lut, Q, P, Rc, order, buf=buf)
        if fused:
            return self._fused_tail(dev, lut, probes, Q, P, k, pass_1, out, want_heap=bool(self.__dict__.get("_keep_heaps")))
        # 3. scan of the probed lists into a compact estimate buffer, ordered replay, rescoring
        seg_off, _ = self._plan(dev, probes, Q, P, buf=buf)
        est_bytes = Q * P * 16 * max(dev["max_real_chunks"], 1)                     # upper bound; only the planned part is touched
        big = buf if est_bytes <= (1 << 30) else _fresh     # GBs of estimates (100M-vector indexes) go back to the allocator after every block
        est = big("est", (est_bytes,), np.uint8)
        cmin = None
        if CMIN_CHUNKS > 0 and _fp.SCAN_IMPL == "fast" and P * max(dev["max_real_chunks"], 1) >= CMIN_CHUNKS:
            cmin = big("cmin", (est.numel() // 16 + 16,), np.uint8)
        self._scan(dev, lut["tables"], probes, Q, P, est, seg_off, cmin=cmin, buf=buf)
        return self._replay_rescore(dev, lut["q"], probes, Q, P, k, pass_1, est, seg_off, order, out, cmin=cmin, buf=buf)


class PendingBatch:
    """Result of `query_batch(..., to_host="async")`: the copies to pinned host memory are queued on the launching stream
    behind the batch's kernels; `result()` waits for them and returns the numpy arrays."""

    def __init__(self, ids, cnt, dst, return_distances):
        t = D.torch()
        self._dev = (ids, cnt, dst)                                     # keep the device results alive until the copies ran
        self._host = [t.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in self._dev]
        for h, x in zip(self._host, self._dev):
            h.copy_(x, non_blocking=True)
        self._done = t.cuda.Event()
        self._done.record()
        self._return_distances = return_distances

    def result(self):
        self._done.synchronize()
        ids, cnt, dst = (h.numpy() for h in self._host)
        self._dev = None
        return (ids, cnt, dst) if self._return_distances else (ids, cnt)


class GraphedBatch:
    """One captured `IVF.query_batch` of a fixed shape (see `IVF.graphed`). Static device buffers: the queries are copied
    into `q_in`, the graph is replayed, the results are read from `out` (device tensors ids (Q, k) int64, counts (Q,)
    int32, dists (Q, k)); consecutive calls therefore serialise on the stream they are issued on."""

    def __init__(self, ivf, Q, k, n_probes, pass_1=None, sub_batches=None):
        t = D.require_cuda()
        dev = ivf.to_device()
        self.ivf, self.Q, self.k = ivf, Q, k
        self.q_in = D.empty((Q, dev["d"]), np.float32)
        self.q_in.zero_()
        kw = dict(k=k, n_probes=n_probes, pass_1=pass_1, order="device", return_distances=True, to_host=False,
                  sub_batches=sub_batches, fused=False)
        from ._lib import launch_count
        global WORKSPACE_REUSE
        reuse, WORKSPACE_REUSE = WORKSPACE_REUSE, False         # the graph's temporaries must come from ITS memory pool, which lives as
        try:                                                    # long as the graph does (the per-stream workspaces can be dropped)
            warm = t.cuda.Stream()                              # capture needs a warmed-up allocator and cached device state
            warm.wait_stream(t.cuda.current_stream())
            with t.cuda.stream(warm):
                for _ in range(2):
                    ivf.query_batch(self.q_in, **kw)
            t.cuda.current_stream().wait_stream(warm)
            t.cuda.synchronize()
            self.graph = t.cuda.CUDAGraph()
            n0 = launch_count()
            with t.cuda.graph(self.graph):
                self.out = ivf.query_batch(self.q_in, **kw)
            self.launches_per_replay = launch_count() - n0      # kernels inside the graph (tkb_launch_count counts the capture once)
        finally:
            WORKSPACE_REUSE = reuse

    def __call__(self, queries, return_distances=False, to_host=True):
        t = D.torch()
        if isinstance(queries, np.ndarray):
            queries = t.from_numpy(np.ascontiguousarray(queries, dtype=np.float32))
        assert tuple(queries.shape) == tuple(self.q_in.shape), "a GraphedBatch answers batches of exactly the captured shape"
        self.q_in.copy_(queries, non_blocking=True)
        self.graph.replay()
        ids, cnt, dst = self.out
        if to_host:
            host = [t.empty(x.shape, dtype=x.dtype, pin_memory=True) for x in (ids, cnt, dst)]
            for h, x in zip(host, (ids, cnt, dst)):
                h.copy_(x, non_blocking=True)
            t.cuda.current_stream().synchronize()
            ids, cnt, dst = (h.numpy() for h in host)
        return (ids, cnt, dst) if return_distances else (ids, cnt)
