"""List-sharded IVF over the GPUs of one box: one process per GPU, torch.distributed (NCCL over NVLink).

The reference has no multi-device story at all (ref: tinyknn/ivf.py is single-process numpy). The scan of
`IVF.query` is independent per (query, probed list) (ref: ivf.py:140-150), so the inverted lists are
partitioned over the ranks; what is NOT independent is the heap that consumes the estimates: it is
order-dependent and not a true top-R, so "local top-k per GPU + merge" would change the
returned ids. The exact scheme used here moves the *estimates* (1 byte per scanned vector, 1/16..1/26 of the
code bytes read) to the query's home rank, which replays the reference heap in probe order:

  home rank (its Q/G queries)   LUT build -> centroid scan -> heap replay -> probe lists
  all ranks                     all_gather(LUTs, probe lists)                        [NCCL, ~1 KB / query]
  every rank, ALL G*Q/G queries scan of the probed lists it OWNS into a send buffer grouped by home rank
  all ranks                     all_to_all_single(estimates), uneven splits          [NCCL, 1 B / scanned vector]
  home rank                     ordered heap replay -> exact rescoring -> k nearest

exchange="push" (the default when peer memory is available) fuses the scan with the exchange: the home rank's receive
buffer holds the segments of its queries in the single-GPU (query, probe slot) layout, every rank maps every receive
buffer (CUDA IPC, NVLink peer access) and the scan kernel of the owning rank stores each estimate chunk straight into
the home rank's HBM (tkb_ivf_plan_push_dev gives it absolute addresses). No send buffer, no all-to-all, no host sync
for split sizes: the only collective after the all-gather is a one-element all-reduce that orders "every scan has
finished" before the replays. exchange="nccl" keeps the all-to-all path.

exchange="pull" keeps the estimates where they are computed: every rank scans the lists it owns into its OWN peer-visible
buffer (local stores, the scan runs at its single-GPU speed, chunk minima next to the estimates), the ranks all-gather the
G + 1 numbers that describe their layouts (that collective is also the "every scan has finished" barrier), the home rank
copies the chunk minima of its queries (1/16 of the bytes) and its replay fetches from the owners, through the peer
mappings, only the chunks that can hold a candidate -- a few percent at 100M. 100M x 128, 8 GPUs: the pushed estimates were
9 GB per rank and step over NVLink and doubled the scan's time; the pull moves ~1 GB.

Both sides of a (scanning rank, home rank) pair order the segments by (query, probe slot), so offsets are
computed locally from replicated metadata (tkb_ivf_plan_dev) and never travel. Replicated per rank:
centroids + centroid codes, list sizes/owners, `ids`, the raw vectors for rescoring; sharded: the PQ codes.
"""
import os

import numpy as np

from . import _device as D
from ._lib import PLAN_SEND, PLAN_RECV, PLAN_PUSH, PROBE_SKIP


# ---------------------------------------------------------------------------------------------------------
# host-side logic (pure numpy / torch.distributed; exercised on CPU with gloo in tests/test_sharded_cpu.py)
# ---------------------------------------------------------------------------------------------------------

# default exchange of ShardedIVF.query_batch: "pull" (estimates stay with the owner, the home rank's replay fetches minima and
# candidate chunks over NVLink: 100M x 128 on 8 GPUs 4.30 M q/s against 1.79 M for "push"), "push" (NVLink peer stores from the
# scan kernel) or "nccl" (all-to-all)
EXCHANGE = os.environ.get("TKB_EXCHANGE", "pull")
# "pull": estimates stay in the owner's HBM, the home rank's replay fetches minima + candidate chunks over NVLink (see above)
# chunk minima inside the push exchange (the home buffer carries a minima region): the home rank's replay of long probe
# lists reads 1 byte per chunk instead of 16 (100M x 128, 2 GPUs: replay 9.9 -> ~4.5 ms per step). Validated on hardware in
# round 2 (tests/test_gpu_build_and_batch.py, tests/test_sharded_gpu.py).
PUSH_CMIN = os.environ.get("TKB_PUSH_CMIN", "1") != "0"


def assign_owners(list_sizes, n_ranks):
    """Size-balanced list -> rank map: largest list first onto the least loaded rank (ties: lowest rank).
    Deterministic, so every rank computes the same map. Returns int32 (n_lists,)."""
    sizes = np.asarray(list_sizes, dtype=np.int64)
    owner = np.zeros(len(sizes), dtype=np.int32)
    load = np.zeros(n_ranks, dtype=np.int64)
    for l in np.argsort(-sizes, kind="stable"):
        r = int(np.argmin(load))
        owner[l] = r
        load[r] += sizes[l]
    return owner


def plan_host(probes, list_size, list_owner, mode, rank, n_ranks, q_per_rank):
    """numpy restatement of tkb_ivf_plan_dev / tkb_ivf_plan_push_dev (csrc/tkb_plan.cu): (seg_off, group_bytes, group_base)."""
    probes = np.asarray(probes)
    Q, P = probes.shape
    n_lists = len(list_size)
    if mode == PLAN_RECV and n_ranks > 1:
        q_lo, q_n = rank * q_per_rank, max(0, min(q_per_rank, Q - rank * q_per_rank))
    else:
        q_lo, q_n = 0, Q
    seg_off = np.full((q_n, P), -1, dtype=np.int64)
    nbytes = np.zeros((q_n, P), dtype=np.int64)
    group = np.zeros((q_n, P), dtype=np.int64)
    mine = np.ones((q_n, P), dtype=bool)
    for i in range(q_n):
        q = q_lo + i
        for s in range(P):
            l = int(probes[q, s])
            if l == PROBE_SKIP:
                continue
            if l < 0:
                l += n_lists
            owner = 0 if list_owner is None else int(list_owner[l])
            if mode == PLAN_PUSH:
                group[i, s] = q // q_per_rank if n_ranks > 1 else 0
                mine[i, s] = list_owner is None or owner == rank
            elif mode == PLAN_SEND:
                if list_owner is not None and owner != rank:
                    continue
                group[i, s] = q // q_per_rank if n_ranks > 1 else 0
            else:
                group[i, s] = owner
            nbytes[i, s] = 16 * ((int(list_size[l]) + 15) // 16)
    group_bytes = np.array([nbytes[group == g].sum() for g in range(n_ranks)], dtype=np.int64)
    group_base = np.concatenate([[0], np.cumsum(group_bytes)[:-1]]).astype(np.int64)
    # PLAN_PUSH: offsets are relative to the start of the home rank's own buffer (the device plan adds its address)
    run = np.zeros(n_ranks, dtype=np.int64) if mode == PLAN_PUSH else group_base.copy()
    for i in range(q_n):
        for s in range(P):
            if nbytes[i, s] > 0:
                g = group[i, s]
                if mine[i, s]:
                    seg_off[i, s] = run[g]
                run[g] += nbytes[i, s]
    return seg_off, group_bytes, group_base


def all_to_all_bytes(send, send_splits, recv_splits, group=None):
    """One uneven all-to-all of uint8 buffers (torch.distributed.all_to_all_single; NCCL on GPUs, gloo on CPU).
    Falls back to pairwise send/recv on backends without all_to_all."""
    import torch
    import torch.distributed as dist
    send_splits, recv_splits = [int(x) for x in send_splits], [int(x) for x in recv_splits]
    recv = torch.empty(max(sum(recv_splits), 1), dtype=torch.uint8, device=send.device)[:sum(recv_splits)]
    send = send[:sum(send_splits)]
    try:
        dist.all_to_all_single(recv, send, recv_splits, send_splits, group=group)
    except RuntimeError:                                                     # e.g. gloo without all_to_all support
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        so = np.concatenate([[0], np.cumsum(send_splits)]).astype(np.int64)
        ro = np.concatenate([[0], np.cumsum(recv_splits)]).astype(np.int64)
        recv[ro[rank]:ro[rank + 1]] = send[so[rank]:so[rank + 1]]
        reqs = []
        for peer in range(world):
            if peer == rank:
                continue
            if send_splits[peer]:
                reqs.append(dist.isend(send[so[peer]:so[peer + 1]].contiguous(), peer, group=group))
        for peer in range(world):
            if peer != rank and recv_splits[peer]:
                buf = torch.empty(recv_splits[peer], dtype=torch.uint8, device=send.device)
                dist.recv(buf, peer, group=group)
                recv[ro[peer]:ro[peer + 1]] = buf
        for r in reqs:
            r.wait()
    return recv


def all_gather_rows(x, group=None):
    """Concatenation over ranks of equally shaped per-rank rows."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


# ---------------------------------------------------------------------------------------------------------
# peer-mapped receive buffers of the push exchange
# ---------------------------------------------------------------------------------------------------------

class _Raw:
    """A device buffer that torch does not own (cudaMalloc'ed by tkb_peer_alloc)."""

    def __init__(self, address, nbytes):
        self.address, self.nbytes = int(address), int(nbytes)

    def data_ptr(self):
        return self.address

    def numel(self):
        return self.nbytes


class PeerBuffers:
    """`n_buf` receive buffers of `nbytes` on every rank, each mapped into every other rank's process.

    Collective over `group`: the 64-byte CUDA IPC handles (tkb_peer_alloc) travel in one all-gather and are opened with
    tkb_peer_open. `bases[b]` is a device int64[world]: the address of rank g's buffer b as seen from this process.
    Consecutive batches alternate between the buffers, so a fast rank can push batch i+1 while a slow home rank is
    still replaying batch i (batch i+2 cannot start before the home rank has passed the barrier of batch i+1)."""

    def __init__(self, nbytes, group=None, rank=0, world=1, n_buf=2):
        import ctypes
        from ._lib import lib, check
        t = D.require_cuda()
        nbytes = -(-int(nbytes) // 256) * 256
        self.nbytes, self.rank, self.world, self.n_buf = nbytes, rank, world, n_buf
        self.local, self.local_cmin, self._opened = [], [], []
        handles = np.zeros((n_buf, 64), dtype=np.uint8)
        for b in range(n_buf):
            p = ctypes.c_void_p()
            h = (ctypes.c_ubyte * 64)()
            # one allocation = [estimates: nbytes][chunk minima: nbytes / 16 + 16]
            check(lib.tkb_peer_alloc(self.nbytes + self.nbytes // 16 + 16, ctypes.byref(p), h))
            self.local.append(_Raw(p.value, self.nbytes))
            self.local_cmin.append(_Raw(p.value + self.nbytes, self.nbytes // 16 + 16))
            handles[b] = np.frombuffer(h, dtype=np.uint8)
        addr = np.zeros((n_buf, world), dtype=np.int64)
        addr[:, rank] = [x.address for x in self.local]
        if world > 1:
            every = all_gather_rows(D.upload(handles)[None], group).cpu().numpy()          # (world, n_buf, 64)
            for g in range(world):
                if g == rank:
                    continue
                for b in range(n_buf):
                    p = ctypes.c_void_p()
                    hb = (ctypes.c_ubyte * 64).from_buffer_copy(every[g, b].tobytes())
                    check(lib.tkb_peer_open(hb, ctypes.byref(p)))
                    self._opened.append(p.value)
                    addr[b, g] = p.value
        self.bases = [D.upload(addr[b]) for b in range(n_buf)]
        # minima region of home rank g, addressed by the scan as table[g] + (absolute estimate address >> 4)
        self.cm_tables = [D.upload(addr[b] + self.nbytes - (addr[b] >> 4)) for b in range(n_buf)]
        self.flag = t.zeros(1, dtype=t.int32, device=D.device())
        self.turn = 0

    def close(self, group=None):
        """Collective when world > 1: every rank unmaps its peers' buffers, THEN (after a barrier) frees its own -- an
        exporter must not free memory that another process still has mapped."""
        from ._lib import lib
        D.torch().cuda.synchronize()
        for p in self._opened:
            lib.tkb_peer_close(p)
        self._opened = []
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=group)
        for x in self.local:
            lib.tkb_peer_free(x.address)
        self.local, self.local_cmin = [], []


# ---------------------------------------------------------------------------------------------------------
# the sharded index (device side)
# ---------------------------------------------------------------------------------------------------------

class ShardedIVF:
    """Wraps a built `IVF`: this rank keeps the PQ codes of the lists it owns and everything replicated.

    `query_batch(queries, k, n_probes)` is collective: every rank passes ITS OWN block of queries (the same
    number on every rank) and gets the results of that block."""

    def __init__(self, ivf, rank=None, world=None, group=None, drop_full_codes=False):
        """drop_full_codes: give up the unsharded copy of the PQ codes (see `drop_full_codes()`). An index that has no
        device copy yet is sharded from its HOST arrays: only the codes of this rank's lists are uploaded."""
        import torch.distributed as dist
        self.ivf = ivf
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        t = D.torch()
        full = ivf.__dict__.get("_dev")
        if full is None:                                                      # host-built index: upload my lists only
            sizes = np.array([0 if (td is None or not isinstance(td, tuple)) else td[0] for td in ivf.pq_transformed_points],
                             dtype=np.int64)
            self.owner = assign_owners(sizes, self.world)
            dev = ivf._build_device(owned=(self.owner == self.rank))
            dev["list_owner"] = D.upload(self.owner)
            self.dev = dev
            return
        sizes = np.asarray(full["host_sizes"], dtype=np.int64)
        chunks = np.asarray(full["host_chunks"], dtype=np.int64)              # tile-padded CSR of the full index
        self.owner = assign_owners(sizes, self.world)
        mine = np.nonzero(self.owner == self.rank)[0]
        M = full["M"]
        tile_bytes = M * 8                                                    # bytes per chunk in the native layout
        local_chunks = np.zeros(len(sizes) + 1, dtype=np.int64)
        nc = np.diff(chunks)
        local_chunks[1:] = np.cumsum(np.where(self.owner == self.rank, nc, 0))
        # native codes are tile-major and every list starts on a tile: a list is one contiguous byte range
        codes = full["codes"]
        parts = [codes[int(chunks[l]) * tile_bytes:int(chunks[l + 1]) * tile_bytes] for l in mine if chunks[l + 1] > chunks[l]]
        local_codes = t.cat(parts) if parts else D.empty((16,), np.uint8)
        dev = dict(full)
        dev.update(local_codes=local_codes, local_chunk_off=D.upload(local_chunks),
                   list_owner=D.upload(self.owner), n_chunks_total=int(local_chunks[-1]))
        self.dev = dev
        if drop_full_codes:
            self.drop_full_codes()

    def drop_full_codes(self):
        """Free the unsharded copy of the PQ codes: from here on this rank holds only the codes of its own lists (capacity
        scales with the number of ranks) and the wrapped IVF's own `query_batch` fails loudly (null code pointer)."""
        self.dev["codes"] = None
        full = self.ivf.__dict__.get("_dev")
        if full is not None:
            full["codes"] = None
            full.pop("codes_ref", None)

    # The three local phases of a batch; query_batch strings them together with the two collectives (the
    # single-GPU test drives them for every rank in turn and moves the buffers by hand).
    def _home(self, queries, n_probes):
        """LUTs and probe lists of this rank's own queries."""
        ivf, dev = self.ivf, self.dev
        if isinstance(queries, np.ndarray):
            queries = D.upload(np.ascontiguousarray(queries, dtype=np.float32))
        Qh, d = queries.shape
        assert dev["d"] == d
        P = min(n_probes, dev["C"])
        Rc = min(2 * P + 10, dev["C"])
        with ivf._stage("lut"):
            lut = ivf.pq.distance_tables(queries, signed=True, normalize=(ivf.metric == "angular"))
        probes_h = ivf._coarse(dev, lut, Qh, P, Rc, "device")
        return dict(lut=lut, probes=probes_h, Qh=Qh, P=P)

    def _scan_owned(self, tables, probes, Qh, P):
        """Scan, for ALL G*Qh queries, of the probed lists this rank owns. Returns the send buffer, its split
        sizes, and the receive-side plan of this rank's own queries (offsets + split sizes)."""
        ivf, dev, G, r = self.ivf, self.dev, self.world, self.rank
        Q = G * Qh
        seg_s, gb_s = ivf._plan(dev, probes, Q, P, PLAN_SEND, r, G, Qh)
        seg_r, gb_r = ivf._plan(dev, probes, Q, P, PLAN_RECV, r, G, Qh, rows=Qh)
        splits = D.torch().stack([gb_s[:G + 1], gb_r[:G + 1]]).cpu().numpy()   # the one host sync of the batch
        est_s = D.empty((max(int(splits[0, G]), 16),), np.uint8)
        ivf._scan(dev, tables, probes, Q, P, est_s, seg_s, codes_key="local_codes", off_key="local_chunk_off")
        return est_s, splits[0, :G], seg_r, splits[1, :G]

    def _scan_push(self, tables, probes, Qh, P, home_base, cm_table=None):
        """Fused scan + exchange: scan, for ALL G*Qh queries, of the probed lists this rank owns, every estimate chunk
        stored directly at its place in the home rank's receive buffer (`home_base`: device int64[G] of addresses
        valid in this process). Returns the per-home buffer sizes (device int64[G])."""
        ivf, dev, G, r = self.ivf, self.dev, self.world, self.rank
        Q = G * Qh
        seg = D.empty((Q, P), np.int64)
        gb = D.empty((2 * G + 1,), np.int64)
        ws = D.empty((max(Q, 1) * G,), np.int64)
        from ._lib import lib, check
        with ivf._stage("plan"):
            check(lib.tkb_ivf_plan_push_dev(D.ptr(probes), Q, P, D.ptr(dev["list_size"]), D.ptr(dev["list_owner"]), dev["n_lists"],
                                            r, G, Qh, D.ptr(home_base), D.ptr(seg), D.ptr(gb), D.ptr(ws), 8 * ws.numel(),
                                            D.stream_ptr()))
        ivf._scan(dev, tables, probes, Q, P, None, seg, codes_key="local_codes", off_key="local_chunk_off",
                  push_cm=None if cm_table is None else (cm_table, Qh))
        return gb[:G]

    def _scan_pull(self, tables, probes, Qh, P, buf_est, buf_cmin, capacity):
        """Pull exchange, owner side: scan, for ALL G*Qh queries, of the probed lists this rank owns into this rank's own
        peer-visible buffer (layout: grouped by the home rank of the query, then (q, s)), chunk minima next to it.
        Returns the device int64[G + 1] that describes the layout: (total bytes, base of every home group)."""
        ivf, dev, G, r = self.ivf, self.dev, self.world, self.rank
        Q = G * Qh
        seg = D.empty((Q, P), np.int64)
        gb = D.empty((2 * G + 1,), np.int64)
        ws = D.empty((max(Q, 1) * G,), np.int64)
        from ._lib import lib, check
        with ivf._stage("plan"):
            check(lib.tkb_ivf_plan_pull_owner_dev(D.ptr(probes), Q, P, D.ptr(dev["list_size"]), D.ptr(dev["list_owner"]),
                                                  dev["n_lists"], r, G, Qh, int(capacity), D.ptr(seg), D.ptr(gb), D.ptr(ws),
                                                  8 * ws.numel(), D.stream_ptr()))
        ivf._scan(dev, tables, probes, Q, P, buf_est, seg, codes_key="local_codes", off_key="local_chunk_off", cmin=buf_cmin)
        return gb[G:2 * G + 1]

    def _pull_home(self, probes_all, home, owner_base, owner_groups, cm_table):
        """Pull exchange, home side: where the segments of this rank's queries live in the owners' buffers (absolute
        addresses), the compact local layout the minima are copied into, and that copy. Returns (seg_addr, seg_local, cmin)."""
        ivf, dev, G, r = self.ivf, self.dev, self.world, self.rank
        Qh, P = home["Qh"], home["P"]
        from ._lib import lib, check
        seg_local, _ = ivf._plan(dev, home["probes"], Qh, P)
        seg_addr = D.empty((Qh, P), np.int64)
        gb = D.empty((2 * G + 1,), np.int64)
        ws = D.empty((max(Qh, 1) * G,), np.int64)
        with ivf._stage("plan"):
            check(lib.tkb_ivf_plan_pull_home_dev(D.ptr(probes_all), G * Qh, P, D.ptr(dev["list_size"]), D.ptr(dev["list_owner"]),
                                                 dev["n_lists"], r, G, Qh, D.ptr(owner_base), D.ptr(owner_groups), D.ptr(seg_addr),
                                                 D.ptr(gb), D.ptr(ws), 8 * ws.numel(), D.stream_ptr()))
        cmin = None
        if cm_table is not None:
            need = self.push_capacity(Qh, P) // 16 + 64
            cmin = self.__dict__.get("_pull_cmin")
            if cmin is None or cmin.numel() < need:
                self.__dict__["_pull_cmin"] = cmin = D.empty((need,), np.uint8)
            with ivf._stage("pull_minima"):
                check(lib.tkb_ivf_pull_minima_dev(D.ptr(home["probes"]), Qh, P, D.ptr(dev["list_size"]), D.ptr(dev["list_owner"]),
                                                  dev["n_lists"], D.ptr(seg_addr), D.ptr(seg_local), D.ptr(cm_table), D.ptr(cmin),
                                                  D.stream_ptr()))
        return seg_addr, seg_local, cmin

    def push_capacity(self, Qh, P):
        """Upper bound of a home rank's receive buffer, from replicated metadata only (every rank computes the same
        number): Qh queries x the P largest lists, padded to chunks."""
        sizes = np.sort(16 * ((np.asarray(self.dev["host_sizes"], dtype=np.int64) + 15) // 16))[::-1]
        return int(max(Qh, 1) * max(int(sizes[:P].sum()), 16))

    def _peers(self, Qh, P, pull=False):
        """The peer-mapped buffers for batches of this shape (collective when they have to be (re)allocated); None when they
        would not fit comfortably (then the all-to-all path is used). Push: a home rank's receive buffer, bounded by
        `push_capacity`. Pull: an owner's buffer holds its lists' segments for the queries of ALL ranks -- the same bytes on
        average, up to `world` times as many when every query probes this rank's lists; it gets what memory allows up to that
        bound and the plan's capacity guard covers the rest (an overflowing batch is repeated through the push exchange)."""
        need = self.push_capacity(Qh, P)
        pb = self.__dict__.get("_pb")
        if pb is not None and pb.nbytes >= need:
            return pb
        t = D.torch()
        t.cuda.empty_cache()                                  # the buffers come from cudaMalloc, not from torch's cache
        free, _ = t.cuda.mem_get_info()
        budget = (free // 2) // 2                             # two buffers (+ 1/16 for the minima) inside half of the free memory
        room = t.tensor([int(budget * 16 // 17)], dtype=t.int64, device=D.device())
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(room, op=dist.ReduceOp.MIN, group=self.group)
        room = int(room.item())
        if room < need:
            return None
        if pb is not None:
            pb.close(self.group)
        size = min(self.world * need, room) if pull else need
        self.__dict__["_pb"] = pb = PeerBuffers(size, self.group, self.rank, self.world)
        return pb

    def close(self):
        """Release the peer-mapped receive buffers (collective when world > 1). The index itself stays usable: the next
        push batch maps fresh buffers."""
        pb = self.__dict__.pop("_pb", None)
        if pb is not None:
            pb.close(self.group)

    def _defer_overflow_check(self, groups, nbytes):
        """to_host=False callers get device tensors and no synchronisation: the owners' totals are copied to pinned host
        memory behind the batch and looked at once the copy has completed (`check_overflow()`, also called by every later
        batch for the copies that are done by then)."""
        t = D.torch()
        host = t.empty((groups.shape[0],), dtype=t.int64).pin_memory()
        host.copy_(groups[:, 0], non_blocking=True)
        ev = t.cuda.Event()
        ev.record()
        self.__dict__.setdefault("_pending_checks", []).append((host, ev, int(nbytes)))
        self.check_overflow(wait=False)

    def check_overflow(self, wait=True):
        """Raises if a `to_host=False` pull batch left segments out because an owner's buffer was too small (its results are
        then incomplete; repeat it with exchange="push"). wait=False: only the batches whose totals have already arrived."""
        pend = self.__dict__.get("_pending_checks", [])
        keep = []
        for host, ev, nbytes in pend:
            if not wait and not ev.query():
                keep.append((host, ev, nbytes))
                continue
            ev.synchronize()
            if int(host.max()) > nbytes:
                self.__dict__["_pending_checks"] = []
                raise RuntimeError("pull exchange: an owner's estimate buffer overflowed (%d > %d bytes); the results of that "
                                   "to_host=False batch are incomplete -- repeat it with exchange='push'" % (int(host.max()), nbytes))
        self.__dict__["_pending_checks"] = keep

    def _finish(self, home, est_r, seg_r, k, pass_1, cmin_r=None, cm_seg=None):
        """Ordered heap replay over the received estimates, exact rescoring, k nearest (device tensors).
        Pull exchange: est_r None, seg_r absolute addresses, cm_seg the compact layout of the minima."""
        return self.ivf._replay_rescore(self.dev, home["lut"]["q"], home["probes"], home["Qh"], home["P"], k, pass_1,
                                        est_r, seg_r, "device", cmin=cmin_r, cm_seg=cm_seg, absolute=est_r is None)

    def query_batch(self, queries, k, n_probes=1, pass_1=None, return_distances=False, to_host=True, exchange=None):
        """Collective. queries: this rank's f32 (Qh, d) block (same Qh on every rank). Selections use the
        device order (ascending distance, ties by heap slot). Returns the results of this rank's block.
        exchange: "push" (scan stores into the home rank's HBM over NVLink peer memory), "pull" (estimates stay with the
        owner, the home rank's replay fetches minima and candidate chunks over NVLink), "nccl" (send buffer + all-to-all)
        or None = the module default EXCHANGE; every rank must pass the same value."""
        ivf, G = self.ivf, self.world
        exchange = EXCHANGE if exchange is None else exchange
        assert exchange in ("push", "pull", "nccl")
        if pass_1 is None:
            pass_1 = (n_probes + 1) * k + 1                                     # ref: ivf.py:135-136
        home = self._home(queries, n_probes)
        tables, probes = home["lut"]["tables"], home["probes"]
        Qh, P = home["Qh"], home["P"]
        pb = self._peers(Qh, P, exchange == "pull") if (exchange in ("push", "pull") and G > 1) else None
        self.last_exchange = exchange if pb is not None else ("nccl" if G > 1 else "local")
        if G > 1:
            with ivf._stage("all_gather"):
                tables = all_gather_rows(tables, self.group)
                probes = all_gather_rows(probes, self.group)
        groups = None
        if pb is not None and exchange == "pull":
            b = pb.turn
            pb.turn = (b + 1) % pb.n_buf
            from . import ivf as _ivf_mod
            use_cm = (PUSH_CMIN and _ivf_mod.CMIN_CHUNKS > 0
                      and P * max(self.dev["max_real_chunks"], 1) >= _ivf_mod.CMIN_CHUNKS)
            # bytes of the owner buffer a batch may use: all of it, unless the caller set `pull_capacity` (tests of the guard)
            cap = min(pb.nbytes, int(self.__dict__.get("pull_capacity") or pb.nbytes))
            mine = self._scan_pull(tables, probes, Qh, P, pb.local[b], pb.local_cmin[b] if use_cm else None, cap)
            with ivf._stage("barrier"):          # every rank's layout numbers; also orders "every scan kernel has completed"
                groups = all_gather_rows(mine[None], self.group)                 # (G, G + 1)
            seg_r, cm_seg, cmin_r = self._pull_home(probes, home, pb.bases[b], groups, pb.cm_tables[b] if use_cm else None)
            est_r = None
        elif pb is not None:
            import torch.distributed as dist
            b = pb.turn
            pb.turn = (b + 1) % pb.n_buf
            seg_r, _ = ivf._plan(self.dev, home["probes"], Qh, P)                # the single-GPU layout of my own queries
            from . import ivf as _ivf_mod
            use_cm = (PUSH_CMIN and _ivf_mod.CMIN_CHUNKS > 0
                      and P * max(self.dev["max_real_chunks"], 1) >= _ivf_mod.CMIN_CHUNKS)
            self._scan_push(tables, probes, Qh, P, pb.bases[b], pb.cm_tables[b] if use_cm else None)
            with ivf._stage("barrier"):                                         # every rank's scan kernel has completed
                dist.all_reduce(pb.flag, group=self.group)
            est_r = pb.local[b]
            cmin_r = pb.local_cmin[b] if use_cm else None
        else:
            est_s, send_splits, seg_r, recv_splits = self._scan_owned(tables, probes, Qh, P)
            if G > 1:
                with ivf._stage("all_to_all"):
                    est_r = all_to_all_bytes(est_s, send_splits, recv_splits, self.group)
            else:
                est_r = est_s
            cmin_r = None
        if groups is not None:
            ids, cnt, dst = self._finish(home, None, seg_r, k, pass_1, cmin_r, cm_seg)
            # an owner whose segments did not fit its buffer left them out (the plan's capacity guard): every rank sees the same
            # totals, so every rank repeats the batch through the push exchange
            if to_host:
                if int(groups[:, 0].max().item()) > cap:
                    self.pull_overflows = self.__dict__.get("pull_overflows", 0) + 1
                    return self.query_batch(queries, k, n_probes, pass_1, return_distances, to_host, "push")
            else:
                self._defer_overflow_check(groups, cap)
        else:
            ids, cnt, dst = self._finish(home, est_r, seg_r, k, pass_1, cmin_r)
        if to_host:
            ids, cnt, dst = ids.cpu().numpy(), cnt.cpu().numpy(), dst.cpu().numpy()
        return (ids, cnt, dst) if return_distances else (ids, cnt)
