"""World-size-2 `gloo` test of the list-sharded query path's HOST logic (runs on CPU, no CUDA):
list->rank assignment, the segment plan (numpy restatement of tkb_ivf_plan_dev), the all-gather of LUTs/probe
lists and the uneven all-to-all of estimates. The scan and the heap are played by the oracle here; the GPU
tests check the device plan/scan/replay against the same restatement."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ivf.npz")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, prefix, n_probes, k, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import restate as O
    from tinyknn_b200 import sharded as SH
    from tinyknn_b200._lib import PLAN_SEND, PLAN_RECV

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.load(GOLDEN)
        S = O.ivf_state_from_arrays(z, prefix)
        K = O.Kernels("port", "avx")
        qs = np.asarray(z[prefix + "q"], dtype=np.float32)
        Qh = len(qs) // world
        home = qs[rank * Qh:(rank + 1) * Qh]
        sizes = np.array([t[0] for t in S.pq_transformed_points], dtype=np.int32)
        owner = SH.assign_owners(sizes, world)
        C = S.pq_transformed_centers[0]
        P = min(n_probes, C)
        M = S.pq_transformed_centers[1].shape[1]

        # home phase: LUTs + probe lists of my own queries (reference semantics through the oracle)
        tabs = np.zeros((Qh, 2 * M), dtype=np.uint64)
        probes_h = np.zeros((Qh, P), dtype=np.int32)
        qn = []
        for i, q in enumerate(home):
            q = np.array(q, dtype=np.float32)
            if S.metric == "angular":
                q /= np.linalg.norm(q)
            dt = O.make_dtable(S.pq, q, K)
            tabs[i] = dt.tables[:2 * M]
            probes_h[i] = dt.top(S.pq_transformed_centers, S.active_centers, k=n_probes)
            qn.append(q)
        tables = SH.all_gather_rows(torch.from_numpy(tabs.view(np.int64)), None).numpy().view(np.uint64)
        probes = SH.all_gather_rows(torch.from_numpy(probes_h), None).numpy()
        Q = world * Qh
        assert tables.shape == (Q, 2 * M) and probes.shape == (Q, P)

        # scan of the lists I own, for every query, into the planned send buffer
        seg_s, gb_s, _ = SH.plan_host(probes, sizes, owner, PLAN_SEND, rank, world, Qh)
        seg_r, gb_r, _ = SH.plan_host(probes, sizes, owner, PLAN_RECV, rank, world, Qh)
        send = np.zeros(max(int(gb_s.sum()), 1), dtype=np.uint8)
        for q in range(Q):
            for s in range(P):
                if seg_s[q, s] < 0:
                    continue
                l = int(probes[q, s])
                assert owner[l] == rank
                n, packed = S.pq_transformed_points[l]
                est = np.zeros(2 * len(packed), dtype=np.uint64)
                K.estimate_pq(packed, np.ascontiguousarray(tables[q]), est, True)
                send[seg_s[q, s]:seg_s[q, s] + 16 * len(packed)] = est.view(np.uint8)
        recv = SH.all_to_all_bytes(torch.from_numpy(send), gb_s, gb_r, None).numpy()
        assert len(recv) == int(gb_r.sum())

        # home phase 2: ordered replay of the received estimates, exact rescoring
        bad = 0
        for i in range(Qh):
            R = (n_probes + 1) * k + 1
            hi, hv = np.zeros(R, np.int64), np.zeros(R, np.int32)
            O.init_heap(hi, hv, True)
            for s in range(P):
                l = int(probes_h[i, s])
                n = int(sizes[l])
                if n == 0:
                    continue
                nb = 16 * ((n + 15) // 16)
                assert seg_r[i, s] >= 0
                O.replay(recv[seg_r[i, s]:seg_r[i, s] + nb], n, hi, hv, True, np.ascontiguousarray(S.ids[l], dtype=np.int64))
            tr = {}
            exp = O.ivf_query(S, home[i], k, n_probes=n_probes, kernels=K, trace=tr)
            if not (np.array_equal(hi, tr["heap_indices"]) and np.array_equal(hv, tr["heap_values"])):
                bad += 1
            cand = hi[hi != -1]
            got = cand if len(cand) <= k else cand[O.bottom_k(O.exact_dists(qn[i], S.data[cand]), k)]
            if set(got) != set(exp):
                bad += 1
        np.save(out, np.array([bad, Qh, int(gb_s.sum()), int(gb_r.sum())]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("prefix,n_probes", [("euc128_", 8), ("ang_", 3)])
def test_sharded_host_logic_gloo_world2(tmp_path, prefix, n_probes):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    outs = [str(tmp_path / ("r%d.npy" % r)) for r in range(world)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, prefix, n_probes, 10, outs[r])) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    res = [np.load(o) for o in outs]
    assert all(r[0] == 0 for r in res), res                   # every home query: heap arrays and ids == single-process oracle
    assert sum(r[2] for r in res) == sum(r[3] for r in res)   # bytes sent == bytes received over the box
    assert all(r[2] > 0 for r in res)                         # both ranks own probed lists


def test_assign_owners_balanced_and_deterministic():
    from tinyknn_b200.sharded import assign_owners
    rng = np.random.default_rng(0)
    sizes = rng.integers(0, 5000, size=1000)
    for g in (1, 2, 4, 8):
        o = assign_owners(sizes, g)
        assert np.array_equal(o, assign_owners(sizes.copy(), g)) and o.min() >= 0 and o.max() < g
        load = np.bincount(o, weights=sizes, minlength=g)
        assert load.max() - load.min() <= sizes.max()


def test_plan_host_layout_properties():
    """Send layout of rank a towards home b and receive layout of b from a describe the same byte stream."""
    from tinyknn_b200.sharded import plan_host, assign_owners
    from tinyknn_b200._lib import PLAN_SEND, PLAN_RECV, PROBE_SKIP
    rng = np.random.default_rng(1)
    n_lists, G, Qh, P = 37, 4, 9, 6
    sizes = rng.integers(0, 300, size=n_lists).astype(np.int32)
    owner = assign_owners(sizes, G)
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(G * Qh)]).astype(np.int32)
    probes[3, 2] = PROBE_SKIP
    send = [plan_host(probes, sizes, owner, PLAN_SEND, r, G, Qh) for r in range(G)]
    recv = [plan_host(probes, sizes, owner, PLAN_RECV, r, G, Qh) for r in range(G)]
    for a in range(G):
        for b in range(G):
            assert send[a][1][b] == recv[b][1][a]             # split sizes agree
            # walking b's queries in (q, s) order visits the same segments at the same relative offsets
            for i in range(Qh):
                q = b * Qh + i
                for s in range(P):
                    l = probes[q, s]
                    if l == PROBE_SKIP or owner[l] != a or sizes[l] == 0:
                        continue
                    assert send[a][0][q, s] - send[a][2][b] == recv[b][0][i, s] - recv[b][2][a]


def test_plan_host_push_layout_is_the_home_ranks_single_gpu_layout():
    """TKB_PLAN_PUSH: what rank a writes for home b's queries lands exactly where b's own single-GPU plan of its block
    expects it; every segment is written by exactly one rank; the per-home buffer sizes agree."""
    from tinyknn_b200.sharded import plan_host, assign_owners
    from tinyknn_b200._lib import PLAN_SEND, PLAN_PUSH, PROBE_SKIP
    rng = np.random.default_rng(2)
    n_lists, G, Qh, P = 41, 3, 11, 5
    sizes = rng.integers(0, 300, size=n_lists).astype(np.int32)
    sizes[4] = 0
    owner = assign_owners(sizes, G)
    probes = np.stack([rng.permutation(n_lists)[:P] for _ in range(G * Qh)]).astype(np.int32)
    probes[3, 2] = PROBE_SKIP
    probes[17, 0] = -2
    push = [plan_host(probes, sizes, owner, PLAN_PUSH, r, G, Qh) for r in range(G)]
    for b in range(G):
        blk = probes[b * Qh:(b + 1) * Qh]
        home, hb, _ = plan_host(blk, sizes, None, PLAN_SEND, 0, 1, 0)
        writers = np.zeros(home.shape, dtype=np.int64)
        for a in range(G):
            assert push[a][1][b] == hb.sum()
            seg = push[a][0][b * Qh:(b + 1) * Qh]
            hit = seg >= 0
            assert np.array_equal(seg[hit], home[hit])
            writers += hit
        assert np.array_equal(writers, (home >= 0).astype(np.int64))


def _push_worker(rank, world, port, prefix, n_probes, k, shm_names, cap, out):
    """The push exchange on CPU: the "peer-mapped receive buffers" are POSIX shared-memory blocks, the scan is the oracle;
    every rank stores the segments of the lists it owns straight into the home rank's block at the offsets of
    plan_host(PLAN_PUSH), a barrier orders the stores before the replays."""
    sys.path.insert(0, ROOT)
    from multiprocessing import shared_memory
    import torch
    import torch.distributed as dist
    from oracle import restate as O
    from tinyknn_b200 import sharded as SH
    from tinyknn_b200._lib import PLAN_SEND, PLAN_PUSH

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blocks = [shared_memory.SharedMemory(name=n) for n in shm_names]
    try:
        z = np.load(GOLDEN)
        S = O.ivf_state_from_arrays(z, prefix)
        K = O.Kernels("port", "avx")
        qs = np.asarray(z[prefix + "q"], dtype=np.float32)
        Qh = len(qs) // world
        home = qs[rank * Qh:(rank + 1) * Qh]
        sizes = np.array([t[0] for t in S.pq_transformed_points], dtype=np.int32)
        owner = SH.assign_owners(sizes, world)
        P = min(n_probes, S.pq_transformed_centers[0])
        M = S.pq_transformed_centers[1].shape[1]
        tabs = np.zeros((Qh, 2 * M), dtype=np.uint64)
        probes_h = np.zeros((Qh, P), dtype=np.int32)
        qn = []
        for i, q in enumerate(home):
            q = np.array(q, dtype=np.float32)
            if S.metric == "angular":
                q /= np.linalg.norm(q)
            dt = O.make_dtable(S.pq, q, K)
            tabs[i] = dt.tables[:2 * M]
            probes_h[i] = dt.top(S.pq_transformed_centers, S.active_centers, k=n_probes)
            qn.append(q)
        tables = SH.all_gather_rows(torch.from_numpy(tabs.view(np.int64)), None).numpy().view(np.uint64)
        probes = SH.all_gather_rows(torch.from_numpy(probes_h), None).numpy()
        Q = world * Qh
        seg, gb, _ = SH.plan_host(probes, sizes, owner, PLAN_PUSH, rank, world, Qh)
        assert gb.max() <= cap
        homes = [np.ndarray((cap,), dtype=np.uint8, buffer=b.buf) for b in blocks]
        written = 0
        for q in range(Q):
            for s in range(P):
                if seg[q, s] < 0:
                    continue
                l = int(probes[q, s])
                n, packed = S.pq_transformed_points[l]
                est = np.zeros(2 * len(packed), dtype=np.uint64)
                K.estimate_pq(packed, np.ascontiguousarray(tables[q]), est, True)
                homes[q // Qh][seg[q, s]:seg[q, s] + 16 * len(packed)] = est.view(np.uint8)     # the "peer store"
                written += 16 * len(packed)
        dist.barrier()                                                   # every rank's scan has finished
        seg_r, gb_r, _ = SH.plan_host(probes_h, sizes, None, PLAN_SEND, 0, 1, 0)                  # my own single-rank layout
        mine = homes[rank]
        bad = 0
        for i in range(Qh):
            R = (n_probes + 1) * k + 1
            hi, hv = np.zeros(R, np.int64), np.zeros(R, np.int32)
            O.init_heap(hi, hv, True)
            for s in range(P):
                l = int(probes_h[i, s])
                n = int(sizes[l])
                if n == 0:
                    continue
                nb = 16 * ((n + 15) // 16)
                O.replay(np.array(mine[seg_r[i, s]:seg_r[i, s] + nb]), n, hi, hv, True, np.ascontiguousarray(S.ids[l], dtype=np.int64))
            tr = {}
            exp = O.ivf_query(S, home[i], k, n_probes=n_probes, kernels=K, trace=tr)
            if not (np.array_equal(hi, tr["heap_indices"]) and np.array_equal(hv, tr["heap_values"])):
                bad += 1
        np.save(out, np.array([bad, written, int(gb_r.sum())]))
        dist.barrier()
    finally:
        del homes, mine
        for b in blocks:
            b.close()
        dist.destroy_process_group()


def test_push_exchange_host_logic_gloo_world2(tmp_path):
    from multiprocessing import shared_memory
    import torch.multiprocessing as mp
    world, port, cap = 2, _free_port(), 4 << 20
    blocks = [shared_memory.SharedMemory(create=True, size=cap) for _ in range(world)]
    try:
        outs = [str(tmp_path / ("p%d.npy" % r)) for r in range(world)]
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_push_worker, args=(r, world, port, "euc128_", 8, 10, [b.name for b in blocks], cap, outs[r]))
                 for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        res = [np.load(o) for o in outs]
        assert all(r[0] == 0 for r in res), res                              # heaps == the single-process oracle on every rank
        assert sum(r[1] for r in res) == sum(r[2] for r in res)             # bytes stored == bytes the home layouts hold
        assert all(r[1] > 0 for r in res)
    finally:
        for b in blocks:
            b.close()
            b.unlink()


# ---- the REAL list-sharded path (ShardedIVF + the kernels' sources) on the CPU emulator, two processes over gloo -----------------

def _emu_worker(rank, world, port, out):
    """tests/test_sharded_gpu.py's worker with gloo for NCCL and the emulated library for the GPU (tests/emulate): both
    exchanges -- the all-to-all and the push exchange, where the scan kernel of one process stores into the other's receive
    buffer (emulated CUDA IPC = POSIX shared memory) --, chunk minima inside the push exchange, and the broadcast that gives
    every rank rank 0's index."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emulate"))
    import torch.distributed as dist
    import emu_torch
    emu_torch.install()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tinyknn_b200 import synth, sharded as SH, ivf as ivf_mod
        from tinyknn_b200.sharded import ShardedIVF
        X = synth.clustered(12_000 + 256, 32, 40, seed=3)
        ivf = synth.build_ivf(X[:12_000], "euclidean", 24, seed=3, deterministic=True)
        qs = X[12_000:].contiguous()
        Qh = 256 // world
        mine = qs[rank * Qh:(rank + 1) * Qh].contiguous()
        bad = 0
        bad += not synth.index_consistent(ivf, dist)                   # same seed, deterministic build: identical indexes
        ref = ivf.query_batch(mine, 10, n_probes=5, order="device", return_distances=True)
        sh = ShardedIVF(ivf)
        used = []
        for exchange in ("nccl", "push", "push", "push"):              # repeated pushes alternate the receive buffers
            got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange=exchange)
            bad += sum(not np.array_equal(a, b) for a, b in zip(ref, got))
            used.append(sh.last_exchange)
        bad += used != ["nccl", "push", "push", "push"]
        SH.PUSH_CMIN, ivf_mod.CMIN_CHUNKS = True, 1                    # chunk minima travel with the pushed estimates
        for _ in range(2):
            got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange="push")
            bad += sum(not np.array_equal(a, b) for a, b in zip(ref, got))
        # pull exchange: the estimates stay in the owner's buffer, the home rank's replay reads the other process's memory;
        # with chunk minima (copied to the home rank first) and without, device results without a host copy, and an owner
        # buffer that is too small (the guard leaves segments out, every rank repeats the batch through the push exchange)
        for cm in (1, 0, 1):
            ivf_mod.CMIN_CHUNKS = cm
            got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange="pull")
            bad += sum(not np.array_equal(a, b) for a, b in zip(ref, got))
            bad += sh.last_exchange != "pull"
        got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange="pull", to_host=False)
        sh.check_overflow()
        bad += sum(not np.array_equal(a, b.cpu().numpy()) for a, b in zip(ref, got))
        sh.pull_capacity = 4096
        got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange="pull")
        bad += sum(not np.array_equal(a, b) for a, b in zip(ref, got))
        bad += sh.__dict__.get("pull_overflows", 0) != 1 or sh.last_exchange != "push"
        try:
            sh.query_batch(mine, 10, n_probes=5, exchange="pull", to_host=False)
            sh.check_overflow()
            bad += 1                                                   # the deferred check must raise
        except RuntimeError:
            pass
        sh.pull_capacity = None
        ivf_mod.CMIN_CHUNKS = 1
        sh.close()
        # one index for the whole job: rank 1 loses its copy, the fingerprints differ, rank 0's index is broadcast
        dev = ivf.to_device()
        if rank == 1:
            dev["codes"].zero_()
            dev["centers"].mul_(0.5)
            dev["ids"].add_(7)
        bad += synth.index_consistent(ivf, dist)
        synth.sync_index_from_rank0(ivf, dist)
        bad += not synth.index_consistent(ivf, dist)
        sh = ShardedIVF(ivf)
        got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange="nccl")
        bad += sum(not np.array_equal(a, b) for a, b in zip(ref, got))
        sh.close()
        # an index built ONCE: rank 0 passes its index, rank 1 passes nothing and receives the device copy (bench.py's N > 1 path)
        ivf2 = synth.replicate_index(ivf if rank == 0 else None, dist)
        bad += not synth.index_consistent(ivf2, dist)
        bad += (rank == 1 and ivf2 is ivf)
        ref2 = ivf2.query_batch(mine, 10, n_probes=5, order="device", return_distances=True)
        bad += sum(not np.array_equal(a, b) for a, b in zip(ref, ref2))
        sh = ShardedIVF(ivf2)
        got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange="push")
        bad += sum(not np.array_equal(a, b) for a, b in zip(ref, got))
        # dropping the unsharded copy really drops it: the wrapped index cannot answer on its own any more, the shard can
        sh.drop_full_codes()
        bad += ivf2.to_device()["codes"] is not None or sh.dev["codes"] is not None
        try:
            ivf2.query_batch(mine, 10, n_probes=5, order="device")
            bad += 1
        except Exception:                                              # noqa: BLE001  (TinyKnnError: null pointer)
            pass
        got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange="nccl")
        bad += sum(not np.array_equal(a, b) for a, b in zip(ref, got))
        sh.close()
        # a host-built index is sharded from its host arrays: a rank uploads the codes of its own lists only
        if rank == 0:
            ivf.__dict__.pop("_dev", None)
            full_bytes = int(ivf._build_device()["codes"].numel())
        box = [ivf if rank == 0 else None]
        dist.broadcast_object_list(box, 0)                             # pickled host-side index (device copy excluded)
        ivf3 = box[0]
        ivf3.data = np.ascontiguousarray(X[:12_000].cpu().numpy())
        ivf3.__dict__.pop("_dev", None)
        sh = ShardedIVF(ivf3)
        bad += "_dev" in ivf3.__dict__ or sh.dev["codes"] is not None
        if rank == 0:
            bad += not (0 < int(sh.dev["local_codes"].numel()) < full_bytes)
        got = sh.query_batch(mine, 10, n_probes=5, return_distances=True, exchange="nccl")
        bad += sum(not np.array_equal(a, b) for a, b in zip(ref, got))
        sh.close()
        np.save(out, np.array([bad]))
    finally:
        dist.destroy_process_group()


def test_sharded_index_on_the_emulator_gloo_world2(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emulate"))
    import emu_lib
    try:
        emu_lib.load()                                              # build once, before the workers start
    except Exception as e:                                           # noqa: BLE001
        pytest.skip("cannot build the emulated library: %s" % str(e)[-300:])
    import torch.multiprocessing as mp
    port = _free_port()
    outs = [str(tmp_path / ("e%d.npy" % r)) for r in range(2)]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_emu_worker, args=(r, 2, port, outs[r])) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(900)
        assert p.exitcode == 0
    assert all(int(np.load(o)[0]) == 0 for o in outs)
