#!/usr/bin/env python3
"""bench.py -- IVF-PQ queries/s (fixed n_probes, k=10) + PQ-scan roofline on B200, reference CPU arm beside it.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ivf100m|glove|sift|ivf10m] [--n-probes P]

One "step" = one pass of the query hot path over one batch of `--queries` synthetic queries against a
resident index: LUT build -> centroid scan -> heap replay -> probe selection -> inverted-list scan ->
ordered heap replay -> exact rescoring -> top-k. Prints ONE JSON line (rank 0).

Default workload = BASELINE.json configs[4], the north star's target: IVF euclidean 100M x 128, 16384 lists, n_probes=32,
k=10 -- 1.6 GB of PQ codes, streamed from HBM every step; with --gpus N the inverted lists are sharded over the N ranks
(the index is built ONCE, on rank 0, and replicated over NCCL). `--workload glove` is configs[1] (the shape the reference's
published q/s are quoted on; its 31 MB of codes live in L2), `sift` configs[2].
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ORDER = "avx"

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the reference's published q/s are quoted on
    "glove": dict(n=1_183_514, d=100, metric="angular", n_clusters=1087, components=2000,
                  name="IVF angular, GloVe-100 shape synthetic 1183514x100, 1087 lists"),
    # BASELINE.json configs[2]
    "sift": dict(n=1_000_000, d=128, metric="euclidean", n_clusters=1024, components=2000,
                 name="IVF euclidean, SIFT-1M shape synthetic 1000000x128, 1024 lists"),
    # BASELINE.json configs[4] (one GPU holds the whole index; `--gpus N` shards its lists) and a 10M stand-in
    "ivf100m": dict(n=100_000_000, d=128, metric="euclidean", n_clusters=16384, components=50_000,
                    name="IVF euclidean 100M x 128 synthetic, 16384 lists"),
    "ivf10m": dict(n=10_000_000, d=128, metric="euclidean", n_clusters=4096, components=20_000,
                   name="IVF euclidean 10M x 128 synthetic, 4096 lists"),
    "tiny": dict(n=60_000, d=100, metric="angular", n_clusters=128, components=200,
                 name="IVF angular tiny (CI)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ivf100m", choices=list(WORKLOADS))
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--n-probes", type=int, default=None,
                    help="default: 32 for the 10M/100M workloads, 10 otherwise")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--order", default="avx", choices=["avx", "sse"],
                    help="accumulation order of the 4-bit scan: the reference's default AVX2 build or its SSE build")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--cpu-worker", nargs=4, metavar=("DIR", "OUT", "SPEC", "SLICE"), default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parity-queries", type=int, default=2000,
                    help="queries of the first batch checked against the oracle before anything is timed (LUT bytes, probe "
                         "lists, heap arrays, id sets; reference selection order and the device order)")
    ap.add_argument("--recall-queries", type=int, default=500,
                    help="recall@k of the timed mode against exact brute force, on this many queries of the first batch (0: skip)")
    ap.add_argument("--cpu-queries", type=int, default=2048,
                    help="workloads whose raw vectors stay on the GPU (10M/100M): the CPU arm cycles through this many queries of "
                         "the batch; the rows their rescoring reads are recorded in an untimed pass of the oracle and kept on the host")
    ap.add_argument("--exchange", default=os.environ.get("TKB_EXCHANGE", "pull"), choices=["push", "pull", "nccl"],
                    help="--shard lists: 'push' = the scan kernel stores estimates into the home rank's HBM over NVLink peer "
                         "memory, 'pull' = estimates stay in the owner's HBM and the home rank's replay fetches chunk minima and "
                         "candidate chunks over NVLink (both fall back to nccl when peer buffers cannot be mapped), "
                         "'nccl' = send buffer + all-to-all")
    ap.add_argument("--no-e2e-pipeline", action="store_true",
                    help="skip e2e_pipelined (the same host-buffer loop with two batches in flight, query_batch(to_host='async'): "
                         "batch i+1 is submitted before batch i is collected)")
    ap.add_argument("--e2e-pipeline", action="store_true", help=argparse.SUPPRESS)      # old spelling: now the default
    ap.add_argument("--graph", action="store_true",
                    help="replay one captured CUDA graph per step (IVF.graphed) instead of launching the kernels one by one; "
                         "single GPU / replicas only. Opt-in: not yet validated on hardware")
    ap.add_argument("--shard", default="auto", choices=["auto", "lists", "replicas"],
                    help="N>1: 'lists' = inverted lists sharded over the ranks, estimates stored into the query's home rank "
                         "(north star; what a 100M-vector index needs for aggregate bandwidth); 'replicas' = every rank holds "
                         "the whole index and answers its own queries, no data-path collective; 'auto' = lists when the PQ "
                         "codes of the workload are >= 1 GB, else replicas (a 31 MB index gains nothing from being split: "
                         "measured 8 GPUs, GloVe shape: lists 24.4M q/s, DESIGN.md 6)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own Cython kernels (oracle/_ref) driven by the restated Python layer
# ------------------------------------------------------------------------------------------------------

class SparseRows:
    """`data` of an index whose raw vectors are too large for the host (100M x 128 f32 = 51 GB): the rows the CPU arm's
    rescoring reads, recorded beforehand (sorted ids + their rows). `rows[ids]` like a numpy matrix; an id that was not
    recorded is an error (the CPU arm would be rescoring something the recording pass never saw)."""

    def __init__(self, ids, rows):
        self.ids, self.rows = ids, rows
        self.shape = (int(ids[-1]) + 1 if len(ids) else 0, rows.shape[1])
        self.dtype = rows.dtype

    def __getitem__(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        pos = np.searchsorted(self.ids, idx)
        pos[pos >= len(self.ids)] = 0
        if not np.array_equal(self.ids[pos], idx):
            raise KeyError("row not recorded for the CPU arm")
        return self.rows[pos]


class RecordingRows:
    """Wraps a row source and remembers which rows were asked for."""

    def __init__(self, src):
        self.src, self.shape, self.seen = src, src.shape, []

    def __getitem__(self, idx):
        self.seen.append(np.asarray(idx, dtype=np.int64).copy())
        return self.src[idx]


def cpu_worker(dirname, out, spec, slc):
    """Runs IVF.query (reference semantics) over a slice of the query batch: W warm-up + K timed steps, each a
    bounded sample sized from a pilot so the whole run takes about `seconds`. spec = "seconds:steps:warmup"."""
    from oracle import restate as O, ref_loader
    z = {f[:-4]: np.load(os.path.join(dirname, f), mmap_mode="c") for f in os.listdir(dirname) if f.endswith(".npy")}
    S = O.ivf_state_from_arrays(z)
    if "sparse_ids" in z:
        S.data = SparseRows(np.asarray(z["sparse_ids"]), z["sparse_rows"])
    kind = "ref" if ref_loader.have_ref_kernels() else "port"
    K = O.Kernels(kind, str(z["order"]) if "order" in z else "avx")
    lo, hi = (int(x) for x in slc.split(":"))
    qs = np.array(z["queries"])[lo:hi]
    n_probes, k = int(z["n_probes"]), int(z["k"])
    seconds, steps, warmup = spec.split(":")
    seconds, steps, warmup = float(seconds), int(steps), int(warmup)

    def run(count, start):
        for i in range(count):
            O.ivf_query(S, qs[(start + i) % len(qs)], k, n_probes=n_probes, kernels=K)

    run(4, 0)
    t0 = time.perf_counter()
    run(16, 4)
    rate = 16 / (time.perf_counter() - t0)
    per_step = int(min(max(8, rate * seconds / (steps + warmup)), 50_000))
    times, pos = [], 20
    for s_ in range(warmup + steps):
        t0 = time.perf_counter()
        run(per_step, pos)
        times.append(time.perf_counter() - t0)
        pos += per_step
    json.dump(dict(per_step=per_step, times=times[warmup:], kind=kind), open(out, "w"))


def record_rows(S, queries, n_probes, k, order):
    """The rows `IVF.query`'s rescoring reads for these queries: one untimed pass of the oracle itself over a row source
    that remembers what it was asked for. Returns (sorted unique ids, their rows)."""
    from oracle import restate as O, ref_loader
    K = O.Kernels("ref" if ref_loader.have_ref_kernels() else "port", order)
    src, rec = S.data, RecordingRows(S.data)
    S.data = rec
    try:
        for q in queries:
            O.ivf_query(S, q, k, n_probes=n_probes, kernels=K)
    finally:
        S.data = src
    ids = np.unique(np.concatenate(rec.seen)) if rec.seen else np.zeros(0, np.int64)
    rows = np.concatenate([src[ids[i:i + 65536]] for i in range(0, len(ids), 65536)]) if len(ids) else np.zeros((0, src.shape[1]), np.float32)
    return ids, rows


def run_cpu_arm(ivf, queries, n_probes, k, modes, steps=3, warmup=1, cpu_queries=2048):
    """The reference's CPU path on this box's host cores. modes = [(processes, seconds), ...]: every mode starts that many
    worker processes (the reference never releases the GIL, so processes, not threads); every worker holds the whole index
    and a 1/processes slice of the queries. Returns one (cpu_baseline dict, step seconds, queries
    per step) triple per mode."""
    from oracle import restate as O
    S = O.IVFState.from_ivf(ivf)
    M = S.pq_transformed_centers[1].shape[1]
    S.pq_transformed_points = [t if t is not None else O.TransformedData(0, np.zeros((0, M), np.uint64))
                               for t in S.pq_transformed_points]
    S.ids = [i if i is not None else np.zeros(0, np.int64) for i in S.ids]
    sparse = None
    note = ""
    if not isinstance(S.data, np.ndarray):               # raw vectors live on the GPU only
        queries = queries[:max(16, min(len(queries), cpu_queries))]
        log("CPU arm: recording the rows the rescoring of %d queries reads" % len(queries))
        sparse = record_rows(S, queries, n_probes, k, ORDER)
        S.data = np.zeros((0, S.data.shape[1]), np.float32)
        note = ("; the workers cycle through the first %d queries of the batch, the %d raw rows their rescoring reads were "
                "recorded by an untimed pass of the oracle and are looked up by id (the raw vector matrix stays on the GPU)"
                % (len(queries), len(sparse[0])))
    arrs = O.ivf_state_to_arrays(S)
    arrs.update(queries=queries, n_probes=np.array(n_probes), k=np.array(k), order=np.array(ORDER))
    if sparse is not None:
        arrs.update(sparse_ids=sparse[0], sparse_rows=sparse[1])
    need = sum(np.asarray(a).nbytes for a in arrs.values())
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    if shm is not None:
        import shutil
        if shutil.disk_usage(shm).free < need * 1.25 + (1 << 28):
            shm = None
    results = []
    with tempfile.TemporaryDirectory(dir=shm) as tmp:
        for name, a in arrs.items():
            np.save(os.path.join(tmp, name + ".npy"), np.asarray(a))
        del arrs
        env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
        for cores, seconds in modes:
            cores = max(1, min(cores, len(queries) // 8))
            per = max(1, len(queries) // cores)
            procs, outs = [], []
            for w in range(cores):
                out = os.path.join(tmp, "out%d_%d.json" % (cores, w))
                outs.append(out)
                procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", tmp, out,
                                               "%g:%d:%d" % (seconds, steps, warmup),
                                               "%d:%d" % (w * per, min(len(queries), (w + 1) * per))], env=env))
            for p in procs:
                p.wait()
            res = [json.load(open(o)) for o in outs if os.path.exists(o)]
            assert res, "no CPU worker finished"
            qps = sum(r["per_step"] / float(np.mean(r["times"])) for r in res)
            per_step = sum(r["per_step"] for r in res)
            step_s = float(np.mean([np.mean(r["times"]) for r in res]))
            kind = "reference" if res[0]["kind"] == "ref" else "port"
            results.append((dict(value=qps, unit="queries/s", cores=len(res), kind=kind,
                                 sample="%d timed steps of %d queries of the same batch (%d worker processes x %d queries, ~%.2f s/step); "
                                        "IVF.query reference semantics: the reference's Cython kernels (oracle/_ref) under the numpy host "
                                        "layer restated in oracle/restate.py%s" % (steps, per_step, len(res), res[0]["per_step"], step_s, note)),
                            step_s, per_step))
    return results


# ------------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


_T0 = time.perf_counter()


def log(msg):
    """Phase timestamps on stderr (TKB_BENCH_LOG=0 silences them): where a long run spends its wall-clock, per rank."""
    if os.environ.get("TKB_BENCH_LOG", "1") != "0":
        print("[bench rank %s +%.1fs] %s" % (os.environ.get("RANK", "0"), time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


def resolve_shard(args):
    """--shard auto: lists when the PQ codes of the workload are >= 1 GB, else replicas."""
    if args.shard == "auto":
        w = WORKLOADS[args.workload]
        code_bytes = w["n"] * (-(-w["d"] // 8) * 8 // 2 if w["d"] == 100 else 32) // 2      # M/2 bytes per vector (M = 52 / 32)
        args.shard = "lists" if code_bytes >= (1 << 30) else "replicas"
    return args.shard


def build_index(args, torch, seed=10):
    from tinyknn_b200 import synth
    w = WORKLOADS[args.workload]
    X = synth.clustered(w["n"] + 4 * args.queries, w["d"], w["components"], seed, normalize=False)
    data, qpool = X[:w["n"]], X[w["n"]:]
    big = w["n"] * w["d"] * 4 > (8 << 30)                          # raw vectors stay on the GPU only (rows fetched on demand)
    big = big or os.environ.get("TKB_BENCH_DEVICE_ROWS", "0") != "0"      # dry runs of that path on a small workload
    ivf = synth.build_ivf(data, w["metric"], w["n_clusters"], seed=seed, host_data=not big)
    return ivf, qpool.cpu().numpy()


def parity_gate(ivf, queries, args, kw):
    """Before anything is timed (rank 0): `--parity-queries` queries of the first batch through the reference selection order
    (`order="numpy"`) against the oracle, stage by stage -- LUT bytes, probe lists, heap arrays (slot for slot), id sets --
    then the timed mode (`order="device"`) against (a) the oracle driven with the device's selection rule and (b) the
    UNMODIFIED reference order: (b) is the parity delta of the throughput mode."""
    from oracle import restate as O, ref_loader
    import torch
    S = O.IVFState.from_ivf(ivf)
    have_ref = ref_loader.have_ref_kernels()
    K = O.Kernels("ref" if have_ref else "port", args.order)
    ns = min(len(queries), args.parity_queries if have_ref else min(args.parity_queries, 128))
    qs = queries[:ns]
    # blocks of 500 queries: one block of `query_batch` each, so that `_last` holds the stages of exactly these queries
    ids, cnt, dids, dcnt, tables, probes, hi, hv = ([] for _ in range(8))
    for lo in range(0, ns, 500):
        blk = qs[lo:lo + 500]
        i_, c_ = ivf.query_batch(blk, order="numpy", **kw)
        last = ivf._last
        assert last["tables"].shape[0] == len(blk)
        tables.append(last["tables"].cpu().numpy().reshape(len(blk), -1))
        probes.append(last["probes"].cpu().numpy())
        hi.append(last["heap_idx"].cpu().numpy()); hv.append(last["heap_val"].cpu().numpy())
        ids.append(i_); cnt.append(c_)
        i_, c_ = ivf.query_batch(blk, order="device", **kw)
        dids.append(i_); dcnt.append(c_)
    ids, cnt, dids, dcnt, tables, probes, hi, hv = (np.concatenate(x) for x in (ids, cnt, dids, dcnt, tables, probes, hi, hv))
    fa = ivf.query_batch(qs[:256], order="device", return_distances=True, fused=True, **kw)
    fb = ivf.query_batch(qs[:256], order="device", return_distances=True, fused=False, **kw)
    bad = dict(lut=0, probes=0, heap=0, ids=0, dev_rule=0, dev_vs_ref=0)
    for i in range(ns):
        tr = {}
        exp = O.ivf_query(S, qs[i], args.k, n_probes=args.n_probes, kernels=K, trace=tr)
        bad["lut"] += int(not np.array_equal(tr["tables"].view(np.uint8), tables[i]))
        top = np.asarray(tr["top"]).astype(np.int64)
        bad["probes"] += int(not np.array_equal(top, probes[i][:len(top)].astype(np.int64)))
        bad["heap"] += int(not (np.array_equal(tr["heap_indices"], hi[i]) and np.array_equal(tr["heap_values"], hv[i])))
        bad["ids"] += int(set(ids[i][:cnt[i]]) != set(exp))
        bad["dev_vs_ref"] += int(set(dids[i][:dcnt[i]]) != set(exp))
        exd = O.ivf_query(S, qs[i], args.k, n_probes=args.n_probes, kernels=K, select=O.bottom_k_sorted)
        bad["dev_rule"] += int(set(dids[i][:dcnt[i]]) != set(exd))
    torch.cuda.synchronize()
    return dict(checked=ns, oracle="oracle/restate.py + " + ("the reference's compiled kernels (oracle/_ref)" if have_ref else "pq_oracle.c"),
                mode="order=numpy (the reference's selection order), every stage compared",
                lut_byte_mismatch=bad["lut"], probe_list_mismatch=bad["probes"], heap_array_mismatch=bad["heap"],
                id_set_mismatch=bad["ids"],
                fused_vs_staged_mismatch=int(sum(not np.array_equal(x, y) for x, y in zip(fa, fb))),
                device_order_id_set_mismatch=bad["dev_rule"],
                device_vs_reference_order_id_set_mismatch=bad["dev_vs_ref"],
                note="device_order_* : the timed mode (order=device) vs the oracle run with the device's selection rule (ascending "
                     "distance, ties by heap slot); device_vs_reference_order_* : the same ids vs the UNMODIFIED reference "
                     "(np.argpartition visiting order) -- the parity delta of the throughput mode")


def recall_at_k(ivf, queries, ids, cnt, k, torch):
    """recall@k of the returned ids against exact nearest neighbours (brute force over the raw vectors on the GPU, torch
    matmuls: ground truth only, like the reference's benchmark computes it with knn_brute, ref: examples/bench.py:76-86)."""
    dev = ivf.to_device()
    X = dev["data"]
    q = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)).to(X.device)
    if ivf.metric == "angular":
        q = q / q.norm(dim=1, keepdim=True)
    best_d = torch.full((len(q), k), float("inf"), device=X.device)
    best_i = torch.full((len(q), k), -1, dtype=torch.int64, device=X.device)
    step = 1 << 20
    for lo in range(0, X.shape[0], step):
        x = X[lo:lo + step].float()
        d2 = (x * x).sum(1)[None, :] - 2.0 * (q @ x.T)               # |x|^2 - 2 q.x  (+ |q|^2, constant per query)
        dd, ii = torch.topk(d2, min(k, d2.shape[1]), dim=1, largest=False)
        cat_d, cat_i = torch.cat([best_d, dd], 1), torch.cat([best_i, ii + lo], 1)
        sel = torch.topk(cat_d, k, dim=1, largest=False).indices
        best_d, best_i = torch.gather(cat_d, 1, sel), torch.gather(cat_i, 1, sel)
    truth = best_i.cpu().numpy()
    hit = sum(len(set(ids[i][:cnt[i]].tolist()) & set(truth[i].tolist())) for i in range(len(q)))
    return hit / float(len(q) * k)


def main():
    args = parse()
    if os.environ.get("TKB_BENCH_WATCHDOG"):            # seconds: dump every thread's Python stack periodically (where is a rank stuck?)
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["TKB_BENCH_WATCHDOG"]), repeat=True, file=sys.stderr)
    if args.cpu_worker:
        return cpu_worker(*args.cpu_worker)
    global ORDER
    ORDER = args.order
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w = WORKLOADS[args.workload]
    if args.n_probes is None:
        args.n_probes = 32 if args.workload in ("ivf100m", "ivf10m") else 10
    # `config` names the workload and nothing else: it is identical for both arms (--impl ours / reference)
    cfg = dict(workload="%s, %d queries/step, k=%d, n_probes=%d%s" % (w["name"], args.queries, args.k, args.n_probes,
                                                                      "" if args.order == "avx" else ", sse accumulation order"),
               n_probes=args.n_probes, k=args.k, queries_per_step=args.queries, order=args.order)

    if args.impl == "reference" and rank != 0:
        return 0
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: tinyknn_b200 has no CPU fallback", "impl": args.impl}))
        return 1
    torch.cuda.set_device(local)
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    dist = None
    if world > 1 and args.impl == "ours":
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=30))
        dist.barrier()

    import tinyknn_b200 as tinyknn                      # noqa: F401
    from tinyknn_b200 import _lib, synth
    tinyknn.fast_pq.set_order(args.order)
    sharded = world > 1 and args.impl == "ours" and resolve_shard(args) == "lists"
    # ONE index per job: rank 0 builds it, the other ranks receive its device copy over NCCL (synth.replicate_index)
    ivf, qpool = None, None
    if rank == 0:
        log("building the synthetic index")
        ivf, qpool = build_index(args, torch)
        log("index built")
    index_note = "built once"
    if dist is not None:
        ivf = synth.replicate_index(ivf, dist)
        from tinyknn_b200 import _device as D_
        qp = torch.from_numpy(qpool).cuda() if rank == 0 else D_.empty((4 * args.queries, w["d"]), np.float32)
        dist.broadcast(qp, 0)
        qpool = qp.cpu().numpy()
        del qp
        same = synth.index_consistent(ivf, dist)
        log("index replicated from rank 0; identical on all ranks: %s" % same)
        assert same, "the replicated index differs between ranks"
        index_note = "built once on rank 0, device copy replicated to %d ranks over NCCL (fingerprints identical)" % world
    Qn = args.queries
    batches = [np.ascontiguousarray(qpool[i * Qn:(i + 1) * Qn]) for i in range(4)]

    # ---------------- reference arm: the reference's CPU path on this box's host cores ----------------
    if args.impl == "reference":
        cores = os.cpu_count() or 1
        (cb, step_s, per_step), = run_cpu_arm(ivf, batches[0], args.n_probes, args.k, [(cores, max(4.0, args.cpu_seconds * 2))],
                                              steps=args.steps, warmup=args.warmup, cpu_queries=args.cpu_queries)
        v = cb["value"]
        print(json.dumps(dict(impl="reference", metric="IVF-PQ queries/s", value=v, unit="queries/s", n_gpus=args.gpus,
                              steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * step_s, higher_is_better=True,
                              scaling="weak", vs_baseline=None, dtype="i8", data="synthetic",
                              config=cfg, step_sample_queries=per_step, cpu_baseline=cb,
                              e2e=dict(value=v, unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                              gpu_launches=0)))
        return 0

    # ---------------- our arm ---------------------------------------------------------------------------
    dev_batches = [torch.from_numpy(b).cuda() for b in batches]
    pinned = [torch.from_numpy(b).pin_memory() for b in batches]
    kw = dict(k=args.k, n_probes=args.n_probes)
    extra = {}
    if sharded:
        # rank r keeps the codes of its lists and answers its own block of queries: a different slice of the query pool
        # per rank, the same number on every rank
        from tinyknn_b200.sharded import ShardedIVF
        engine = ShardedIVF(ivf)
        log("lists sharded")
        roll = rank * 4 * Qn // world
        dev_batches = [torch.roll(b, roll, 0) for b in dev_batches]
        pinned = [torch.roll(b, roll, 0).pin_memory() for b in pinned]
        run = lambda q, **o: engine.query_batch(q, exchange=args.exchange, **kw, **o)
    else:
        if world > 1:                                   # replicas: every rank answers its own slice of the query pool
            roll = rank * 4 * Qn // world
            dev_batches = [torch.roll(b, roll, 0) for b in dev_batches]
            pinned = [torch.roll(b, roll, 0).pin_memory() for b in pinned]
        run = lambda q, **o: ivf.query_batch(q, order="device", **kw, **o)
    eager_run, graphed = run, None
    if args.graph and not sharded:
        graphed = ivf.graphed(Qn, args.k, n_probes=args.n_probes)
        extra["cuda_graph"] = True
        run = lambda q, **o: (graphed(q, to_host=o.get("to_host", True)) if not (set(o) - {"to_host"}) else eager_run(q, **o))

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    parity, recall = None, None
    if rank == 0:
        log("parity gate")
        parity = parity_gate(ivf, batches[0], args, kw)
        log("parity gate: %s" % {k_: v for k_, v in parity.items() if k_.endswith("mismatch")})
        if args.recall_queries > 0:
            nr = min(args.recall_queries, Qn)
            r_ids, r_cnt = ivf.query_batch(batches[0][:nr], order="device", **kw)
            recall = dict(k=args.k, queries=nr, value=recall_at_k(ivf, batches[0][:nr], r_ids, r_cnt, args.k, torch),
                          truth="exact brute force over the raw vectors (torch, GPU)")
            log("recall@%d = %.4f" % (args.k, recall["value"]))
    if sharded:
        # the sharded path must return exactly what the unsharded path returns for the same queries
        a = run(dev_batches[0][:256].contiguous(), return_distances=True)
        b = ivf.query_batch(dev_batches[0][:256].contiguous(), order="device", return_distances=True, **kw)
        same = all(np.array_equal(x, y) for x, y in zip(a, b))
        flag = torch.tensor([0 if same else 1], device="cuda")
        dist.all_reduce(flag)
        if rank == 0:
            parity["sharded_vs_single_gpu_mismatching_ranks"] = int(flag.item())
        torch.cuda.synchronize()
        before = torch.cuda.memory_allocated()
        engine.drop_full_codes()                         # from here on a rank holds the codes of its own lists only
        extra["code_bytes_per_rank"] = int(engine.dev["local_codes"].numel())
        extra["full_code_bytes_freed"] = int(before - torch.cuda.memory_allocated())

    log("parity gate done")
    for i in range(args.warmup):
        run(dev_batches[i % 4], to_host=False)
    sync_all()
    log("warm-up done")
    # -- value: inputs resident in HBM
    calls0 = _lib.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    torch.cuda.profiler.start()                 # `ncu --profile-from-start off` captures exactly the timed steps
    e0.record()
    t_issue = time.perf_counter()
    for i in range(args.steps):
        run(dev_batches[i % 4], to_host=False)
    t_issue = time.perf_counter() - t_issue            # wall-clock the host needed to ISSUE the steps (nothing waited for)
    e1.record()
    sync_all()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    if sharded:
        engine.check_overflow()                        # pull exchange: no owner buffer overflowed in the timed steps (raises otherwise)
        pb_ = engine.__dict__.get("_pb")
        extra["peer_buffer_bytes"] = None if pb_ is None else int(pb_.nbytes)
        extra["push_capacity_bytes"] = int(engine.push_capacity(Qn, min(args.n_probes, ivf.to_device()["C"])))
        extra["pull_overflows"] = int(engine.__dict__.get("pull_overflows", 0))
    log("timed region done: %.3f ms/step" % (ms / args.steps))
    launches = _lib.launch_count() - calls0            # kernels launched by libtinyknn_b200.so in the timed region (counted in C)
    if graphed is not None:                            # replays do not pass through the C launch counter
        launches = graphed.launches_per_replay * args.steps
    # -- per-kernel times for the roofline: the same steps again, one stream, a CUDA-event pair around every stage
    #    (in the timed loop above the sub-batches of a step overlap on side streams, which no event pair can untangle)
    ivf.profile(True)
    one = {} if sharded else dict(sub_batches=1)
    for i in range(min(args.steps, 10)):
        eager_run(dev_batches[i % 4], to_host=False, **one)
    stages = ivf.stage_times()
    last = dict(ivf._last)
    n_prof = min(args.steps, 10)
    scan_log = list(ivf.__dict__.get("_scan_log") or [])
    blocks_per_step = max(1, len(scan_log) // n_prof)                # a batch larger than the estimate workspace runs in blocks
    scan_log = scan_log[-blocks_per_step:]
    if "tc_ws" in last:                             # one more step with the role cycle counters compiled in (they cost registers: not in the timed steps)
        os.environ["TKB_TC_CLOCKS"] = "1"
        eager_run(dev_batches[(n_prof - 1) % 4], to_host=False, **one)
        torch.cuda.synchronize()
        os.environ.pop("TKB_TC_CLOCKS")
        last["tc_ws_clk"] = ivf._last["tc_ws"][:32 + 8 * 13].clone()
    staged = None
    if "fused" in stages and not sharded:           # the stage-by-stage kernels of the same path, for the scan kernel's own roofline
        ivf.profile(True)
        for i in range(min(args.steps, 5)):
            eager_run(dev_batches[i % 4], to_host=False, sub_batches=1, fused=False)
        staged = ivf.stage_times()
    ivf.profile(False)
    # -- e2e: host (pinned) queries in, ids out, through the public API
    for i in range(min(2, args.warmup)):
        run(pinned[i % 4].numpy())
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ids_h, cnt_h = run(pinned[i % 4].numpy())
    sync_all()
    e2e_s = time.perf_counter() - t0
    e2e_pipe_s = None
    if not sharded and graphed is None and not args.no_e2e_pipeline:
        pending = None
        sync_all()
        t0 = time.perf_counter()
        for i in range(args.steps):
            nxt = eager_run(pinned[i % 4].numpy(), to_host="async")
            if pending is not None:
                pending.result()
            pending = nxt
        pending.result()
        sync_all()
        e2e_pipe_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    log("e2e done")

    tms = torch.tensor([ms, e2e_s * 1e3, (e2e_pipe_s or 0.0) * 1e3], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_pipe_ms = tms.tolist()
    total_q = Qn * args.steps * world
    value = total_q / (ms * 1e-3)
    e2e_v = total_q / (e2e_ms * 1e-3)

    if rank == 0:
        dev = ivf.to_device()
        M = dev["M"]
        # algorithmic bytes of the dominant kernel (inverted-list scan): M/2 B of codes per scanned
        # (query, vector) + 1 B estimate written; scanned vectors counted from the probe lists of the last step
        # (on this rank: with sharded lists, the segments of the lists rank 0 owns, for the queries of all ranks)
        real_chunks = (dev["host_sizes"].astype(np.int64) + 15) // 16      # the reference pads each list to 16 (not to our tiles)
        scanned, Qk = 0, 0
        for pr_, so_ in (scan_log or [(last["scan_probes"], last.get("scan_seg_off"))]):   # the blocks of ONE step
            probes = pr_.cpu().numpy().astype(np.int64)
            present = np.ones_like(probes, dtype=bool) if so_ is None else so_.cpu().numpy() >= 0
            present &= probes != -(2 ** 31)
            probes = np.where(probes < 0, probes + dev["n_lists"], probes)
            scanned += int(16 * (real_chunks[np.where(present, probes, 0)] * present).sum())
            Qk += probes.shape[0]
        per_step = lambda v: float(np.sum(v)) / max(1, len(v) // blocks_per_step)      # ms per step of a stage
        fused_run = "fused" in stages
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        try:                                # dram__bytes_read+write per launch, from the committed ncu captures
            traffic_all = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload, {})
        except (OSError, ValueError):
            traffic_all = {}
        if sharded or args.n_probes != traffic_all.get("n_probes", args.n_probes):
            traffic_all = {}                # the captures are single-GPU runs at the workload's default n_probes

        def scan_roof(ms, flagged):
            ab = scanned * (M // 2 + 1)
            return dict(kernel="ivf_scan_fast", achieved=ab / (ms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                        frac=ab / (ms * 1e-3) / 1e9 / peak, traffic=traffic_all.get("ivf_scan_tc" if "tc_ws" in last else "ivf_scan_fast"),
                        algorithmic_bytes_per_launch=ab, kernel_ms=ms, codes_per_s=scanned / (ms * 1e-3), flagged_chunks=flagged)

        def counter(ws_key):
            if ws_key == "patch_ws" and "tc_ws" in last:              # tensor-core scan: refolded (vector, query) pairs
                return None
            return int(last[ws_key][:16].cpu().numpy().view(np.int64)[0 if ws_key == "patch_ws" else 1]) if ws_key in last else None

        if fused_run:
            # one launch = LUT rows read (16*M B/query) + M/2 code bytes per scanned vector + the raw rows of the heap's
            # candidates (R*d*itemsize per query; every heap is full on this workload); the estimates (1 B per scanned
            # vector, written and read back) stay in the CTA's L2-resident scratch and are NOT counted
            k_ms = per_step(stages["fused"])
            R1 = (args.n_probes + 1) * args.k + 1
            itemsize = 4 if dev["data_dtype"] == 0 else 8
            alg_bytes = scanned * (M // 2) + Qk * 16 * M + Qk * R1 * dev["d"] * itemsize
            roof = dict(bound="hbm", kernel="ivf_fused", achieved=alg_bytes / (k_ms * 1e-3) / 1e9, peak=peak, unit="GB/s",
                        frac=alg_bytes / (k_ms * 1e-3) / 1e9 / peak, traffic=traffic_all.get("ivf_fused"),
                        algorithmic_bytes_per_launch=alg_bytes, scanned_vectors_per_launch=scanned, kernel_ms=k_ms,
                        codes_per_s=scanned / (k_ms * 1e-3), flagged_chunks=counter("fused_ws"))
            if staged:
                roof["scan_kernel_alone"] = scan_roof(per_step(staged["scan"]), None)
                roof["staged_stage_ms"] = {k: per_step(v) for k, v in staged.items()}
        else:
            roof = dict(bound="hbm", scanned_vectors_per_launch=scanned, launches_per_step=blocks_per_step,
                        **scan_roof(per_step(stages["scan"]), counter("patch_ws")))
        if "tc_ws" in last:
            hdr = last["tc_ws_clk"][:32].cpu().numpy().view(np.int32)
            roof.update(kernel="ivf_scan_tc", tc_work_items=int(hdr[0]), tc_refolded_pairs=int(hdr[2]), tc_tiles=int(hdr[3]),
                        tc_mean_group_columns=16.0 * int(hdr[4]) / max(1, int(hdr[3])),
                        tc_role_cycles_per_tile={k_: round(float(v_) / max(1, int(hdr[3])) , 1) for k_, v_ in zip(
                            ("expand_wait", "expand_work", "mma_wait_a", "mma_wait_d", "mma_issue", "epi_wait", "epi_work", "epi_barrier",
                             "epi_copy", "epi_flush", "load_wait", "load_stage", "total"),
                            last["tc_ws_clk"][32:32 + 8 * 13].cpu().numpy().view(np.int64))},
                        note="list-major tensor-core scan (tcgen05.mma kind::i8): the code bytes are read once per list and batch, "
                             "so `achieved` (algorithmic bytes / time) is not bounded by the HBM peak; see `traffic`")
        code_bytes = dev["n_chunks_total"] * M * 8 if not sharded else extra.get("code_bytes_per_rank", 0)
        roof.update(peak_source="measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                    kernel_timing="CUDA events around the launch on its stream, one stream, %d steps right after the timed region"
                                  % min(args.steps, 10),
                    stage_ms={k: per_step(v) for k, v in stages.items()},
                    # whole path: the same algorithmic bytes over the whole device-timed step
                    step_frac=scanned * (M // 2 + 1) / (ms / args.steps * 1e-3) / 1e9 / peak)
        cb = None
        if not args.no_cpu_baseline and world == 1:      # rank 0 at N=1 only
            log("CPU baseline")
            cores = os.cpu_count() or 1
            (cb, _, _), (cb1, _, _) = run_cpu_arm(ivf, batches[0], args.n_probes, args.k,
                                                  [(cores, args.cpu_seconds), (1, max(2.0, args.cpu_seconds / 2))],
                                                  cpu_queries=args.cpu_queries)
            cb["single_process"] = dict(value=cb1["value"], unit="queries/s", cores=1,
                                        note="the reference's own operating mode: one process, one thread (BASELINE.md 3, mode i)")
        parallelism = (("lists sharded over %d ranks, %s" % (world, {
            "push": "estimates stored into the home rank's HBM by the scan kernel (NVLink peer memory)",
            "pull": "estimates kept in the owner's HBM, chunk minima and candidate chunks fetched by the home rank's replay (NVLink peer memory)",
        }.get(engine.last_exchange, "NCCL all-to-all of estimates")))
                       if sharded else ("query-sharded replicas x%d" % world))
        l2 = ("estimate buffer rewritten every step and query batches rotate between steps; the codes %s (%d MB) %s"
              % ("this rank scans" if sharded else "of this workload", code_bytes >> 20,
                 "are L2-resident by nature, see DESIGN.md" if code_bytes < (100 << 20)
                 else "are larger than the 126 MB L2: every step streams them from HBM"))
        line = dict(metric="IVF-PQ queries/s", value=value, unit="queries/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="i8", data="synthetic", config=cfg, parallelism=parallelism, l2=l2, index=index_note,
                    clocks=clocks, e2e=dict(value=e2e_v, unit="queries/s", h2d_bytes_per_step=int(Qn * w["d"] * 4),
                                            d2h_bytes_per_step=int(Qn * args.k * 8 + Qn * 4)),
                    gpu_launches=int(launches), roofline=roof, cpu_baseline=cb, parity=parity, recall=recall,
                    # host-side issue time of one step on rank 0: close to ms_per_step = the loop is bound by Python/launch cost,
                    # far below = the GPU is the bottleneck
                    host_issue_ms_per_step=t_issue * 1e3 / args.steps, **extra)
        if e2e_pipe_s is not None:
            line["e2e_pipelined"] = dict(value=total_q / (e2e_pipe_ms * 1e-3), unit="queries/s", in_flight=2,
                                         note="same host buffers and copies as e2e, batch i+1 submitted before batch i is collected "
                                              "(query_batch(to_host='async'))")
        print(json.dumps(line))
    if sharded:
        engine.close()                                  # unmap / free the peer buffers (collective)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main() or 0)
