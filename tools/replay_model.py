#!/usr/bin/env python3
"""Step 2 of the replay model (CPU only): replays the streams of tools/replay_model_streams.py with the round structure of
replay_rq2_kernel (window doubling, stale bound, queue cap, at most two records per step, an accepted insert every two steps)
and counts rounds / records / accepted inserts / steps per query, and what lock-stepping G queries in one warp costs.
Round 1 numbers (R=111, 10.9 k vectors per query): 6.9 rounds, 701 queued records, 399 accepted, 936 steps per query;
8 queries per warp: 1178 steps (+26 %), 4 per warp: +17 %; 77 % of the steps fall into the first 39 chunks of the stream."""
import numpy as np, heapq
streams = np.load("/tmp/rq/streams.npy", allow_pickle=True)
R = 111; QCAP = 4 * R
def simulate(e):
    """window-doubling rounds with a stale bound; returns per round (chunks, records, accepted, steps)."""
    n_chunks = len(e) // 16
    heap = [-127] * R          # max-heap via negatives (values only; ties irrelevant for counts)
    cursor = 0; rnd = 0; out = []
    while cursor < n_chunks:
        W = (R + 15) // 16 + 1 if rnd == 0 else max(32, cursor)
        end = min(n_chunks, cursor + W)
        bound = -heap[0]
        seg = e[16 * cursor:16 * end]
        idx = np.nonzero(seg < bound)[0]
        if len(idx) > QCAP:
            # cut at chunk boundary
            cutpos = idx[QCAP]; end = cursor + cutpos // 16
            idx = idx[idx < 16 * (end - cursor)]
        recs = seg[idx]; chunks = idx // 16
        # consume: bound frozen per chunk
        acc = 0; steps = 0; i = 0; cur_chunk = -1; frozen = 0; cooldown = 0
        nrec = len(recs)
        active_until = 0
        while i < nrec or steps < active_until:
            if cooldown == 0 and i < nrec:
                took = 0; started = False
                for _ in range(2):
                    if i >= nrec: break
                    if chunks[i] != cur_chunk:
                        cur_chunk = chunks[i]; frozen = -heap[0]
                    v = recs[i]; i += 1; took += 1
                    if v < frozen:
                        heapq.heapreplace(heap, -int(v)); acc += 1; started = True
                        active_until = max(active_until, steps + 7)
                        break
                cooldown = 1 if started else 0
            else:
                cooldown = max(0, cooldown - 1)
            steps += 1
        out.append((end - cursor, nrec, acc, steps))
        cursor = end; rnd += 1
    return out
allr = [simulate(s) for s in streams]
rounds = [len(r) for r in allr]
recs = [sum(x[1] for x in r) for r in allr]
accs = [sum(x[2] for x in r) for r in allr]
steps = [sum(x[3] for x in r) for r in allr]
print("rounds/query %.1f  records %.0f  accepted %.0f  steps %.0f" % (np.mean(rounds), np.mean(recs), np.mean(accs), np.mean(steps)))
# lockstep: a warp handles G queries: steps per round = max over its queries (rounds aligned by index)
for G in (1, 4, 8, 16):
    tot = []
    for g0 in range(0, len(allr) - G + 1, G):
        grp = allr[g0:g0 + G]; nr = max(len(r) for r in grp)
        s = 0
        for j in range(nr):
            s += max((r[j][3] if j < len(r) else 0) for r in grp)
        tot.append(s)
    print("queries in lockstep %2d: steps per warp %.0f (x%.2f of the mean single-query steps)" % (G, np.mean(tot), np.mean(tot) / np.mean(steps)))
per_round = np.zeros((20, 4))
cnt = np.zeros(20)
for r in allr:
    for j, x in enumerate(r):
        per_round[j] += x; cnt[j] += 1
for j in range(int(max(rounds))):
    print("round %2d: chunks %6.0f records %6.1f accepted %5.1f steps %6.1f  (queries %d)" % ((j,) + tuple(per_round[j] / max(cnt[j], 1)) + (cnt[j],)))
