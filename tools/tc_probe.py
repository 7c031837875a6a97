"""Where does the tensor-core scan spend its time? Builds bench.py's 100M x 128 index once, plans one 10 000-query batch and times
the scan launch with phases of the kernel switched off one by one (TKB_TC_DBG bits, csrc/tkb_scan_tc.cu: 1 no copy-out, 2 copy-out
without its global stores, 4 no epilogue arithmetic, 8 no expansion, 16 no MMAs, 32 certificate failures ignored). The estimates of
those variants are wrong by design; nothing here is a bench value. Usage: python tools/tc_probe.py [--workload ivf100m] [--variants 0,1,...]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="ivf100m")
    ap.add_argument("--queries", type=int, default=10000)
    ap.add_argument("--n-probes", type=int, default=32)
    ap.add_argument("--variants", default="0,32,1,2,4,36,8,16,48,5,13,61")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import bench
    from tinyknn_b200 import _device as D
    ivf, qpool = bench.build_index(args, torch)
    dev = ivf.to_device()
    qs = D.upload(np.ascontiguousarray(qpool[:args.queries], dtype=np.float32))
    Q, P = qs.shape[0], min(args.n_probes, dev["C"])
    lut = ivf.pq.distance_tables(qs, signed=True, normalize=(ivf.metric == "angular"))
    probes = ivf._coarse(dev, lut, Q, P, min(2 * P + 10, dev["C"]), "device")
    seg, gb = ivf._plan(dev, probes, Q, P)
    total = int(gb.cpu().numpy()[1])
    est = D.empty((total + 256,), np.uint8)
    cmin = D.empty((total // 16 + 272,), np.uint8)
    for v in [int(x) for x in args.variants.split(",")]:
        if v:
            os.environ["TKB_TC_DBG"] = str(v)
        else:
            os.environ.pop("TKB_TC_DBG", None)
        ms = []
        for rep in range(args.reps + 1):
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record()
            ivf._scan(dev, lut["tables"], probes, Q, P, est, seg, cmin=cmin)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                ms.append(e0.elapsed_time(e1))
        hdr = ivf._last["tc_ws"][:32].cpu().numpy().view(np.int32)
        print(json.dumps(dict(dbg=v, scan_ms=float(np.median(ms)), tiles=int(hdr[3]), refolded=int(hdr[2]))), flush=True)


if __name__ == "__main__":
    main()
