"""Summaries of an `ncu --set full --import-source on` report for profiles/ (run in the build container; ncu reads the report
without a GPU). Usage: python tools/ncu_summarize.py <report.ncu-rep> <profiles/prefix> [launch index for the SASS page, default 0]
Writes <prefix>_raw.csv (every raw metric of every captured launch, one row per metric) and <prefix>_sass.txt (instruction mix by
opcode, the tcgen05 / TMEM mnemonics found, the 40 SASS instructions with the most stall samples, executed instructions per
400-byte region)."""
import collections
import csv
import io
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, launches = raw[0], raw[1], raw[2:]
    with open(prefix + "_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch %d" % i for i in range(len(launches))])
        for j, (h, u) in enumerate(zip(hdr, units)):
            w.writerow([h, u] + [l[j] for l in launches])
    launch = sys.argv[3] if len(sys.argv) > 3 else "0"
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", launch,
                                           "--launch-count", "1"]))))
    name = src[0][1] if src and len(src[0]) > 1 else "?"
    h = src[1]
    rows = []
    for r in src[2:]:                                   # (the page may repeat itself: stop at the next "Kernel Name" header)
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(h):
            rows.append(r)
    i_src, i_smp, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    tot_ex = sum(int(r[i_ex]) for r in rows)
    tot_smp = sum(int(r[i_smp]) for r in rows)
    mix = collections.Counter()
    for r in rows:
        op = r[i_src].strip().split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        mix[op.split(".")[0]] += int(r[i_ex])
    with open(prefix + "_sass.txt", "w") as f:
        f.write("kernel: %s\nSASS instructions: %d, warp instructions executed: %d, stall samples: %d\n\n" % (name, len(rows), tot_ex, tot_smp))
        f.write("tensor-core / tensor-memory / mbarrier mnemonics present (static count, executed warp instructions):\n")
        for key in ("UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "SYNCS", "UTMALDG", "ELECT"):
            hits = [r for r in rows if key in r[i_src]]
            if hits:
                f.write("  %-8s static %4d  executed %12d   e.g. %s\n" % (key, len(hits), sum(int(r[i_ex]) for r in hits), hits[0][i_src].strip()))
        f.write("\nexecuted warp instructions by opcode (share of all):\n")
        for op, n in mix.most_common(30):
            f.write("  %-12s %6.2f %%\n" % (op, 100.0 * n / max(1, tot_ex)))
        f.write("\n40 instructions with the most stall samples (share of samples, share of executed, instruction):\n")
        for r in sorted(rows, key=lambda r: -int(r[i_smp]))[:40]:
            f.write("  %5.2f %%  %5.2f %%  #%d  %s\n" % (100.0 * int(r[i_smp]) / max(1, tot_smp), 100.0 * int(r[i_ex]) / max(1, tot_ex),
                                                        rows.index(r), r[i_src].strip()))
        f.write("\nper region of 40 SASS instructions: share of executed instructions, share of stall samples\n")
        for b in range(0, len(rows), 40):
            seg = rows[b:b + 40]
            f.write("  #%4d..%4d  %5.1f %%  %5.1f %%\n" % (b, b + len(seg) - 1, 100.0 * sum(int(r[i_ex]) for r in seg) / max(1, tot_ex),
                                                           100.0 * sum(int(r[i_smp]) for r in seg) / max(1, tot_smp)))


if __name__ == "__main__":
    main()
