#!/usr/bin/env python
"""Per-kernel summary of the LAST training step in an ncu `--metrics gpu__time_duration.sum --csv` launch list of
`bench.py --steps 1 --warmup W` (steps are delimited by their adam_kernel launch)."""
import collections
import csv
import sys


def main(path, thresh_us=60.0):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    names = [r["Kernel Name"].split("(")[0].replace("pnvo::", "").replace("void ", "") for r in rows]
    vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
    idx = [i for i, n in enumerate(names) if "adam" in n]
    # (a capture of exactly one step -- PNVO_PROFILE_STEP=1 with ncu --profile-from-start off -- has a single marker)
    a, b = (idx[-2] + 1 if len(idx) >= 2 else 0), (idx[-1] + 1 if idx else len(names))
    by = collections.defaultdict(lambda: [0, 0.0])
    for n, v in zip(names[a:b], vals[a:b]):
        by[n][0] += 1
        by[n][1] += v
    tot = sum(vals[a:b])
    print("%d launches in the last step, total %.3f ms (serialised, cold-cache ncu timing)" % (b - a, tot / 1e6))
    for k, (n, v) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %4d %8.3f ms %5.1f%%" % (k[:58], n, v / 1e6, 100 * v / tot))
    print("\nlaunches above %.0f us, in program order:" % thresh_us)
    for n, v, r in zip(names[a:b], vals[a:b], rows[a:b]):
        if v > thresh_us * 1e3:
            print("  %-50s grid=%-14s %8.1f us" % (n[:50], r.get("Grid Size", ""), v / 1e3))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 60.0)
