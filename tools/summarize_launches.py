#!/usr/bin/env python
"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the top launches."""
import csv
import collections
import sys


def main(path, top=25):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        if unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        rows.append((int(r["ID"]), r["Kernel Name"].split("(")[0], r.get("Grid Size", ""), v))
    tot = sum(r[3] for r in rows)
    by = collections.defaultdict(lambda: [0, 0.0])
    for _, k, _, v in rows:
        by[k][0] += 1
        by[k][1] += v
    print(f"{len(rows)} launches, total {tot/1e6:.3f} ms")
    print("%-40s %6s %10s %7s" % ("kernel", "count", "ms", "share"))
    for k, (n, v) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print("%-40s %6d %10.3f %6.1f%%" % (k[:40], n, v / 1e6, 100 * v / tot))
    print("\ntop launches:")
    for i, k, g, v in sorted(rows, key=lambda r: -r[3])[:top]:
        print("  #%-4d %-32s grid=%-18s %9.3f ms %5.1f%%" % (i, k[:32], g, v / 1e6, 100 * v / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
