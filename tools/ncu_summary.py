#!/usr/bin/env python
"""Summarises an `ncu --set full` report (.ncu-rep) into the text committed under profiles/: per launch the duration,
DRAM bytes (read + write) and achieved DRAM GB/s, tensor-pipe / issue / occupancy percentages and registers.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_ncu_full_x.txt
    python tools/ncu_summary.py gpurun_out/x.ncu-rep --json conv_stem2_fwd > stem.json   # summed DRAM bytes of the matches
"""
import csv
import json
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct_active"),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma_pct"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma_pipe_pct"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit.lower(), 1e-3)


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = []
    for vals in rows[2:]:
        r = {"name": vals[hdr.index("Kernel Name")], "id": vals[hdr.index("ID")]}
        for key, short in KEYS:
            if key in hdr:
                i = hdr.index(key)
                r[short] = (vals[i], units[i])
        recs.append(r)
    return recs


def main():
    rep = sys.argv[1]
    recs = load(rep)
    if "--json" in sys.argv:
        pat = sys.argv[sys.argv.index("--json") + 1]
        sel = [r for r in recs if pat in r["name"]]
        print(json.dumps({"kernel": pat, "launches": len(sel),
                          "dram_bytes": int(sum(to_bytes(*r["dram_rd"]) + to_bytes(*r["dram_wr"]) for r in sel)),
                          "time_us": round(sum(to_us(*r["time"]) for r in sel), 1), "report": rep}))
        return
    print("# %s: %d launches (ncu --set full --clock-control none; per-launch values are cold-cache and serialised)" % (rep, len(recs)))
    print("%-4s %-58s %9s %10s %10s %9s %7s %7s %7s %7s %6s %5s" % ("id", "kernel", "time us", "dram rd MB", "dram wr MB",
                                                                   "DRAM GB/s", "dram %", "tensor%", "issue %", "occup %", "L2 %", "regs"))
    for r in recs:
        t = to_us(*r["time"])
        rd, wr = to_bytes(*r["dram_rd"]), to_bytes(*r["dram_wr"])
        g = lambda k: r[k][0] if k in r else "-"  # noqa: E731
        name = r["name"].split("(")[0].replace("pnvo::", "").replace("void ", "")
        print("%-4s %-58s %9.1f %10.1f %10.1f %9.0f %7s %7s %7s %7s %6s %5s" % (
            r["id"], name[:58], t, rd / 1e6, wr / 1e6, (rd + wr) / (t * 1e-6) / 1e9, g("dram_pct"),
            g("tensor_pct_active") if g("tensor_pct_active") != "-" else g("hmma_pipe_pct"), g("issue_pct"),
            g("occupancy_pct"), g("l2_pct"), g("regs")))


if __name__ == "__main__":
    main()
