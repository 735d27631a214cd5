#!/usr/bin/env python3
"""Per-kernel table from an .ncu-rep (ncu --set full): launches, total time, and for the longest launch of every kernel
its duration, DRAM bytes and rate, pipe utilisation and issue activity.  usage: ncu_table.py file.ncu-rep > table.md"""
import csv
import io
import re
import subprocess
import sys


def num(d, k):
    try:
        return float(d.get(k, "").replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = {}
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").strip()
        t = num(d, "gpu__time_duration.sum") * scale.get(unit["gpu__time_duration.sum"], 1.0)
        a = agg.setdefault(name, {"n": 0, "tot": 0.0, "best": None, "t": -1})
        a["n"] += 1
        a["tot"] += t
        if t > a["t"]:
            a["t"], a["best"] = t, d
    print("| kernel | launches | total us | longest us | grid | regs | DRAM MB (r+w) | DRAM GB/s | fma pipe % | alu pipe % | issue active % | top stall |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    stalls = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["tot"]):
        d = a["best"]
        by = (num(d, "dram__bytes_read.sum") * scale.get(unit["dram__bytes_read.sum"], 1.0) +
              num(d, "dram__bytes_write.sum") * scale.get(unit["dram__bytes_write.sum"], 1.0))
        top = max(stalls, key=lambda k: num(d, k) if num(d, k) == num(d, k) else -1)
        top = top.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
        print("| %s | %d | %.1f | %.1f | %s | %s | %.1f | %.0f | %.1f | %.1f | %.1f | %s |" % (
            name[:70], a["n"], a["tot"], a["t"], d.get("Grid Size", ""), d.get("launch__registers_per_thread", ""), by / 1e6,
            by / 1e9 / (a["t"] * 1e-6) if a["t"] > 0 else 0,
            num(d, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            num(d, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"), top))


if __name__ == "__main__":
    main()
