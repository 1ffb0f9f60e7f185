#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name the count and the mean / min / max
duration in microseconds (launches in order of first appearance).   python tools/ncu_kernel_times.py launches.csv [skip]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, errors="replace") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = r["Kernel Name"].split("(")[0]
        rows.append((name, us))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, us in rows:
        agg.setdefault(name, []).append(us)
    total = sum(us for _, us in rows)
    print(f"{len(rows)} launches, {total:.1f} us in total")
    for name, v in agg.items():
        print(f"{name[:90]:90s} n={len(v):4d} mean {sum(v) / len(v):8.1f} us  min {min(v):8.1f}  max {max(v):8.1f}  share {100 * sum(v) / total:5.1f} %")


if __name__ == "__main__":
    main()
