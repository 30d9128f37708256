#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    dram = defaultdict(lambda: [0.0, 0.0])  # kernel -> [bytes read, bytes written] when the list carries dram__bytes_*.sum
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    bscale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "B": 1.0, "KB": 1e3, "MB": 1e6, "GB": 1e9}
    for r in rd:
        metric = r.get("Metric Name")
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        if metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            dram[name][0 if metric.endswith("read.sum") else 1] += v * bscale.get(r.get("Metric Unit", "byte"), 1.0)
            continue
        if metric != "gpu__time_duration.sum":
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, grid, block, us in rows:
        a = agg[name]
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms total (cold-cache, serialised: compare SHARES)")
    if dram:
        rd_b, wr_b = sum(v[0] for v in dram.values()), sum(v[1] for v in dram.values())
        print(f"DRAM traffic of the listed launches: {rd_b / 1e9:.2f} GB read + {wr_b / 1e9:.2f} GB written = {(rd_b + wr_b) / 1e9:.2f} GB")
    print(f"{'kernel':60s} {'count':>7s} {'total ms':>10s} {'share':>7s} {'avg us':>9s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'GB/s':>8s}")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        rb, wb = dram.get(name, (0.0, 0.0))
        gbs = (rb + wb) / us / 1e3 if us > 0 else 0.0
        print(f"{name[:60]:60s} {cnt:7d} {us / 1e3:10.3f} {us / total:7.1%} {us / cnt:9.1f} {rb / 1e6:11.1f} {wb / 1e6:11.1f} {gbs:8.0f}")
    return rows


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
