#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for name, grid, block, us in rows:
        a = agg[name]
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms total (cold-cache, serialised: compare SHARES)")
    print(f"{'kernel':60s} {'count':>7s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{name[:60]:60s} {cnt:7d} {us / 1e3:10.3f} {us / total:7.1%} {us / cnt:9.1f}")
    return rows


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
