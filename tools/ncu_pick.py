#!/usr/bin/env python
"""Reads `ncu -i X.ncu-rep --page raw --csv` on stdin and prints, per profiled launch, the handful of metrics that say what
limits a kernel (duration, DRAM / L2 / shared throughput, achieved occupancy, registers, issue-slot utilisation, top stall)."""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]  # fmt: skip


def main():
    rows = list(csv.reader(ln for ln in sys.stdin if not ln.startswith("==")))
    if len(rows) < 3:
        print("no data")
        return
    header, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(header)}
    for r in rows[2:]:
        name = r[idx.get("Kernel Name", 4)][:70]
        print(f"--- {name}  grid {r[idx.get('Grid Size', 0)]} block {r[idx.get('Block Size', 0)]}")
        for m in WANT:
            if m in idx and idx[m] < len(r):
                print(f"    {m:75s} {r[idx[m]]:>14s} {units[idx[m]]}")


if __name__ == "__main__":
    main()
