#!/usr/bin/env python
"""Warp stall reasons and issue / pipe figures of an `ncu --set full` report:
ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_stalls.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
extra = ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
         "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
         "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
         "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active"]
for r in data:
    for k in extra:
        if k in hdr:
            print(f"{k} [{units[hdr.index(k)]}] = {r[hdr.index(k)]}")
    print("warp stall reasons (average warps stalled per issue-active cycle):")
    pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
    for i, h in enumerate(hdr):
        if h.startswith(pre) and h.endswith(suf) and "not_issued" not in h:
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.05:
                print(f"  {h[len(pre):-len(suf)]:28s} {v:.2f}")
