#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report: ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "lts__t_bytes.sum"]
for i, h in enumerate(hdr):
    if h in keys or any(h == k + s for k in keys for s in (".per_second", ".pct_of_peak_sustained_elapsed")):
        print(f"{h} [{units[i]}] = {[r[i] for r in data]}")
