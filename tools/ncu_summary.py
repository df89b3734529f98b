#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports (one row per captured launch) into the few metrics the roofline argument
uses.  Usage: ncu_summary.py raw1.csv [raw2.csv ...] > profiles/rNN_ncu_full_summary.txt"""
import csv
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]

for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("--- %s   [%s]" % (r[idx["Kernel Name"]].split("(")[0], path.split("/")[-1]))
        for k in KEYS:
            if k in idx:
                print("%-78s %s %s" % (k, r[idx[k]], units[idx[k]]))
