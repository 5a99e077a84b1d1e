#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box): python scripts/ncu_summary.py rep [regex]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|"
                 r"sm__warps_active.avg.pct|launch__(registers_per_thread|grid_size|block_size|occupancy_limit|shared_mem_per_block_dynamic)|"
                 r"smsp__inst_executed.sum$|smsp__issue_active.avg.pct|sm__cycles_elapsed.avg$|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared|lts__t_sector_hit_rate.pct|"
                 r"smsp__average_warps?_issue_stalled.*_per_issue_active|smsp__average_warp_latency_issue_stalled|sm__throughput.avg.pct|"
                 r"l1tex__throughput.avg.pct|lts__throughput.avg.pct|lts__t_sectors_op_read.sum$|lts__t_sectors.sum$")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:90], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            print("  %-95s %-14s %s" % (h, u, v))
