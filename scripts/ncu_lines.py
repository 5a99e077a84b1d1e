#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample breakdown of an .ncu-rep captured with --import-source on
and -lineinfo:  python scripts/ncu_lines.py rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
items = []
for r in rows:
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0] or r[2] != "-":
        continue
    d = dict(zip(range(len(hdr)), r))
    try:
        inst = int(r[hdr.index("Instructions Executed")])
        samp = int(r[hdr.index("Warp Stall Sampling (All Samples)")] or 0)
    except ValueError:
        continue
    items.append((inst, samp, r[0], r[1].strip()[:100]))
ti = sum(i[0] for i in items) or 1
ts = sum(i[1] for i in items) or 1
print("total warp-instructions %d, stall samples %d" % (ti, ts))
print("by instructions:")
for inst, samp, ln, src in sorted(items, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% samp  L%-4s %s" % (100 * inst / ti, 100 * samp / ts, ln, src))
