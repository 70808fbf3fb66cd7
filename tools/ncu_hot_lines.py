#!/usr/bin/env python3
"""Aggregate an ncu source page (cuda,sass view) into the hottest CUDA source lines.

usage: ncu_hot_lines.py <report.ncu-rep> [top_n]
"""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
cur_file = "?"
items = []
tot_i = tot_s = 0
hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or not r[0].isdigit():
        continue
    try:
        ns = int(float(r[hdr["# Samples"]]))
        ni = int(float(r[hdr["Instructions Executed"]]))
        nt = int(float(r[hdr["Thread Instructions Executed"]]))
    except (ValueError, KeyError):
        continue
    tot_i += ni
    tot_s += ns
    items.append((ns, ni, nt, cur_file, int(r[0]), r[1].strip()))
items.sort(reverse=True)
print(f"total warp instructions {tot_i:,}   stall samples {tot_s:,}")
for ns, ni, nt, f, line, src in items[:top]:
    eff = nt / ni if ni else 0
    print(f"{100.0 * ns / max(tot_s, 1):5.1f}% smp {100.0 * ni / max(tot_i, 1):5.1f}% inst eff{eff:5.1f} {f}:{line:<4} {src[:100]}")
