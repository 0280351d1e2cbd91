#!/usr/bin/env python
"""Summarise the source page of an .ncu-rep: runs of equal execution count (address order) with stall samples,
and optionally the SASS of a line range.  usage: ncu_src.py rep [lo hi]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Source" in r)
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) > idx["Instructions Executed"] and r[idx["Instructions Executed"]].isdigit()]
half = len(data)
E, S = idx["Instructions Executed"], idx["# Samples"]
if len(sys.argv) > 3:
    for i in range(int(sys.argv[2]), int(sys.argv[3])):
        r = data[i]
        print(i, r[E].rjust(9), r[S].rjust(6), r[idx["Source"]][:120])
    sys.exit()
tot_i = sum(int(r[E]) for r in data); tot_s = sum(int(r[S]) for r in data)
print("instructions", tot_i, "samples", tot_s)
runs = []
for i, r in enumerate(data):
    n, s = int(r[E]), int(r[S])
    if runs and runs[-1][0] == n: runs[-1][1] += 1; runs[-1][2] += s
    else: runs.append([n, 1, s, i])
for n, k, s, i in runs:
    if n * k > tot_i / 400: print(f"line {i:5d} exec={n:>9} x{k:<4} samples={s:>6} ({100*s/tot_s:.1f}%)")
