#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: executed instructions and stall samples per stall reason,
total and grouped by address ranges of N instructions (default 512) so hot regions stand out.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_sass_summary.py src.csv [group]"""
import csv
import sys

path = sys.argv[1]
group = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]


def f(r, n):
    try:
        return float(r[col[n]])
    except ValueError:
        return 0.0


tot_exec = sum(f(r, "Instructions Executed") for r in data)
tot_samp = sum(f(r, "# Samples") for r in data)
executed = sum(1 for r in data if f(r, "Instructions Executed") > 0)
print("instructions in kernel: %d (%.0f KB); ever executed: %d; executed warp-instr: %.4g; samples: %.0f" % (
    len(data), len(data) * 16 / 1024, executed, tot_exec, tot_samp))
print("stall share of samples:")
for n in sorted(stalls, key=lambda n: -sum(f(r, n) for r in data)):
    s = sum(f(r, n) for r in data)
    if s > 0:
        print("  %-24s %6.2f%%" % (n, 100 * s / tot_samp))
print("regions of %d instructions: [index] exec%% samples%% top-stalls  first-source-op" % group)
for g in range(0, len(data), group):
    blk = data[g:g + group]
    e = sum(f(r, "Instructions Executed") for r in blk)
    s = sum(f(r, "# Samples") for r in blk)
    if e / tot_exec < 0.005 and s / tot_samp < 0.005:
        continue
    top = sorted(stalls, key=lambda n: -sum(f(r, n) for r in blk))[:3]
    tops = " ".join("%s=%.0f%%" % (n[6:], 100 * sum(f(r, n) for r in blk) / max(s, 1)) for n in top)
    print("  [%6d] %5.1f%% %5.1f%%  %s" % (g, 100 * e / tot_exec, 100 * s / tot_samp, tops))
