#!/usr/bin/env python
"""Top source lines per stall reason from an `ncu --page source --csv --print-source cuda,sass` export.
usage: python tools/ncu_stall_lines.py src2.csv stall_long_sb[,stall_no_inst,...] [top]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
stalls = sys.argv[2].split(",")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
col = None; cur = None
agg = defaultdict(lambda: defaultdict(float)); src = {}; total = 0.0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': col = {n: i for i, n in enumerate(r)}; continue
    if col is None or len(r) < len(col) or r[col['Address']] != '-': continue
    k = (cur, r[0]); src[k] = r[1]
    try: total += float(r[col['# Samples']] or 0)
    except ValueError: pass
    for st in stalls:
        try: agg[st][k] += float(r[col[st]] or 0)
        except (ValueError, KeyError): pass
for st in stalls:
    t = sum(agg[st].values())
    print("== %s: %.1f%% of all samples" % (st, 100 * t / total))
    for k, v in sorted(agg[st].items(), key=lambda x: -x[1])[:top]:
        print("  %5.1f%%  %s:%s  %s" % (100 * v / t, k[0], k[1], src[k][:100]))
