#!/usr/bin/env python
"""Per-function / per-line breakdown of an `ncu --page source --csv --print-source cuda,sass` export.
usage: python tools/ncu_line_summary.py src2.csv [n_frames] [top_lines]
Instruction counts are inclusive of inlined callees' own lines (each source line is attributed once)."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
n_frames = float(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cur_file, col = None, None
per_line = defaultdict(lambda: [0.0, 0.0, ""])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Line No":
        col = {n: i for i, n in enumerate(r)}
        continue
    if col is None or len(r) < len(col) or r[col["Address"]] != "-":
        continue
    try:
        ln = int(r[0])
        ex = float(r[col["Instructions Executed"]] or 0)
        sm = float(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    e = per_line[(cur_file, ln)]
    e[0] += ex
    e[1] += sm
    e[2] = r[1]

fn_re = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|MBE_HD|__host__)[^(;]*?\b(\w+)\s*\(")
func_of = {}
for f in set(k[0] for k in per_line):
    try:
        lines = open(f).read().split("\n")
    except OSError:
        continue
    cur = "?"
    pend = ""
    for i, text in enumerate(lines, 1):
        t = (pend + " " + text).strip() if pend else text
        m = fn_re.match(t)
        if m:
            cur = m.group(1)
            pend = ""
        elif text.startswith("template"):
            pend = text
        else:
            pend = ""
        func_of[(f, i)] = cur
tot_e = sum(v[0] for v in per_line.values())
tot_s = sum(v[1] for v in per_line.values())
per_fn = defaultdict(lambda: [0.0, 0.0])
for k, v in per_line.items():
    fn = func_of.get(k, "?")
    per_fn[fn][0] += v[0]
    per_fn[fn][1] += v[1]
print("total executed warp-instr %.4g, samples %.0f%s" % (tot_e, tot_s, (", %.0f instr/frame" % (tot_e / n_frames)) if n_frames else ""))
print("%-28s %8s %8s %10s" % ("function", "exec%", "samp%", "instr/frame"))
for fn, v in sorted(per_fn.items(), key=lambda kv: -kv[1][1]):
    if v[0] / tot_e < 0.002 and v[1] / tot_s < 0.002:
        continue
    print("%-28s %7.1f%% %7.1f%% %10.0f" % (fn, 100 * v[0] / tot_e, 100 * v[1] / tot_s, v[0] / n_frames if n_frames else 0))
print("top lines by samples:")
for k, v in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
    print("  %5.1f%% samp %5.1f%% exec  %s:%d  %s" % (100 * v[1] / tot_s, 100 * v[0] / tot_e, k[0].split("/")[-1], k[1], v[2].strip()[:90]))
