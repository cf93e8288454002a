#!/usr/bin/env python
"""Static SASS size per source function of one kernel, from `nvdisasm -g -c <cubin>` (line info needs -lineinfo builds).
usage: cuobjdump -xelf all lib.so; nvdisasm -g -c mbe_b200.sm_100a.cubin > k.dis; python tools/sass_code_size.py k.dis <kernel substring> [repo root]
Each instruction is attributed to the innermost source line nvdisasm reports for it (inlined callees count for themselves)."""
import re
import sys
from collections import defaultdict

dis, pat = sys.argv[1], sys.argv[2]
root = sys.argv[3] if len(sys.argv) > 3 else None
fn_re = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|MBE_HD|__host__)[^(;]*?\b(\w+)\s*\(")
func_cache = {}


def funcs_of(path):
    if path in func_cache:
        return func_cache[path]
    m = {}
    try:
        p = path
        if root and path.startswith("/root/repo/"):
            p = root + path[len("/root/repo"):]
        lines = open(p).read().split("\n")
    except OSError:
        func_cache[path] = m
        return m
    cur, pend = "?", ""
    for i, text in enumerate(lines, 1):
        t = (pend + " " + text).strip() if pend else text
        mm = fn_re.match(t)
        if mm:
            cur, pend = mm.group(1), ""
        elif text.startswith("template"):
            pend = text
        else:
            pend = ""
        m[i] = cur
    func_cache[path] = m
    return m


inside = False
cur = ("?", 0)
per = defaultdict(int)
total = 0
for line in open(dis):
    if line.startswith("//---") and ".text." in line:
        inside = pat in line
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        fn = funcs_of(cur[0]).get(cur[1], "?")
        per[fn] += 1
        total += 1
print("kernel %s: %d instructions, %.1f KB" % (pat, total, total * 16 / 1024))
for fn, n in sorted(per.items(), key=lambda kv: -kv[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    print("%-28s %6d  %5.1f KB" % (fn, n, n * 16 / 1024))
