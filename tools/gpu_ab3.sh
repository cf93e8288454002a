#!/bin/bash
# A/B of experimental builds (paths relative to the repo root) on the AMBE+2 and IMBE hard-decision arms, no tests.
# usage: bash tools/gpu_ab3.sh <tag> build/var/a.so build/var/b.so ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_ab.txt
for lib in "$@"; do
  for args in "" "--codec imbe7200x4400"; do
    MBE_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
    python - "$lib" "$args" >> $OUT/${TAG}_ab.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-24s %-24s %.4g frames/s  %.2f ms/step" % (sys.argv[1], sys.argv[2] or "(ambe+2)", d["value"], d["ms_per_step"]))
except Exception as e:
    print("%-24s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-400:])
PY
  done
done
cat $OUT/${TAG}_ab.txt
