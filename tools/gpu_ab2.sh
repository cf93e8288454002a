#!/bin/bash
# parity tests on the product build, then A/B of builds on the AMBE+2 and IMBE hard-decision arms.
# usage: bash tools/gpu_ab2.sh <tag> lib1.so lib2.so ...   (paths relative to mbelib-neo_b200/)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
: > $OUT/${TAG}_ab.txt
for rep in 1 2; do
for lib in "$@"; do
  for args in "" "--codec imbe7200x4400"; do
    MBE_B200_LIB=$PWD/mbelib-neo_b200/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
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
done
cat $OUT/${TAG}_ab.txt
