#!/bin/bash
# parity tests on the working tree's build, then A/B of variant builds on IMBE (131072 streams) and AMBE+2 (65536)
# usage: bash tools/gpu_r02b.sh <tag> [lib ...]   (libs relative to the repo root; the in-tree build is always measured first)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -6 $OUT/${TAG}_pytest.log
: > $OUT/${TAG}_ab.txt
for lib in mbelib-neo_b200/libmbe_b200.so "$@"; do
  for args in "--codec imbe7200x4400 --streams 131072" "--codec ambe3600x2450 --streams 65536"; do
    MBE_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
    python - "$lib" "$args" >> $OUT/${TAG}_ab.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-36s %-44s %.4g frames/s  %.2f ms/step" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"]))
except Exception as e:
    print("%-36s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-400:])
PY
  done
done
cat $OUT/${TAG}_ab.txt
