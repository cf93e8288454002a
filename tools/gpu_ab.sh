#!/bin/bash
# A/B experimental builds: bash tools/gpu_ab.sh <tag> lib1.so lib2.so ... (paths relative to repo root)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_ab.txt
for lib in "$@"; do
  MBE_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 $AB_ARGS > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
  python - "$lib" >> $OUT/${TAG}_ab.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-28s %.4g frames/s  %.2f ms/step" % (sys.argv[1], d["value"], d["ms_per_step"]))
except Exception as e:
    print("%-28s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-400:])
PY
done
cat $OUT/${TAG}_ab.txt
