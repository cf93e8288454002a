#!/bin/bash
# split path: parity tests, then A/B bench fused vs split (env MBE_B200_SPLIT) for chosen libs
# usage: bash tools/gpu_r02h.sh <tag> "<pytest args>" "<libs>" ["<extra bench arg sets separated by |>"]
TAG=$1; PYT=$2; LIBS=${3:-mbelib-neo_b200/libmbe_b200.so}
OUT=gpurun_out; mkdir -p $OUT
if [ -n "$PYT" ]; then
  timeout 1500 python -m pytest $PYT -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -25 $OUT/${TAG}_pytest.log
fi
: > $OUT/${TAG}_ab.txt
for lib in $LIBS; do
 for split in 0 1; do
  for args in "--codec imbe7200x4400 --streams 131072" "--codec ambe3600x2450 --streams 65536"; do
    MBE_B200_SPLIT=$split MBE_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
    python - "$lib split=$split" "$args" >> $OUT/${TAG}_ab.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-40s %-44s %.4g frames/s  %.2f ms/step launches %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches")))
except Exception as e:
    print("%-40s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-600:])
PY
  done
 done
done
cat $OUT/${TAG}_ab.txt
