#!/bin/bash
# quick parity subset + A/B + stage timing of variant builds.  usage: bash tools/gpu_r02c.sh <tag> "<ab libs>" "<timing libs>"
TAG=$1; ABLIBS=$2; TLIBS=$3
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_ab.txt
for lib in $ABLIBS; do
  MBE_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/${TAG}_pytest_$(basename $lib).log 2>&1
  echo "$lib parity: $(tail -1 $OUT/${TAG}_pytest_$(basename $lib).log)" >> $OUT/${TAG}_ab.txt
  for args in "--codec imbe7200x4400 --streams 131072" "--codec ambe3600x2450 --streams 65536"; do
    MBE_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
    python - "$lib" "$args" >> $OUT/${TAG}_ab.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-24s %-44s %.4g frames/s  %.2f ms/step" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"]))
except Exception as e:
    print("%-24s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-400:])
PY
  done
done
cat $OUT/${TAG}_ab.txt
: > $OUT/${TAG}_stages.txt
for lib in $TLIBS; do
  for codec in 0 3; do
    echo "== $lib codec $codec" >> $OUT/${TAG}_stages.txt
    MBE_B200_LIB=$PWD/$lib timeout 300 python tools/gpu_stage_timing.py $codec 65536 >> $OUT/${TAG}_stages.txt 2>&1
  done
done
cat $OUT/${TAG}_stages.txt
