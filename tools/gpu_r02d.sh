#!/bin/bash
# A/B of variant builds + one ncu capture of the IMBE hard kernel for a chosen lib
# usage: bash tools/gpu_r02d.sh <tag> "<ab libs>" <ncu lib> [codec]
TAG=$1; ABLIBS=$2; NLIB=$3; NCODEC=${4:-imbe7200x4400}
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_ab.txt
for lib in $ABLIBS; do
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
if [ -n "$NLIB" ]; then
MBE_B200_LIB=$PWD/$NLIB timeout 900 ncu --set full --clock-control none --import-source on -k regex:mbe_stream_kernel -s 1 -c 1 -f -o $OUT/${TAG}_ncu \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --codec $NCODEC --streams 16576 > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log
fi
