#!/bin/bash
# tests on the in-tree build (both paths), then split-path A/B over variant libs
TAG=$1; LIBS=$2; PYT=${3:-"tests/test_gpu_split.py tests/test_gpu_parity.py -m gpu"}
OUT=gpurun_out; mkdir -p $OUT
if [ "$PYT" != "none" ]; then
timeout 1500 python -m pytest $PYT -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
fi
: > $OUT/${TAG}_ab.txt
for lib in $LIBS; do
  for args in "--codec imbe7200x4400 --streams 131072" "--codec ambe3600x2450 --streams 65536"; do
    MBE_B200_SPLIT=1 MBE_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 4 --warmup 3 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
    python - "$lib" "$args" >> $OUT/${TAG}_ab.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-36s %-44s %.4g frames/s  %.2f ms/step launches %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches")))
except Exception as e:
    print("%-36s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-600:])
PY
  done
done
cat $OUT/${TAG}_ab.txt
