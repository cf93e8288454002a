#!/bin/bash
# soft-decision check: parity tests, then the soft bench arms.  usage: bash tools/gpu_soft.sh <tag>
TAG=${1:-soft}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
: > $OUT/${TAG}_bench.txt
for lib in libmbe_b200_old.so libmbe_b200.so; do export MBE_B200_LIB=$PWD/mbelib-neo_b200/$lib; echo "== $lib" >> $OUT/${TAG}_bench.txt
for args in "" "--soft" "--soft-channel" "--codec imbe7200x4400 --soft" "--codec imbe7200x4400 --soft-channel" "--codec imbe7100x4400 --soft-channel" "--codec imbe7200x4400"; do
  if [ "$lib" = libmbe_b200_old.so ] && [[ "$args" != *soft* ]]; then continue; fi
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 3 $args > $OUT/tmp.json 2> $OUT/tmp.err
  python - "$args" >> $OUT/${TAG}_bench.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1])
    print("%-45s %.4g frames/s  %.2f ms/step" % (sys.argv[1] or "(default)", d["value"], d["ms_per_step"]))
except Exception as e:
    print("%-45s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/tmp.err").read()[-600:])
PY
done; done
cat $OUT/${TAG}_bench.txt
