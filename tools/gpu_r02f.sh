#!/bin/bash
# full GPU test-suite on the in-tree build, A/B against variant libs, then the default bench line (configs[2], full size)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
: > $OUT/${TAG}_ab.txt
for lib in "$@"; do
  for args in "--codec imbe7200x4400 --streams 131072" "--codec ambe3600x2450 --streams 65536"; do
    MBE_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
    python - "$lib" "$args" >> $OUT/${TAG}_ab.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-28s %-44s %.4g frames/s  %.2f ms/step" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"]))
except Exception as e:
    print("%-28s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-400:])
PY
  done
done
cat $OUT/${TAG}_ab.txt
( time timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ) 2> $OUT/${TAG}_bench.time
echo "bench exit $?"; tail -5 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.time | tail -3
head -c 3000 $OUT/${TAG}_bench.json
