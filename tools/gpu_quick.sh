#!/bin/bash
# quick GPU check: parity tests + one bench line (no ncu).  usage: bash tools/gpu_quick.sh <tag> [bench args]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --no-cpu-baseline "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.4g frames/s  ms/step %.2f  e2e %.4g  fp32 frac %.3f" % (d["value"], d["ms_per_step"], (d["e2e"] or {}).get("value",0), d["roofline_fp32"]["frac"]))
except Exception as e: print("no bench line", e)
PY
