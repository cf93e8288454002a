#!/bin/bash
# generic env-sweep A/B: each line of the here-doc in $2 is "ENV=.. ENV=.. | bench args"
TAG=$1; SPEC=$2; OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_sweep.txt
while IFS='|' read -r envs args; do
  [ -z "$args" ] && continue
  env $envs python bench.py --no-cpu-baseline --no-e2e --steps 4 --warmup 3 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
  python - "$envs" "$args" >> $OUT/${TAG}_sweep.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-70s %-40s %.4g frames/s  %.2f ms/step launches %s" % (sys.argv[1].strip(), sys.argv[2].strip(), d["value"], d["ms_per_step"], d.get("gpu_launches")))
except Exception as e:
    print("%-70s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-600:])
PY
done < $SPEC
cat $OUT/${TAG}_sweep.txt
