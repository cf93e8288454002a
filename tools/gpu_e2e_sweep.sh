#!/bin/bash
# e2e sweep: each line of $2 is "ENV=.. | bench args"; prints device-resident and e2e values
TAG=$1; SPEC=$2; OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_sweep.txt
while IFS='|' read -r envs args; do
  [ -z "$args" ] && continue
  env $envs python bench.py --no-cpu-baseline --steps 3 --warmup 2 $args > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
  python - "$envs" "$args" >> $OUT/${TAG}_sweep.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    e=d.get("e2e") or {}
    print("%-40s %-34s dev %.4g  e2e bytes %.4g packed %.4g  link %.4g" % (sys.argv[1].strip(), sys.argv[2].strip(), d["value"],
          (e.get("bytes_input") or {}).get("value", 0), (e.get("packed_input") or {}).get("value", 0), (e.get("link_ceiling") or {}).get("value", 0)))
except Exception as ex:
    print("%-40s FAILED %s" % (sys.argv[1], ex)); print(open("gpurun_out/ab_tmp.err").read()[-600:])
PY
done < $SPEC
cat $OUT/${TAG}_sweep.txt
