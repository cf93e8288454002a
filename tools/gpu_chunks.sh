#!/bin/bash
# e2e throughput against the number of pipeline chunks of the host-pointer call
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${1:-c}_chunks.txt
for c in 8 16 24 32 48; do
  MBE_B200_CHUNKS=$c timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > $OUT/tmp.json 2> $OUT/tmp.err
  python - $c >> $OUT/${1:-c}_chunks.txt <<'PY'
import json,sys
d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1])
print("chunks %3s  device %.4g  e2e %.4g  e2e packed %.4g frames/s" % (sys.argv[1], d["value"], d["e2e"]["value"], d["e2e"]["packed_input"]["value"]))
PY
done
cat $OUT/${1:-c}_chunks.txt
