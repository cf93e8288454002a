#!/bin/bash
# sweep of the multi-kernel path's stream-range size and internal stream count
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_sweep.txt
run() {  # label, env..., -- args
  python bench.py --no-cpu-baseline --no-e2e --steps 4 --warmup 3 $ARGS > $OUT/ab_tmp.json 2>$OUT/ab_tmp.err
  python - "$1" "$ARGS" >> $OUT/${TAG}_sweep.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/ab_tmp.json").read().strip().splitlines()[-1])
    print("%-34s %-44s %.4g frames/s  %.2f ms/step launches %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d.get("gpu_launches")))
except Exception as e:
    print("%-34s FAILED %s" % (sys.argv[1], e)); print(open("gpurun_out/ab_tmp.err").read()[-600:])
PY
}
for ARGS in "--codec imbe7200x4400 --streams 131072" "--codec ambe3600x2450 --streams 65536"; do
  MBE_B200_SPLIT=0 run "fused"
  for aux in 1 2 3 4; do for mb in 256 512 1024 2048; do
    MBE_B200_SPLIT=1 MBE_B200_AUX=$aux MBE_B200_DESC_MB=$mb run "split aux=$aux desc_mb=$mb"
  done; done
done
cat $OUT/${TAG}_sweep.txt
