#!/bin/bash
# occupancy sensitivity: same library, 2 blocks/SM vs 1 block/SM (padded dynamic shared memory)
OUT=gpurun_out; mkdir -p $OUT
for pad in 0 20000; do
  MBE_B200_PAD_SMEM=$pad timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > $OUT/occ_tmp.json 2>$OUT/occ_tmp.err
  python - $pad <<'PY'
import json,sys
d=json.loads(open("gpurun_out/occ_tmp.json").read().strip().splitlines()[-1])
print("pad %s: %.4g frames/s %.2f ms" % (sys.argv[1], d["value"], d["ms_per_step"]))
PY
done
