#!/bin/bash
# throughput against frames per launch (a real-time server launches ONE 20 ms frame per stream at a time)
TAG=${1:-sweep}
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_frames_sweep.txt
for lib in ${LIBS:-libmbe_b200.so}; do
echo "== $lib" >> $OUT/${TAG}_frames_sweep.txt
export MBE_B200_LIB=$PWD/mbelib-neo_b200/$lib
for spec in "1 1048576" "2 524288" "5 262144" "10 131072" "50 65536"; do
  set -- $spec
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps $((50 / $1 > 5 ? 50 / $1 : 5)) --warmup 3 --frames $1 --streams $2 $EXTRA > $OUT/tmp.json 2> $OUT/tmp.err
  python - "$1" "$2" >> $OUT/${TAG}_frames_sweep.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1])
    print("frames/launch %3s  streams %8s  %.4g frames/s  %.2f ms/launch  hbm %.0f GB/s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["achieved"]))
except Exception as e:
    print("FAILED", sys.argv[1:], e); print(open("gpurun_out/tmp.err").read()[-500:])
PY
done
done
cat $OUT/${TAG}_frames_sweep.txt
