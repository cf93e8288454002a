#!/bin/bash
# e2e throughput of the host-pointer call against the pipeline shape: compute streams x taper x chunk count
# usage: bash tools/gpu_pipe_sweep.sh <tag> "k:taper:chunks" ...
TAG=${1:-p}; shift
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_pipe.txt
for spec in "$@"; do
  IFS=: read k t c <<< "$spec"
  MBE_B200_KSTREAMS=$k MBE_B200_TAPER=$t MBE_B200_CHUNKS=$c timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 $PIPE_ARGS > $OUT/tmp.json 2> $OUT/tmp.err
  python - $k $t $c >> $OUT/${TAG}_pipe.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1])
    print("kstreams %s taper %3s chunks %3s  device %.4g  e2e %.4g (%.2f ms)  e2e packed %.4g frames/s" % (sys.argv[1], sys.argv[2], sys.argv[3], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], (d["e2e"].get("packed_input") or {"value": float("nan")})["value"]))
except Exception as e:
    print("FAILED", sys.argv[1:], e, open("gpurun_out/tmp.err").read()[-300:])
PY
done
cat $OUT/${TAG}_pipe.txt
