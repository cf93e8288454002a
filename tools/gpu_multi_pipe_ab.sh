#!/bin/bash
# N-GPU e2e of the host-pointer call: old pipeline shape (2 compute streams, 16 equal chunks) against the default
# usage (under gpurun --gpus N): bash tools/gpu_multi_pipe_ab.sh <tag> <N>
TAG=${1:-ab}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_pipe_n$N.txt
run() {
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > $OUT/tmp.json 2> $OUT/tmp.err
  python - "$*" >> $OUT/${TAG}_pipe_n$N.txt <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/tmp.json").read().strip().splitlines()[-1])
    print("%-60s device %.4g  e2e %.4g (%.2f ms)" % (sys.argv[1], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("FAILED", sys.argv[1], e, open("gpurun_out/tmp.err").read()[-300:])
PY
}
run MBE_B200_KSTREAMS=2 MBE_B200_TAPER=0 MBE_B200_CHUNKS=16
run MBE_B200_KSTREAMS=4 MBE_B200_TAPER=18 MBE_B200_CHUNKS=32
run MBE_B200_KSTREAMS=2 MBE_B200_TAPER=0 MBE_B200_CHUNKS=16
run MBE_B200_KSTREAMS=4 MBE_B200_TAPER=18 MBE_B200_CHUNKS=32
run MBE_B200_KSTREAMS=4 MBE_B200_TAPER=0 MBE_B200_CHUNKS=16
run MBE_B200_KSTREAMS=2 MBE_B200_TAPER=0 MBE_B200_CHUNKS=32
cat $OUT/${TAG}_pipe_n$N.txt
