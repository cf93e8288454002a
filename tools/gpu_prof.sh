#!/bin/bash
# ncu full capture of the stream kernel on a reduced batch.  usage: bash tools/gpu_prof.sh <tag> [bench args]
TAG=${1:-p}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mbe_stream_kernel -s 1 -c 1 -f -o $OUT/${TAG}_stream \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --streams 16576 "$@" > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
