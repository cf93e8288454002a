#!/bin/bash
# e2e of the host-pointer call against the frame-window split of the tail chunks (MBE_B200_TAILSPLIT = windows, 1 = off)
TAG=${1:-t}; shift
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_tail.txt
for g in "$@"; do
  MBE_B200_TAILSPLIT=$g python tools/gpu_e2e_probe.py 2>&1 | tail -1 | sed "s/^/tailsplit $g | /" >> $OUT/${TAG}_tail.txt
done
cat $OUT/${TAG}_tail.txt
