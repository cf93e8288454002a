#!/bin/bash
# compute-sanitizer over every codec / input kind (tools/gpu_sanitize.py) + the full GPU test suite.
# usage: bash tools/gpu_sanitize.sh <tag>
TAG=${1:-san}
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/${TAG}_compute_sanitizer.txt
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool python tools/gpu_sanitize.py (45 streams x 6 frames, all four codecs, hard + soft (channel-like and random reliabilities) + ECC-only + packed + every stage-level entry point)" >> $OUT/${TAG}_compute_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tools/gpu_sanitize.py 2>&1 | grep -v "^=========     \|^$" | tail -25 >> $OUT/${TAG}_compute_sanitizer.txt
done
cat $OUT/${TAG}_compute_sanitizer.txt
