#!/bin/bash
# final evidence of the round: full GPU suite on both kernel paths, smoke, default bench + reference arm, compute-sanitizer
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
bash tools/gpu_round2.sh $TAG ref
MBE_B200_SPLIT=0 timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_fused_default.log 2>&1; echo "pytest (fused default) exit $?" >> $OUT/${TAG}_pytest_fused_default.log; tail -3 $OUT/${TAG}_pytest_fused_default.log
bash tools/gpu_sanitize.sh $TAG > /dev/null 2>&1; grep -E "^==|SUMMARY|hazard|ERROR" $OUT/${TAG}_compute_sanitizer.txt | head -20
