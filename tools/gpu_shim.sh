#!/bin/bash
# shim check: the reference's test programs on the GPU path + the shim parity test.  usage: bash tools/gpu_shim.sh <tag>
TAG=${1:-shim}
OUT=gpurun_out; mkdir -p $OUT
for t in test_api test_floattoshort_parity test_golden_pcm test_noise_determinism test_frame_paths test_ecc test_params test_input_validation; do
  echo "== $t" >> $OUT/${TAG}_reftests.log
  timeout 120 oracle/_ref/shim_$t >> $OUT/${TAG}_reftests.log 2>&1
  echo "exit $?" >> $OUT/${TAG}_reftests.log
done
cat $OUT/${TAG}_reftests.log | tail -40
timeout 900 python -m pytest tests/test_gpu_shim.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -30 $OUT/${TAG}_pytest.log
