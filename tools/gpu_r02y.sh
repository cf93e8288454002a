#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_pool.py tests/test_gpu_shim.py -x -q -s 2>&1 | tail -15
for mode in "" pageable; do examples/mbe_pool_demo 262144 50 0 $mode; done 2>&1 | tee $OUT/r02y_pool_demo.txt
