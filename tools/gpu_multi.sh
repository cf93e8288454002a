#!/bin/bash
# N-GPU box: pool tests over real devices, the C pool demo on 1 and N devices, bench.py on N GPUs.
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N>
TAG=${1:-multi}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_pool.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
: > $OUT/${TAG}_pool_demo.txt
for d in 1 $N; do
  timeout 300 examples/mbe_pool_demo 131072 50 $d >> $OUT/${TAG}_pool_demo.txt 2>&1
done
cat $OUT/${TAG}_pool_demo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench exit $?"; tail -1 $OUT/${TAG}_bench_n$N.json | cut -c1-600
