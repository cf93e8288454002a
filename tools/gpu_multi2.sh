#!/bin/bash
# N-GPU box: pool tests over real devices, the C pool demo on 1 and N devices (pinned + pageable), bench.py (default = configs[2]) on N GPUs
# usage (under gpurun --gpus N): bash tools/gpu_multi2.sh <tag> <N>
TAG=${1:-multi}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt; nvidia-smi topo -m >> $OUT/${TAG}_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_pool.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
: > $OUT/${TAG}_pool_demo.txt
for d in 1 $N; do
  timeout 300 examples/mbe_pool_demo 262144 50 $d >> $OUT/${TAG}_pool_demo.txt 2>&1
done
timeout 300 examples/mbe_pool_demo 262144 50 $N pageable >> $OUT/${TAG}_pool_demo.txt 2>&1
cat $OUT/${TAG}_pool_demo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench exit $?"; tail -2 $OUT/${TAG}_bench_n$N.err
python - $OUT/${TAG}_bench_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d.get("e2e") or {}
print("N=%d value %.4g  e2e %.4g (%s)  link ceiling %.4g frac_of_link %.3f  bytes-input e2e %.4g packed %.4g" % (d["n_gpus"], d["value"], e.get("value",0), e.get("input"),
      (e.get("link_ceiling") or {}).get("value",0), e.get("frac_of_link",0), (e.get("bytes_input") or {}).get("value",0), (e.get("packed_input") or {}).get("value",0)))
PY
