#!/bin/bash
# N-GPU box: bench.py --config C for each C in $3.. under torchrun
TAG=$1; N=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
for C in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --config $C --steps 5 --warmup 3 > $OUT/${TAG}_config${C}_n$N.json 2> $OUT/${TAG}_config${C}_n$N.err
  python - $OUT/${TAG}_config${C}_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d.get("e2e") or {}
print("N=%d %s: value %.4g  e2e %.4g  link ceiling %.4g frac_of_link %.3f" % (d["n_gpus"], d["config"]["baseline_config"], d["value"], e.get("value",0),
      (e.get("link_ceiling") or {}).get("value",0), e.get("frac_of_link",0)))
PY
done
