#!/bin/bash
# the round as the driver runs it: full GPU suite, smoke, default bench (+ reference arm)
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log; tail -2 $OUT/${TAG}_smoke.log
( time timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ) 2> $OUT/${TAG}_bench.time
echo "bench exit $?"; tail -3 $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.time
python - $OUT/${TAG}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g e2e %.4g launches %s path %s" % (d["value"], (d.get("e2e") or {}).get("value", 0), d["gpu_launches"], d.get("kernel_path")))
print("roofline", json.dumps({k:v for k,v in d["roofline"].items() if k not in ("peak_source","note","ncu")})[:900])
print("kernels", json.dumps(d.get("kernels")))
print("cpu_baseline", json.dumps(d.get("cpu_baseline"))[:400])
PY
if [ "$2" = "ref" ]; then timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2>$OUT/${TAG}_bench_ref.err; head -c 600 $OUT/${TAG}_bench_ref.json; fi
