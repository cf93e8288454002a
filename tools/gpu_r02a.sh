#!/bin/bash
# round-2 baseline probe: box facts + configs[2] at full size on the round-1 kernel
OUT=gpurun_out; mkdir -p $OUT
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv; nproc; free -g; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; ulimit -l; } > $OUT/r02a_box.txt 2>&1
timeout 900 python bench.py --codec imbe7200x4400 --streams 1048576 --no-cpu-baseline --no-e2e --steps 3 --warmup 2 > $OUT/r02a_imbe_1m_dev.json 2> $OUT/r02a_imbe_1m_dev.err
echo "dev exit $?"; tail -c 600 $OUT/r02a_imbe_1m_dev.json
timeout 900 python bench.py --codec imbe7200x4400 --streams 1048576 --no-cpu-baseline --steps 3 --warmup 2 > $OUT/r02a_imbe_1m_e2e.json 2> $OUT/r02a_imbe_1m_e2e.err
echo "e2e exit $?"; tail -c 300 $OUT/r02a_imbe_1m_e2e.err; python - <<'PY'
import json
for f in ("r02a_imbe_1m_dev","r02a_imbe_1m_e2e"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.4g ms/step %.1f e2e %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
    except Exception as e: print(f, "no line", e)
PY
cat $OUT/r02a_box.txt
