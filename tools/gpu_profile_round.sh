#!/bin/bash
# bench line + reference arm + ncu launch list + one full capture of the stream kernel (no tests).
# usage: bash tools/gpu_profile_round.sh <tag>
TAG=${1:-r}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; cat $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mbe_stream_kernel -s 1 -c 1 -f -o $OUT/${TAG}_stream \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --streams 16576 > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log
ls -la $OUT | tail -12
