#!/bin/bash
# split path profile: launch list (durations) + one full ncu capture per split kernel
# usage: bash tools/gpu_r02i.sh <tag> [codec] [streams_for_full_capture] [kernels regex list]
TAG=$1; CODEC=${2:-imbe7200x4400}; NS=${3:-16576}; KS=${4:-"mbe_split_bank mbe_split_unvoiced mbe_stream_kernel"}
OUT=gpurun_out; mkdir -p $OUT
export MBE_B200_SPLIT=1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:mbe_ -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --codec $CODEC --streams $NS > $OUT/${TAG}_launch.log 2>&1
grep -v "^==" $OUT/${TAG}_launches.csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin))
h=r[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); mi=h.index('Metric Name')
for x in r[1:]:
    print('%-50s %-28s %s' % (x[ki][:50], x[mi], x[vi]))
" | tail -24
for k in $KS; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $OUT/${TAG}_$k \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --codec $CODEC --streams $NS > $OUT/${TAG}_ncu_$k.log 2>&1
tail -1 $OUT/${TAG}_ncu_$k.log
done
