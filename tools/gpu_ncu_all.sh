#!/bin/bash
# one `ncu --set full` capture per kernel of the multi-kernel path (and, with FUSED=1, of the fused kernel) for a list of workloads
# usage: bash tools/gpu_ncu_all.sh <tag> "<workload> ..."   workload = codec:kind  (kind: hard | softch | tones)
# captures land in gpurun_out/<tag>_<codec>_<kind>_<kernel>_s<streams>x50.ncu-rep; summarise with tools/ncu_constants.py
TAG=$1; WL=$2; NS=${NS:-10656}
OUT=gpurun_out; mkdir -p $OUT
for w in $WL; do
  codec=${w%%:*}; kind=${w##*:}
  case $kind in
    hard) args="--codec $codec";;
    softch) args="--codec $codec --soft-channel";;
    tones) args="--tones-unvoiced";;
  esac
  for spec in parameter:mbe_stream_kernel bank:mbe_split_bank unvoiced:mbe_split_unvoiced; do
    kn=${spec%%:*}; rx=${spec##*:}
    MBE_B200_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f \
        -o $OUT/${TAG}_${codec}_${kind}_${kn}_s${NS}x50 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $args --streams $NS \
        > $OUT/${TAG}_ncu.log 2>&1 || tail -3 $OUT/${TAG}_ncu.log
  done
  if [ -n "$FUSED" ]; then
    MBE_B200_SPLIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mbe_stream_kernel -s 1 -c 1 -f \
        -o $OUT/${TAG}_${codec}_${kind}_fused_s${NS}x50 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $args --streams $NS \
        > $OUT/${TAG}_ncu.log 2>&1 || tail -3 $OUT/${TAG}_ncu.log
  fi
done
# launch list of the default workload shape (durations only; cold cache, serialised)
MBE_B200_SPLIT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mbe_ -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --codec imbe7200x4400 --streams 131072 > $OUT/${TAG}_launch.log 2>&1
ls $OUT/${TAG}_*.ncu-rep | wc -l
