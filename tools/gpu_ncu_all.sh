#!/bin/bash
# one `ncu --set full` capture per kernel of the multi-kernel path (and, with FUSED=1, of the fused kernel) for a list of workloads
# usage: bash tools/gpu_ncu_all.sh <tag> "<workload> ..."   workload = codec:kind  (kind: hard | softch | tones)
# every capture is reduced on the box to gpurun_out/<tag>_<codec>_<kind>_<kernel>_s<streams>x50.raw.csv (ncu --page raw --csv:
# a report with imported sources is ~17 MB, gpurun brings back 64 MiB per call); KEEP="codec:kind:kernel ..." keeps those reports too.
# summarise with tools/ncu_constants.py <tag> gpurun_out/<tag>_*.raw.csv
TAG=$1; WL=$2; NS=${NS:-10656}
OUT=gpurun_out; mkdir -p $OUT
reduce() {  # <report base> <workload:kernel>: raw page as csv; the report itself stays only for the tokens listed in $KEEP
  [ -f $1.ncu-rep ] || return
  ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null
  case " $KEEP " in *" $2 "*) ;; *) rm -f $1.ncu-rep;; esac
}
for w in $WL; do
  codec=${w%%:*}; kind=${w##*:}
  case $kind in
    hard) args="--codec $codec";;
    softch) args="--codec $codec --soft-channel";;
    tones) args="--tones-unvoiced";;
  esac
  for spec in ${KERNELS:-parameter:mbe_stream_kernel bank:mbe_split_bank unvoiced:mbe_split_unvoiced}; do
    kn=${spec%%:*}; rx=${spec##*:}
    MBE_B200_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f \
        -o $OUT/${TAG}_${codec}_${kind}_${kn}_s${NS}x50 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $args --streams $NS \
        > $OUT/${TAG}_ncu.log 2>&1 || tail -3 $OUT/${TAG}_ncu.log
    reduce $OUT/${TAG}_${codec}_${kind}_${kn}_s${NS}x50 $w:$kn
  done
  if [ -n "$FUSED" ]; then
    MBE_B200_SPLIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mbe_stream_kernel -s 1 -c 1 -f \
        -o $OUT/${TAG}_${codec}_${kind}_fused_s${NS}x50 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $args --streams $NS \
        > $OUT/${TAG}_ncu.log 2>&1 || tail -3 $OUT/${TAG}_ncu.log
    reduce $OUT/${TAG}_${codec}_${kind}_fused_s${NS}x50 $w:fused
  fi
done
# launch list of the default workload shape (durations only; cold cache, serialised)
[ -n "$NOLIST" ] || MBE_B200_SPLIT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mbe_ -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --codec imbe7200x4400 --streams 131072 > $OUT/${TAG}_launch.log 2>&1
ls $OUT/${TAG}_*.ncu-rep | wc -l
