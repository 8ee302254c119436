#!/bin/bash
# ncu pass (one GPU): full captures of the hot kernels in stepwise mode (+ optionally the fused kernel).
set -u
mkdir -p gpurun_out
WL=${WL:-cfg4_100k}; ORD=${ORD:-1}; TAG=${TAG:-r1}; TIL=${TIL:--1}; FUSED=${FUSED:-0}
LOG=gpurun_out/profile_$TAG.log
STEP="python bench.py --workload $WL --order $ORD --tiling $TIL --mode stepwise --steps 4 --warmup 3 --skip-e2e"
FUSE="python bench.py --workload $WL --order $ORD --tiling $TIL --mode fused --steps 12 --warmup 3 --skip-e2e"
echo "$(date +%T) plain step" > $LOG
timeout 300 $STEP >> $LOG 2>&1 || { echo "plain stepwise run failed" >> $LOG; tail -5 $LOG; exit 1; }
echo "$(date +%T) ncu full step" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_term -s 8 -c 2 -o gpurun_out/prof_step_$TAG -f $STEP >> $LOG 2>&1
echo "$(date +%T) rc=$?" >> $LOG
if [ "$FUSED" = "1" ]; then
  timeout 300 $FUSE >> $LOG 2>&1 || { echo "plain fused run failed" >> $LOG; exit 1; }
  echo "$(date +%T) ncu launches fused" >> $LOG
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fused_$TAG.csv $FUSE >> $LOG 2>&1
  echo "$(date +%T) rc=$? ncu full fused" >> $LOG
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f $FUSE >> $LOG 2>&1
  echo "$(date +%T) rc=$? done" >> $LOG
fi
grep -E "^[0-9]{2}:|Profiling|profiling_run" $LOG | tail -20
