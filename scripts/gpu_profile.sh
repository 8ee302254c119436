#!/bin/bash
# ncu pass (one GPU, tagged): launch list + full captures of the stepwise kernels and the fused kernel.
set -u
mkdir -p gpurun_out
WL=${WL:-cfg4_100k}; ORD=${ORD:-1}; TAG=${TAG:-r1}; TIL=${TIL:--1}
LOG=gpurun_out/profile_$TAG.log
STEP="python bench.py --workload $WL --order $ORD --tiling $TIL --mode stepwise --steps 6 --warmup 3 --skip-e2e"
FUSE="python bench.py --workload $WL --order $ORD --tiling $TIL --mode fused --steps 12 --warmup 3 --skip-e2e"
echo "$(date +%T) plain step" > $LOG
timeout 300 $STEP >> $LOG 2>&1 || { echo "plain stepwise run failed" >> $LOG; tail -5 $LOG; exit 1; }
echo "$(date +%T) ncu launches step" >> $LOG
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_step_$TAG.csv $STEP >> $LOG 2>&1
echo "$(date +%T) rc=$? ncu full step" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_term|k_gram|k_loss|k_dual' -s 40 -c 9 -o gpurun_out/prof_step_$TAG -f $STEP >> $LOG 2>&1
echo "$(date +%T) rc=$? plain fused" >> $LOG
timeout 300 $FUSE >> $LOG 2>&1 || { echo "plain fused run failed" >> $LOG; exit 1; }
echo "$(date +%T) ncu launches fused" >> $LOG
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fused_$TAG.csv $FUSE >> $LOG 2>&1
echo "$(date +%T) rc=$? ncu full fused" >> $LOG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f $FUSE >> $LOG 2>&1
echo "$(date +%T) rc=$? done" >> $LOG
grep -E "^[0-9]{2}:|Profiling|profiling_run" $LOG | tail -30
