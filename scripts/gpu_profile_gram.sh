#!/bin/bash
# reduced ncu pass after a Gram-kernel change: stepwise launch list, full capture of k_gram and of k_fused
set -u
mkdir -p gpurun_out
TAG=${TAG:-r1c}
LOG=gpurun_out/profile_$TAG.log
STEP="python bench.py --workload cfg4_100k --mode stepwise --steps 6 --warmup 3 --skip-e2e"
FUSE="python bench.py --workload cfg4_100k --mode fused --steps 12 --warmup 3 --skip-e2e"
timeout 300 $STEP > $LOG 2>&1 || { echo "plain stepwise run failed"; tail -5 $LOG; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_step_$TAG.csv $STEP >> $LOG 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gram' -s 4 -c 2 -o gpurun_out/prof_gram_$TAG -f $STEP >> $LOG 2>&1
timeout 300 $FUSE >> $LOG 2>&1 || { echo "plain fused run failed"; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f $FUSE >> $LOG 2>&1
grep -E "Profiling|profiling_run" $LOG | tail -8 | cut -c1-200
