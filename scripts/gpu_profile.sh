#!/bin/bash
# ncu pass (one GPU): launch list of the stepwise pipeline, full captures of the hot kernels.
set -u
mkdir -p gpurun_out
WL=${WL:-cfg4_100k}; ORD=${ORD:-32}; TAG=${TAG:-r1}
STEP="python bench.py --workload $WL --order $ORD --mode stepwise --steps 12 --warmup 3 --skip-e2e"
FUSE="python bench.py --workload $WL --order $ORD --mode fused --steps 20 --warmup 3 --skip-e2e"
$STEP > gpurun_out/plain_step_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step_$TAG.csv $STEP > gpurun_out/ncu_step_$TAG.log 2>&1
$STEP > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_term|k_gram|k_loss|k_dual|k_exp' -s 60 -c 12 -o gpurun_out/prof_step_$TAG -f $STEP > gpurun_out/ncu_full_step_$TAG.log 2>&1
$FUSE > gpurun_out/plain_fused_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 1 -c 1 -o gpurun_out/prof_fused_$TAG -f $FUSE > gpurun_out/ncu_full_fused_$TAG.log 2>&1
cat gpurun_out/plain_step_$TAG.log gpurun_out/plain_fused_$TAG.log | tail -4
ls -la gpurun_out | tail -12
