#!/bin/bash
# ncu over the END-TO-END call (plan build on the GPU, iterations, filtered final factor): launch list of the whole
# bench command, then full captures of the final factor's kernels.  One GPU.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r2c}
CMD="python bench.py --steps 20 --warmup 5 --no-cpu --tte-iters 0"
timeout 300 $CMD > gpurun_out/pe_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/pe_plain_$TAG.log; exit 1; }
tail -1 gpurun_out/pe_plain_$TAG.log | head -c 300; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_e2e_$TAG.csv $CMD > gpurun_out/pe_ncu1_$TAG.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/launches_e2e_$TAG.csv)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_symv_axpy|k_lz_dot|k_lz_sub|k_lz_reduce|k_lz_finish' -s 4000 -c 16 -o gpurun_out/prof_eig_$TAG -f $CMD > gpurun_out/pe_ncu2_$TAG.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/prof_eig_$TAG.ncu-rep
