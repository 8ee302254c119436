#!/bin/bash
# one GPU-box pass: parity tests, smoke, bench lines, ncu launch list.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for wl in cfg4_100k cfg3_20k cfg2_5k cfg1_500; do
  timeout 600 python bench.py --workload $wl --order 0 > gpurun_out/bench_${wl}_o0.json 2> gpurun_out/bench_${wl}_o0.err || tail -5 gpurun_out/bench_${wl}_o0.err
done
timeout 600 python bench.py --workload cfg4_100k --order 1 --no-cpu > gpurun_out/bench_cfg4_100k_o1.json 2> gpurun_out/bench_cfg4_100k_o1.err || tail -5 gpurun_out/bench_cfg4_100k_o1.err
timeout 600 python bench.py --workload cfg4_100k --order 32 --no-cpu > gpurun_out/bench_cfg4_100k_o32.json 2> gpurun_out/bench_cfg4_100k_o32.err
timeout 600 python bench.py --impl reference --steps 150 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_*.json
