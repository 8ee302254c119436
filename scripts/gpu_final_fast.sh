#!/bin/bash
# shorter round-end pass: GPU tests, smoke, the default bench line (with its CPU baseline leg), the
# other configurations without their CPU legs, the batch configuration, the e2e breakdown
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; cat gpurun_out/bench_main.json
for wl in cfg3_20k cfg2_5k cfg1_500; do timeout 600 python bench.py --workload $wl --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; cat gpurun_out/bench_$wl.json; done
timeout 900 python bench.py --workload cfg5_batch --instances 1024 --no-cpu > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; cat gpurun_out/bench_cfg5.json
SIGSDP_PLAN_TIMING=1 timeout 300 python scripts/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1; grep -E "e2e|host plan" gpurun_out/e2e_breakdown.log
