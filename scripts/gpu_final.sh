#!/bin/bash
# round-end style pass: GPU tests, smoke, the default bench line, the reference arm, other configs, batch
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err; cat gpurun_out/bench_main.json
timeout 600 python bench.py --impl reference --steps 150 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
for wl in cfg3_20k cfg2_5k cfg1_500; do timeout 600 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; cat gpurun_out/bench_$wl.json; done
timeout 900 python bench.py --workload cfg5_batch --instances 1024 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; cat gpurun_out/bench_cfg5.json
