#!/bin/bash
# Chebyshev-filtered final factor: tests, then the bench line (driver command line and K = 150)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "filtered or lanczos" 2>&1 | tail -15 | tee gpurun_out/ef_tests.log
timeout 600 python -m pytest tests/test_gpu_plan_builder.py -q 2>&1 | tail -3
for K in 20 150; do
  timeout 400 python bench.py --steps $K --warmup 5 --no-cpu 2>gpurun_out/ef_bench_$K.err | tail -1 > gpurun_out/ef_bench_$K.json
  python -c "
import json; d=json.load(open('gpurun_out/ef_bench_$K.json')); print($K, d['value'], d['e2e']['value'], d['e2e']['breakdown_ms'], d.get('parity'))"
done
SIGSDP_PLAN_TIMING=1 timeout 300 python scripts/e2e_breakdown.py cfg4_100k 2>&1 | tail -40 > gpurun_out/ef_e2e.log
grep -E "e2e |python laps" gpurun_out/ef_e2e.log
