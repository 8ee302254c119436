#!/bin/bash
# device plan builder: equality tests, stage times next to the host builder's, end-to-end effect
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plan_builder.py -q 2>&1 | tail -15 | tee gpurun_out/pb_tests.log
for b in ${BUILDERS:-host device}; do
  echo "==== builder $b"
  SIGSDP_PLAN_BUILDER=$b SIGSDP_PLAN_TIMING=1 timeout 300 python scripts/e2e_breakdown.py cfg4_100k 2>&1 | tail -60 > gpurun_out/pb_e2e_$b.log
  grep -E "e2e |python laps" gpurun_out/pb_e2e_$b.log
done
for b in ${BUILDERS:-host device}; do
  echo "==== bench builder $b"
  SIGSDP_PLAN_BUILDER=$b timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu 2>gpurun_out/pb_bench_$b.err | tail -1 > gpurun_out/pb_bench_$b.json
  python -c "
import json; d=json.load(open('gpurun_out/pb_bench_$b.json')); print(d['value'], d['e2e']['value'], d['e2e']['breakdown_ms'])"
done
