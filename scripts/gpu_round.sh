#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for K in 20 150; do
  timeout 400 python bench.py --steps $K --warmup 5 --no-cpu 2>gpurun_out/rd_bench_$K.err | tail -1 > gpurun_out/rd_bench_$K.json
  python -c "
import json; d=json.load(open('gpurun_out/rd_bench_$K.json')); b=d['e2e']['breakdown_ms']; print($K, round(d['value'],1), round(d['e2e']['value'],1), {k: round(v,1) for k,v in b.items() if k!='eig'})"
done
