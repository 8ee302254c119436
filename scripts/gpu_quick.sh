#!/bin/bash
set -u
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log; fi
for wl in cfg4_100k cfg3_20k cfg2_5k cfg1_500; do echo "== $wl"; timeout 300 python bench.py --workload $wl --skip-e2e 2>&1 | tail -1; done
