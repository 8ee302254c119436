#!/bin/bash
set -u
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log; fi
for bps in 2; do
  echo "== cfg4 bps=$bps"; SIGSDP_BLOCKS_PER_SM=$bps timeout 300 python bench.py --workload cfg4_100k --order 1 --skip-e2e 2>&1 | tail -1
done
for wl in cfg3_20k cfg2_5k cfg1_500; do echo "== $wl"; timeout 300 python bench.py --workload $wl --skip-e2e 2>&1 | tail -1; done
