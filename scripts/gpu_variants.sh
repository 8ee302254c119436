#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for gap in 2 8 16; do
  echo "== gap=$gap cfg4"; SIGSDP_RUN_GAP=$gap timeout 300 python bench.py --workload cfg4_100k --skip-e2e 2>&1 | tail -1 | cut -c1-760
done
