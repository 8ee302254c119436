#!/bin/bash
set -u
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for gap in 0 2 4 8; do
  echo "== cfg4 gap=$gap"; SIGSDP_RUN_GAP=$gap timeout 300 python bench.py --workload cfg4_100k --order 1 --skip-e2e 2>&1 | tail -1
done
