#!/bin/bash
# one profiling line of the headline workload (device-resident timing + block-0 cycle counters)
set -u
timeout 300 python bench.py --workload cfg4_100k --skip-e2e 2>&1 | tail -1
