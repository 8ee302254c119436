#!/bin/bash
# row-sharded single-graph pass on N = $1 GPUs: the torchrun parity test, then bench.py --gpus N
# (cfg4, ONE graph across the ranks) and, with BIG=1, the 1M-node graph
set -u
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${TESTS:-1}" = "1" ]; then timeout 600 python -m pytest tests/test_row_shards.py -m gpu -q -x -k torchrun 2>&1 | tail -15; fi
timeout 600 $TR --master-port 29531 bench.py --gpus $N --steps ${STEPS:-150} --warmup 3 2>gpurun_out/rows_$N.err | tail -1 > gpurun_out/rows_$N.json
tail -5 gpurun_out/rows_$N.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/rows_$N.json'))
    print('rows N=$N value=%.1f ms/step=%.4f e2e=%.1f frac=%.3f parity=%s exchange=%s barrier_wait_ms=%s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('parity'), d.get('exchange'), d['detail']['barrier_wait_ms_max']))
    print(d['detail']['phase_us_rank0'], d['e2e']['breakdown_ms'])
except Exception as e: print('no line', e)
PY
if [ "${BIG:-0}" = "1" ]; then
timeout 900 $TR --master-port 29532 bench.py --gpus $N --workload cfg4x10_1m --steps 50 --warmup 3 --no-cpu 2>gpurun_out/rows1m_$N.err | tail -1 > gpurun_out/rows1m_$N.json
tail -3 gpurun_out/rows1m_$N.err; head -c 600 gpurun_out/rows1m_$N.json; echo
fi
