#!/bin/bash
# multi-GPU pass: N = $1 ranks; replicas (weak) and sketch-column sharding (strong), plus the batch config
set -u
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 150 --warmup 3 --no-cpu 2>gpurun_out/multi_rep_$N.err | tail -1 > gpurun_out/multi_replicas_$N.json
$TR --master-port 29522 bench.py --gpus $N --steps 150 --warmup 3 --parallel sketch 2>gpurun_out/multi_sk_$N.err | tail -1 > gpurun_out/multi_sketch_$N.json
$TR --master-port 29523 bench.py --gpus $N --workload cfg5_batch --instances 1024 --steps 150 2>gpurun_out/multi_b_$N.err | tail -1 > gpurun_out/multi_batch_$N.json
for f in replicas sketch batch; do python -c "
import json,sys
d=json.load(open('gpurun_out/multi_${f}_$N.json'))
print('$f N=$N value=%.1f %s ms/step=%.4f scaling=%s e2e=%s' % (d['value'], d['unit'], d['ms_per_step'], d.get('scaling'), (d.get('e2e') or {}).get('value')))
" 2>&1 | tail -1; done
