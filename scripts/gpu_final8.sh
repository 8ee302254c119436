#!/bin/bash
# 8-GPU (or $1-GPU) closing pass: the driver's command line on the row-sharded solver, the 150-step line, the batch split
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/f_rows20_$N.err | tail -1 > gpurun_out/f_rows20_$N.json
timeout 600 $TR --master-port 29552 bench.py --gpus $N --steps 150 --warmup 3 --no-cpu 2>gpurun_out/f_rows150_$N.err | tail -1 > gpurun_out/f_rows150_$N.json
timeout 600 $TR --master-port 29553 bench.py --gpus $N --workload cfg5_batch --instances 1024 --steps 150 2>gpurun_out/f_batch_$N.err | tail -1 > gpurun_out/f_batch_$N.json
python - <<PY
import json
for f in ("f_rows20_$N", "f_rows150_$N", "f_batch_$N"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.1f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("breakdown_ms", {}).get("state_process"), d.get("config", {}).get("blocks_per_instance"), (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(f, "no line", e)
PY
tail -2 gpurun_out/f_rows20_$N.err
