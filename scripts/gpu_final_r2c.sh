#!/bin/bash
# closing pass of round 2 (one GPU): GPU tests, smoke, the driver's two command lines (ours + reference arm),
# the default line (150 steps), the other single-GPU configurations and the batch mode
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/fz_pytest_gpu.log; tail -3 gpurun_out/fz_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/fz_ref.json 2> gpurun_out/fz_ref.err; cat gpurun_out/fz_ref.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/fz_driver.json 2> gpurun_out/fz_driver.err; head -c 400 gpurun_out/fz_driver.json; echo
timeout 900 python bench.py > gpurun_out/fz_main.json 2> gpurun_out/fz_main.err; head -c 300 gpurun_out/fz_main.json; echo
for wl in cfg3_20k cfg2_5k cfg1_500; do timeout 600 python bench.py --workload $wl > gpurun_out/fz_$wl.json 2> gpurun_out/fz_$wl.err; head -c 200 gpurun_out/fz_$wl.json; echo; done
timeout 900 python bench.py --workload cfg5_batch --instances 1024 > gpurun_out/fz_cfg5.json 2> gpurun_out/fz_cfg5.err; head -c 300 gpurun_out/fz_cfg5.json; echo
python - <<'PY'
import json
for f in ("fz_driver", "fz_main", "fz_cfg3_20k", "fz_cfg2_5k", "fz_cfg1_500"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); e = d["e2e"]; b = e["breakdown_ms"]
        print(f, "value %.1f frac %.3f e2e %.1f" % (d["value"], d["roofline"]["frac"], e["value"]), {k: round(v, 1) for k, v in b.items() if k != "eig"},
              "eig", (b.get("eig") or {}).get("total_s"), "filter" in (b.get("eig") or {}), "parity", (d.get("parity") or {}).get("ok"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d["clocks"])
    except Exception as ex:
        print(f, "no line", ex)
PY
