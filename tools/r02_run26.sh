#!/bin/bash
mkdir -p gpurun_out
python tools/step_probe.py 3 8 > gpurun_out/r02_step_plain_b8.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b8.csv python tools/step_probe.py 3 8 > gpurun_out/r02_ncu_b8.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_b8.csv p_sample_step | head -24
python tools/agg_launches.py gpurun_out/r02_launches_b8.csv p_sample_step --list | sed -n 8,70p
timeout 600 python bench.py --workload train --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab8_train.json 2> gpurun_out/r02_ab8_train.err; echo "bench train rc=$?"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_3.json 2> gpurun_out/r02_bench_3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_ab8_train.json")); print("train standalone", d["value"], d["e2e"]["value"])
d=json.load(open("gpurun_out/r02_bench_3.json")); t=d["secondary"]["train"]; print("default", d["value"], d["e2e"]["value"], "train", t["value"], t["e2e"]["value"])
PY
