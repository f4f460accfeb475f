#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/t_sweep.py > gpurun_out/r02_t_sweep.md 2>&1; echo "t_sweep rc=$?"; tail -8 gpurun_out/r02_t_sweep.md
timeout 400 python bench.py --model unet --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_unet.json 2> gpurun_out/r02_bench_unet.err; echo "unet rc=$?"
timeout 400 python bench.py --model unet --workload train --batch 2 --no-cpu-baseline > gpurun_out/r02_bench_train_unet.json 2> gpurun_out/r02_bench_train_unet.err; echo "unet train rc=$?"
timeout 400 python bench.py --workload train --batch 1 --no-cpu-baseline > gpurun_out/r02_bench_train_b1.json 2> gpurun_out/r02_bench_train_b1.err; echo "train b1 rc=$?"
python - <<'PY'
import json
for n in ("bench_unet","bench_train_unet","bench_train_b1"):
    try:
        d=json.load(open(f"gpurun_out/r02_{n}.json")); print(n, round(d["value"],2), round(d["e2e"]["value"],2), d["clocks"]["sm_mhz"], d["config"]["workload"][:60])
    except Exception as e: print(n, "failed", e)
PY
