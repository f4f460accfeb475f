#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_trainloop_gpu.py tests/test_conv3d_chain_gpu.py -x -q > gpurun_out/r02_gputest_19.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_18.log | cut -c1-300
for v in a b; do
timeout 600 python bench.py --workload train --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab7_train_$v.json 2> gpurun_out/r02_ab7_train_$v.err; echo "bench train rc=$?"
done
python - <<'PY'
import json
for n in ("train_a","train_b"):
    try:
        d=json.load(open(f"gpurun_out/r02_ab7_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_5.csv python tools/train_probe.py 3 2 > gpurun_out/r02_ncu_train5.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_train_5.csv adamw > gpurun_out/r02_train_agg5.txt; grep "finalize\|TOTAL\|pack_all" gpurun_out/r02_train_agg5.txt
