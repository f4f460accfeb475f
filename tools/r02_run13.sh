#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_trainloop_gpu.py -x -q > gpurun_out/r02_gputest_11.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_11.log | cut -c1-300
for v in fused nofused fused_b nofused_b; do
  unset FCWDM_NO_FUSED_COLSUM
  case $v in nofused*) export FCWDM_NO_FUSED_COLSUM=1;; esac
  timeout 600 python bench.py --workload train --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_train_ab_$v.json 2> gpurun_out/r02_train_ab_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("fused","nofused","fused_b","nofused_b"):
    try:
        d=json.load(open(f"gpurun_out/r02_train_ab_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
