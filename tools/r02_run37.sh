#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv3d_gpu.py tests/test_wunet_gpu.py tests/test_backward_gpu.py -x -q --timeout 120 > gpurun_out/r02_gputest_22.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_22.log | cut -c1-300
for v in new old new_b old_b; do
  unset FCWDM_LIB_PATH
  case $v in old*) export FCWDM_LIB_PATH=$PWD/tools/_bin/libfcwdm_nocoal.so;; esac
  timeout 300 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ab12_$v.json 2> gpurun_out/r02_ab12_$v.err; echo "bench $v rc=$?"
done
for v in new old; do
  unset FCWDM_LIB_PATH
  case $v in old*) export FCWDM_LIB_PATH=$PWD/tools/_bin/libfcwdm_nocoal.so;; esac
  timeout 300 python bench.py --workload train --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab12_train_$v.json 2> gpurun_out/r02_ab12_train_$v.err; echo "bench train $v rc=$?"
done
python - <<'PY'
import json
for n in ("new","old","new_b","old_b","train_new","train_old"):
    try:
        d=json.load(open(f"gpurun_out/r02_ab12_{n}.json"))
        r=d.get("roofline",{})
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], r.get("us_by_variant"))
    except Exception as e:
        print(n, "failed", e)
PY
