#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv3d_chain_gpu.py tests/test_wunet_gpu.py tests/test_train_gpu.py -x -q > gpurun_out/r02_gputest_14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_14.log | cut -c1-300
timeout 300 python tools/chain_probe.py 2>&1 | cut -c1-150 > gpurun_out/r02_chain_probe_6.txt; cat gpurun_out/r02_chain_probe_6.txt
for v in new ahead208 new_b ahead208_b; do
  unset FCWDM_LIB_PATH
  case $v in ahead208*) export FCWDM_LIB_PATH=$PWD/tools/_bin/libfcwdm_ahead208.so;; esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ab3_$v.json 2> gpurun_out/r02_ab3_$v.err; echo "bench $v rc=$?"
done
unset FCWDM_LIB_PATH
timeout 600 python bench.py --workload train --batch 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab3_train.json 2> gpurun_out/r02_ab3_train.err; echo "bench train rc=$?"
python - <<'PY'
import json
for n in ("new","ahead208","new_b","ahead208_b","train"):
    try:
        d=json.load(open(f"gpurun_out/r02_ab3_{n}.json"))
        r=d.get("roofline",{})
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"], r.get("us_by_variant"))
    except Exception as e:
        print(n, "failed", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_3.csv python tools/train_probe.py 3 2 > gpurun_out/r02_ncu_train3.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_train_3.csv adamw > gpurun_out/r02_train_agg3.txt; grep "pack_all\|TOTAL\|direct_copy" gpurun_out/r02_train_agg3.txt
