#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_haar_gpu.py tests/test_wunet_gpu.py tests/test_conv3d_chain_gpu.py -x -q > gpurun_out/r02_gputest_21.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_21.log | cut -c1-300
for v in new old new_b old_b; do
  unset FCWDM_NO_FUSED_STATS_HAAR
  case $v in old*) export FCWDM_NO_FUSED_STATS_HAAR=1;; esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ab11_$v.json 2> gpurun_out/r02_ab11_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("new","old","new_b","old_b"):
    try:
        d=json.load(open(f"gpurun_out/r02_ab11_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"], d["config"]["output_finite"])
    except Exception as e:
        print(n, "failed", e)
PY
