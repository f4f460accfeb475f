#!/bin/bash
mkdir -p gpurun_out
for v in rows20k rows130k nochain rows130k_b rows20k_b; do
  unset FCWDM_NO_CHAIN; unset FCWDM_CHAIN_ROWS
  case $v in nochain) export FCWDM_NO_CHAIN=1;; rows130k*) export FCWDM_CHAIN_ROWS=130000;; esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_$v.json 2> gpurun_out/r02_bench_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("rows20k","rows130k","nochain","rows130k_b","rows20k_b"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"], d["config"]["output_finite"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02_bench_rows130k.err
FCWDM_CHAIN_ROWS=130000 timeout 900 python -m pytest tests/test_wunet_gpu.py tests/test_conv3d_chain_gpu.py -q > gpurun_out/r02_gputest_7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_gputest_7.log | cut -c1-200
export FCWDM_CHAIN_ROWS=130000
python tools/step_probe.py 3 > gpurun_out/r02_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_3.csv python tools/step_probe.py 3 > gpurun_out/r02_ncu1.log 2>&1
