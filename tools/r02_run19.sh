#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3d_chain_gpu.py tests/test_wunet_gpu.py -x -q > gpurun_out/r02_gputest_16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_16.log | cut -c1-300
timeout 300 python tools/chain_probe.py 2>&1 | cut -c1-150 > gpurun_out/r02_chain_probe_8.txt; cat gpurun_out/r02_chain_probe_8.txt
timeout 600 python -m pytest tests/test_configs_gpu.py tests/test_reference_scripts_gpu.py -x -q > gpurun_out/r02_gputest_17.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_17.log | cut -c1-300
for v in a b; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ab5_$v.json 2> gpurun_out/r02_ab5_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("a","b"):
    try:
        d=json.load(open(f"gpurun_out/r02_ab5_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"], d["config"]["output_finite"])
    except Exception as e:
        print(n, "failed", e)
PY
