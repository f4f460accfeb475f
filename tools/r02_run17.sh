#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_15.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_15.log | cut -c1-300
timeout 300 python tools/chain_probe.py 2>&1 | cut -c1-150 > gpurun_out/r02_chain_probe_7.txt; cat gpurun_out/r02_chain_probe_7.txt
for v in base fstats base_b fstats_b; do
  unset FCWDM_FUSED_STATS
  case $v in fstats*) export FCWDM_FUSED_STATS=1;; esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ab4_$v.json 2> gpurun_out/r02_ab4_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("base","fstats","base_b","fstats_b"):
    try:
        d=json.load(open(f"gpurun_out/r02_ab4_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"], d["config"]["output_finite"])
    except Exception as e:
        print(n, "failed", e)
PY
