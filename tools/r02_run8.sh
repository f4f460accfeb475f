#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3d_chain_gpu.py -q -x -s > gpurun_out/r02_chain_test_6.log 2>&1; echo "chain pytest rc=$?"; tail -4 gpurun_out/r02_chain_test_6.log | cut -c1-250
for v in nochain chain; do
  unset FCWDM_NO_CHAIN
  case $v in nochain) export FCWDM_NO_CHAIN=1;; esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_$v.json 2> gpurun_out/r02_bench_$v.err; echo "bench $v rc=$?"
done
unset FCWDM_NO_CHAIN
python - <<'PY'
import json
for n in ("nochain","chain"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r02_bench_chain.err
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_6.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_gputest_6.log | cut -c1-300
python tools/step_probe.py 3 > gpurun_out/r02_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_2.csv python tools/step_probe.py 3 > gpurun_out/r02_ncu1.log 2>&1
