#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3d_chain_gpu.py -q -x > gpurun_out/r02_chain_test_5.log 2>&1; echo "chain pytest rc=$?"; tail -3 gpurun_out/r02_chain_test_4.log | cut -c1-250
FCWDM_LIB_PATH=$PWD/fast-cwdm_b200/fcwdm/libfcwdm_trace.so timeout 300 python tools/chain_trace.py 2>&1 | tee gpurun_out/r02_chain_trace_4.txt | grep -v "^| [1-4] " | tail -28
timeout 300 python tools/chain_probe.py 2>&1 | tail -9 | tee gpurun_out/r02_chain_probe_4.txt
for v in nochain chain; do
  unset FCWDM_NO_CHAIN; unset FCWDM_FUSED_STATS
  case $v in nochain) export FCWDM_NO_CHAIN=1;; fstats) export FCWDM_FUSED_STATS=1;; esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_$v.json 2> gpurun_out/r02_bench_$v.err; echo "bench $v rc=$?"
done
unset FCWDM_NO_CHAIN; unset FCWDM_FUSED_STATS
python - <<'PY'
import json
for n in ("nochain","chain"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
