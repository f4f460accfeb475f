#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/chain_debug.py 2>&1 | tail -22
timeout 600 python -m pytest tests/test_conv3d_chain_gpu.py -q -s -x > gpurun_out/r02_chain_test_2.log 2>&1; echo "chain pytest rc=$?"; tail -5 gpurun_out/r02_chain_test_2.log | cut -c1-250
for cs in 4 2; do FCWDM_CHAIN_CLUSTER=$cs timeout 300 python tools/chain_probe.py 2>&1 | tail -9; done | tee gpurun_out/r02_chain_probe_1.txt
FCWDM_CHAIN_CLUSTER=4 FCWDM_CHAIN_SPLIT=2 timeout 300 python tools/chain_probe.py 2>&1 | tail -9 | tee -a gpurun_out/r02_chain_probe_1.txt
for v in nochain chain chain4 chain2; do
  case $v in nochain) export FCWDM_NO_CHAIN=1;; chain) unset FCWDM_NO_CHAIN; unset FCWDM_CHAIN_CLUSTER;; chain4) export FCWDM_CHAIN_CLUSTER=4;; chain2) export FCWDM_CHAIN_CLUSTER=2;; esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_$v.json 2> gpurun_out/r02_bench_$v.err; echo "bench $v rc=$?"
done
unset FCWDM_NO_CHAIN; unset FCWDM_CHAIN_CLUSTER
python - <<'PY'
import json
for n in ("nochain","chain","chain4","chain2"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r02_gputest_3.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r02_gputest_3.log | cut -c1-300
