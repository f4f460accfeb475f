#!/bin/bash
# round-2 GPU call 2: chain kernel tests, the reworked config tests, chain on/off A/B of the headline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3d_chain_gpu.py -q -s -x > gpurun_out/r02_chain_test_1.log 2>&1; echo "chain pytest rc=$?"
tail -25 gpurun_out/r02_chain_test_1.log | cut -c1-300
timeout 900 python -m pytest tests/test_configs_gpu.py tests/test_trainloop_gpu.py tests/test_wunet_gpu.py -q -s > gpurun_out/r02_gputest_2.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02_gputest_2.log | cut -c1-300
FCWDM_NO_CHAIN=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_nochain.json 2> gpurun_out/r02_bench_nochain.err; echo "bench nochain rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_chain.json 2> gpurun_out/r02_bench_chain.err; echo "bench chain rc=$?"
python - <<'PY'
import json
for n in ("nochain","chain"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{n}.json"))
        print(n, d["value"], d["e2e"]["value"], d["clocks"], d["gpu_launches"])
    except Exception as e:
        print(n, "failed", e)
PY
tail -5 gpurun_out/r02_bench_chain.err
