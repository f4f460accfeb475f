#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainloop_gpu.py -x -q > gpurun_out/r02_gputest_20.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_gputest_20.log | cut -c1-200
FCWDM_BENCH_STEP_TRACE=1 timeout 600 python bench.py --workload train --batch 2 --steps 14 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab10.json 2> gpurun_out/r02_ab10.err
python -c "import json; d=json.load(open('gpurun_out/r02_ab10.json')); print('train:', round(d['value'],2), round(d['e2e']['value'],2))"
grep "per-step" gpurun_out/r02_ab10.err
