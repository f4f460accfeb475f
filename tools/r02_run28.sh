#!/bin/bash
mkdir -p gpurun_out
for v in "1 1" "1 3" "1 6"; do
  set -- $v
  FCWDM_BENCH_STEP_TRACE=1 FCWDM_BENCH_LAG=$1 FCWDM_BENCH_E2E_WARM=$2 timeout 600 python bench.py --workload train --batch 2 --steps 14 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab9.json 2> gpurun_out/r02_ab9.err
  python -c "import json; d=json.load(open('gpurun_out/r02_ab9.json')); print('lag $1 warm $2:', round(d['value'],2), round(d['e2e']['value'],2))"
  grep "per-step" gpurun_out/r02_ab9.err
done
