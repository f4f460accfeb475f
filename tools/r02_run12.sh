#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainloop_gpu.py -q -s -k three_steps > gpurun_out/r02_gputest_10.log 2>&1; echo "pytest rc=$?"
grep -n "update norm\|update of\|passed\|failed" gpurun_out/r02_gputest_10.log | cut -c1-200
# training launch list at batch 2 (the bench's training configuration)
python tools/train_probe.py 3 2 > gpurun_out/r02_train_plain.log 2>&1 && tail -3 gpurun_out/r02_train_plain.log && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_1.csv python tools/train_probe.py 3 2 > gpurun_out/r02_ncu_train1.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_train_1.csv adamw | head -30
# same-box A/B: fused statistics in the single-CTA kernel's epilogue
for v in base fstats base_b fstats_b; do
  unset FCWDM_FUSED_STATS
  case $v in fstats*) export FCWDM_FUSED_STATS=1;; esac
  timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ab_$v.json 2> gpurun_out/r02_ab_$v.err; echo "bench $v rc=$?"
done
python - <<'PY'
import json
for n in ("base","fstats","base_b","fstats_b"):
    try:
        d=json.load(open(f"gpurun_out/r02_ab_{n}.json"))
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], d["gpu_launches"], d["config"]["output_finite"])
    except Exception as e:
        print(n, "failed", e)
PY
