#!/bin/bash
mkdir -p gpurun_out
for v in base streamio base_b streamio_b; do
  unset FCWDM_LIB_PATH
  case $v in streamio*) export FCWDM_LIB_PATH=$PWD/tools/_bin/libfcwdm_streamio.so;; esac
  timeout 300 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_ab13_$v.json 2> gpurun_out/r02_ab13_$v.err; echo "bench $v rc=$?"
done
FCWDM_LIB_PATH=$PWD/tools/_bin/libfcwdm_streamio.so timeout 200 python -m pytest tests/test_conv3d_gpu.py -x -q --timeout 120 -k pair 2>&1 | tail -1
python - <<'PY'
import json
for n in ("base","streamio","base_b","streamio_b"):
    try:
        d=json.load(open(f"gpurun_out/r02_ab13_{n}.json"))
        r=d.get("roofline",{})
        print(n, round(d["value"],3), round(d["e2e"]["value"],3), d["clocks"]["sm_mhz"], r.get("us_by_variant"))
    except Exception as e:
        print(n, "failed", e)
PY
