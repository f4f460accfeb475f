#!/bin/bash
# verification pass: the GPU suite (bounded), smoke, both bench arms
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 240 > gpurun_out/r02_gputest_final2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_final2.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final2.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke_final2.log
timeout 600 python bench.py > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_ref_final2.json 2> gpurun_out/r02_bench_ref_final2.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_final2.json")); t=d["secondary"]["train"]; b=d["secondary"]["batch8"]
print("default", round(d["value"],2), round(d["e2e"]["value"],2), d["clocks"]["sm_mhz"], "roofline", round(d["roofline"]["frac"],3), "b8", round(b["value"],2), "train", round(t["value"],2), round(t["e2e"]["value"],2))
r=json.load(open("gpurun_out/r02_bench_ref_final2.json")); print("ref", r["value"], r["ms_per_step"])
PY
