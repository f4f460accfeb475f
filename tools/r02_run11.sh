#!/bin/bash
# full GPU suite + the default bench line (both arms), as the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_9.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_2.json 2> gpurun_out/r02_bench_2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_2.json 2> gpurun_out/r02_bench_ref_2.err; echo "bench ref rc=$?"
cut -c1-600 gpurun_out/r02_bench_2.json; cut -c1-400 gpurun_out/r02_bench_ref_2.json
