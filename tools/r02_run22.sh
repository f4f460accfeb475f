#!/bin/bash
# final 1-GPU pass of the round: the suite, smoke, both bench arms, the driver probe, launch lists and ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_final.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref_final.json 2> gpurun_out/r02_bench_ref_final.err; echo "bench ref rc=$?"
cut -c1-400 gpurun_out/r02_bench_final.json
timeout 600 python tools/driver_probe.py 32 12 4 > gpurun_out/r02_driver_probe_v2.txt 2>&1; tail -3 gpurun_out/r02_driver_probe_v2.txt
python tools/step_probe.py 3 > gpurun_out/r02_step_plain5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_5.csv python tools/step_probe.py 3 > gpurun_out/r02_ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3d_chain_kernel -s 4 -c 4 -o gpurun_out/r02_chain_instep -f python tools/step_probe.py 2 > gpurun_out/r02_ncu6.log 2>&1; echo "ncu chain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3d_igemm_kernel -s 12 -c 6 -o gpurun_out/r02_igemm_instep -f python tools/step_probe.py 2 > gpurun_out/r02_ncu7.log 2>&1; echo "ncu igemm rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
