#!/bin/bash
# round-2 GPU call 1: full GPU test suite, default bench line, reference arm, launch list, HBM-kernel ncu captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r02_gputest_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_1.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_1.json 2> gpurun_out/r02_bench_ref_1.err; echo "ref rc=$?"
python tools/step_probe.py 3 > gpurun_out/r02_step_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_1.csv python tools/step_probe.py 3 > gpurun_out/r02_ncu1.log 2>&1
python tools/step_probe.py 2 > gpurun_out/r02_step_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:gn_stats|gn_apply|dwt3d_cl|idwt3d_cl|p_sample_step" -s 36 -c 40 -o gpurun_out/r02_hbm_kernels python tools/step_probe.py 2 > gpurun_out/r02_ncu2.log 2>&1
python tools/step_probe.py 2 > gpurun_out/r02_step_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:conv3d_pair_kernel" -s 20 -c 4 -o gpurun_out/r02_pair_instep python tools/step_probe.py 2 > gpurun_out/r02_ncu3.log 2>&1
tail -5 gpurun_out/r02_gputest_1.log; cat gpurun_out/r02_bench_1.json | head -c 3000
