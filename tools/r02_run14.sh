#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv3d_gpu.py tests/test_wunet_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/r02_gputest_12.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_gputest_12.log | cut -c1-300
for v in base exp1 exp2 exp3; do
  unset FCWDM_LIB_PATH
  case $v in exp*) export FCWDM_LIB_PATH=$PWD/tools/_bin/libfcwdm_$v.so;; esac
  echo "== $v"; timeout 300 python tools/chain_probe.py 2>&1 | cut -c1-150
done > gpurun_out/r02_chain_probe_5.txt 2>&1
unset FCWDM_LIB_PATH
cat gpurun_out/r02_chain_probe_5.txt
python tools/step_probe.py 3 > gpurun_out/r02_step_plain4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_4.csv python tools/step_probe.py 3 > gpurun_out/r02_ncu4.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_4.csv p_sample_step | head -12
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train_2.csv python tools/train_probe.py 3 2 > gpurun_out/r02_ncu_train2.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_train_2.csv adamw > gpurun_out/r02_train_agg2.txt; head -16 gpurun_out/r02_train_agg2.txt
