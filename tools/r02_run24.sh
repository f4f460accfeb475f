#!/bin/bash
mkdir -p gpurun_out
export FCWDM_FUSED_STATS=1
python tools/step_probe.py 3 > gpurun_out/r02_step_plain6.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_6.csv python tools/step_probe.py 3 > gpurun_out/r02_ncu8.log 2>&1
python tools/agg_launches.py gpurun_out/r02_launches_6.csv p_sample_step | head -20
python tools/agg_launches.py gpurun_out/r02_launches_6.csv p_sample_step --list | sed -n 14,50p
