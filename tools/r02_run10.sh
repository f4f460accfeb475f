#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trainloop_gpu.py tests/test_train_gpu.py tests/test_wunet_gpu.py tests/test_sample_driver_gpu.py tests/test_unet_gpu.py -q -s > gpurun_out/r02_gputest_8.log 2>&1; echo "pytest rc=$?"
grep -n "TrainLoop loss\|update of\|passed\|failed\|^E \|plain unet" gpurun_out/r02_gputest_8.log | cut -c1-250 | head -40
timeout 600 python tools/driver_probe.py > gpurun_out/r02_driver_probe_v1.txt 2>&1; tail -5 gpurun_out/r02_driver_probe_v1.txt
