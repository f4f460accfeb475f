#!/bin/bash
mkdir -p gpurun_out
FCWDM_LIB_PATH=$PWD/fast-cwdm_b200/fcwdm/libfcwdm_trace.so timeout 300 python tools/chain_trace.py 2>&1 | tee gpurun_out/r02_chain_trace_1.txt | tail -60
timeout 900 python -m pytest tests/test_configs_gpu.py::test_config3_batch8_equals_batch1 tests/test_reference_scripts_gpu.py -q -s -x > gpurun_out/r02_gputest_4.log 2>&1; echo "pytest rc=$?"
grep -n "random weights\|contractive\|sample.nii\|passed\|failed\|^E " gpurun_out/r02_gputest_4.log | cut -c1-300 | head -30
