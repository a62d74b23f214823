#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_autograd.py tests/test_gpu_train_ops.py tests/test_gpu_train_step.py -q --timeout 200 -o faulthandler_timeout=180 2>&1 | tail -25 | tee gpurun_out/job47_pytest.log
VBG_TRAIN_PROFILE=1 timeout 400 python scripts/train_bench.py cfg2 5 2>&1 | grep -v Warning | tail -40 | tee gpurun_out/job47_train_profile.log
