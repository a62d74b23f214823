#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_forward.py -q --timeout 200 -o faulthandler_timeout=180 2>&1 | tail -15 | tee gpurun_out/job45_pytest.log
timeout 300 python scripts/train_bench.py cfg2 5 2>&1 | tail -12 | tee gpurun_out/job45_train_bench.log
