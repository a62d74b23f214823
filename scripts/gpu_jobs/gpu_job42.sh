#!/bin/bash
# training-step kernels: unit parity vs torch float64 autograd
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_autograd.py -q --timeout 100 -o faulthandler_timeout=90 2>&1 | tail -40 | tee gpurun_out/job42_pytest.log
