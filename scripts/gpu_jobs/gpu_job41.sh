#!/bin/bash
# conv weight gradient (MN-major tcgen05) parity
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_autograd.py -x -q --timeout 100 -o faulthandler_timeout=90 2>&1 | tail -15 | tee gpurun_out/job41_pytest.log
