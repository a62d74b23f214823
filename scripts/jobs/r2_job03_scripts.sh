#!/bin/bash
# r2 job 3: the reference's own eval_SROIE.py / train_SROIE.py through the drop-in
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_reference_scripts.py -m gpu -q -x -s --timeout 900 2>&1 | tail -80 > gpurun_out/r2_pytest_scripts.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_scripts.log; tail -60 gpurun_out/r2_pytest_scripts.log
