#!/bin/bash
# r2 job 9: training-step tests + timing/profile after the tcgen05 attention backward
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_zz_syncbn.py -m gpu -q -x --timeout 300 2>&1 | tail -5 > gpurun_out/r2_pytest_train_step.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_train_step.log; cat gpurun_out/r2_pytest_train_step.log
VBG_TRAIN_PROFILE=1 timeout 300 python scripts/train_bench.py cfg2 5 > gpurun_out/r2_train_profile_a.log 2>&1; echo "train_bench exit $?"
grep -v "^step" gpurun_out/r2_train_profile_a.log | head -60
