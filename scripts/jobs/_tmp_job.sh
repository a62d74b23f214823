#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -W "default::UserWarning" 2>&1 | grep -E "passed|failed|FAILED|capture failed|^E  " | tail -12 > gpurun_out/r2_pytest_gpu_final_$i.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_gpu_final_$i.log; cat gpurun_out/r2_pytest_gpu_final_$i.log
done
