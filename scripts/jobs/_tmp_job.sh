#!/bin/bash
mkdir -p gpurun_out
bash scripts/jobs/tests.sh
cp gpurun_out/pytest_gpu.log gpurun_out/r2_pytest_gpu_full_c.log
bash scripts/jobs/train_profile.sh > gpurun_out/r2_train_profile_c.txt 2>&1; head -45 gpurun_out/r2_train_profile_c.txt | cut -c1-160
bash scripts/jobs/ncu_launches.sh > gpurun_out/r2_launches_summary.txt 2>&1
head -30 gpurun_out/r2_launches_summary.txt
