#!/bin/bash
# GPU job 53: host-side (cProfile) profile of the training step.
mkdir -p gpurun_out
timeout 200 python scripts/train_host_profile.py cfg2 2>&1 | grep -v -E "Warning|warn\(" > gpurun_out/train_host_profile.log
head -120 gpurun_out/train_host_profile.log
