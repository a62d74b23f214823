#!/bin/bash
mkdir -p gpurun_out
VBG_TRAIN_PROFILE=1 timeout 400 python scripts/train_bench.py cfg2 3 2>&1 | grep -v Warning | tail -56 | tee gpurun_out/job46_train_profile.log
