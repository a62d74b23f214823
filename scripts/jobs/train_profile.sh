#!/bin/bash
# forward / backward / optimizer split of the cfg2 training step and its kernel table
mkdir -p gpurun_out
VBG_TRAIN_PROFILE=1 timeout 600 python scripts/train_bench.py cfg2 8 > gpurun_out/train_profile.log 2>&1; echo "train_bench exit $?"
grep -v "^step" gpurun_out/train_profile.log | head -70
