#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/host_cost.py 2>/dev/null | tee gpurun_out/host_cost.log
