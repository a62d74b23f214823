#!/bin/bash
mkdir -p gpurun_out
VBG_PDL=0 timeout 300 python scripts/train_graph_debug.py > gpurun_out/r2_train_graph_debug.log 2>&1; echo "exit $?" >> gpurun_out/r2_train_graph_debug.log
grep -v Warning gpurun_out/r2_train_graph_debug.log | tail -30
