#!/bin/bash
mkdir -p gpurun_out
for v in alias alias_module; do timeout 100 python scripts/train_graph_debug3.py $v 2>&1 | grep "^\[" ; done | tee gpurun_out/r2_train_graph_debug3.log
